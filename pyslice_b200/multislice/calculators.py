"""`MultisliceCalculator` with the reference's setup()/run() API
(reference src/multislice/calculators.py:39-250), driving the CUDA engine.

Differences from the reference that a caller can observe (all documented in DESIGN.md):
  * results are complex64 CUDA tensors instead of complex128 CPU tensors;
  * frames are processed in batches on the GPU; with `torch.distributed` initialised (one process
    per GPU) each rank propagates a contiguous block of frames and `WFData.shard` records it;
  * no `psi_data/` frame cache is written (the reference's cache key ignores atom positions);
  * new optional kwarg `layer_every`: also record the wave function after every n-th slice
    (layer axis of WFData); the default 0 reproduces the reference's single layer.
"""
from __future__ import annotations

import logging
import time
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import engine, hostmath
from ..postprocessing.wf_data import WFData
from .multislice import Probe, create_batched_probes
from .potentials import gridFromTrajectory
from .trajectory import Trajectory

logger = logging.getLogger(__name__)
complex_dtype = torch.complex64
float_dtype = torch.float32


@dataclass
class FrameShard:
    """Which frames of the trajectory this process holds (multi-GPU runs)."""
    rank: int
    world: int
    counts: List[int]          # frames per rank

    @property
    def start(self) -> int:
        return sum(self.counts[:self.rank])

    @property
    def total(self) -> int:
        return sum(self.counts)


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def split_frames(n_frames: int, world: int) -> List[int]:
    """Contiguous blocks, ragged when world does not divide n_frames."""
    return [n_frames // world + (1 if r < n_frames % world else 0) for r in range(world)]


class MultisliceCalculator:

    def __init__(self, device=None, force_cpu=False):
        """device: CUDA device (None = current).  `force_cpu=True` is refused: this engine has no
        CPU path (use the reference for that)."""
        if force_cpu:
            raise RuntimeError("pyslice_b200 has no CPU path; force_cpu is not supported")
        self.device = engine._device(device)
        logger.info(f"pyslice_b200 calculator initialized on device: {self.device}")
        self.element_map = {i + 1: s for i, s in enumerate(hostmath.ELEMENTS[:36])}

    def setup(
        self,
        trajectory: Trajectory,
        aperture: float = 0.0,
        voltage_eV: float = 60e3,
        defocus: float = 0.0,
        slice_thickness: float = 0.5,
        sampling: float = 0.1,
        probe_positions: Optional[List[Tuple[float, float]]] = None,
        batch_size: int = 10,
        save_path: Optional[Path] = None,
        cleanup_temp_files: bool = False,
        slice_axis: int = 2,
        layer_every: int = 0,
        shard_frames: Optional[bool] = None,
    ):
        """Same keyword arguments and defaults as the reference (calculators.py:96-109).
        `defocus`, `batch_size`, `save_path`, `cleanup_temp_files` are accepted and, as in the
        reference, have no effect on the result."""
        if slice_axis != 2:
            raise NotImplementedError("pyslice_b200 supports slice_axis=2 only")
        self.trajectory = trajectory
        self.aperture = aperture
        self.voltage_eV = voltage_eV
        self.defocus = defocus
        self.slice_thickness = slice_thickness
        self.sampling = sampling
        self.probe_positions = probe_positions
        self.save_path = save_path
        self.cleanup_temp_files = cleanup_temp_files
        self.slice_axis = slice_axis
        self.layer_every = int(layer_every)

        xs, ys, zs, lx, ly, lz = gridFromTrajectory(trajectory, sampling=sampling, slice_thickness=slice_thickness)
        self.xs, self.ys, self.zs = xs, ys, zs
        self.lx, self.ly, self.lz = lx, ly, lz
        self.nx, self.ny, self.nz = len(xs), len(ys), len(zs)
        self.dx = xs[1] - xs[0]
        self.dy = ys[1] - ys[0]

        if self.probe_positions is None:
            self.probe_positions = [(lx / 2, ly / 2)]
        self.base_probe = Probe(xs, ys, self.aperture, self.voltage_eV, device=self.device)
        self.n_frames = trajectory.n_frames
        self.n_probes = len(self.probe_positions)

        rank, world = _dist_info()
        if shard_frames is None:
            shard_frames = world > 1
        self.shard = FrameShard(rank, world, split_frames(self.n_frames, world)) if shard_frames else None

        self._plan = engine.make_plan(xs, ys, zs, trajectory.atom_types.tolist(), voltage_eV, device=self.device)
        self._probes = create_batched_probes(self.base_probe, self.probe_positions).array    # frame-invariant
        self.n_layers = engine.layer_count(self.nz, self.layer_every)
        self.wavefunction_data = None

    # -----------------------------------------------------------------------------------------
    def _local_frames(self):
        if self.shard is None:
            return 0, self.n_frames
        return self.shard.start, self.shard.start + self.shard.counts[self.shard.rank]

    def run(self, timer=None) -> WFData:
        """Propagate every probe through every frame (reference calculators.py:163-250).
        `timer`: optional engine.PhaseTimer collecting CUDA-event times per phase."""
        t_start = time.time()
        timer = timer or engine.NO_TIMER
        plan = self._plan
        f_lo, f_hi = self._local_frames()
        T_loc = f_hi - f_lo
        P, nx, ny = self.n_probes, self.nx, self.ny
        store = torch.empty((self.n_layers, P, T_loc, nx, ny), dtype=torch.complex64, device=self.device)
        fb, pb = engine.batch_sizes(plan, P, max(T_loc, 1))
        work = torch.empty((fb * min(pb, P), nx, ny), dtype=torch.complex64, device=self.device)
        tbuf = torch.empty((fb, plan.nz, nx, ny), dtype=torch.complex64, device=self.device)
        positions = self.trajectory.positions
        for b0 in range(0, T_loc, fb):
            nb = min(fb, T_loc - b0)
            block = positions[f_lo + b0:f_lo + b0 + nb]
            if isinstance(block, torch.Tensor):      # already resident (device arm of bench.py)
                pos_d = block.to(device=self.device, dtype=torch.float64).contiguous()
            else:
                pos = np.ascontiguousarray(block, dtype=np.float64)
                pos_d = torch.from_numpy(pos).to(self.device, non_blocking=True)
            with timer.phase("potential"):
                t = engine.build_transmission(plan, pos_d, out=tbuf[:nb])
            with timer.phase("propagate"):
                for p0 in range(0, P, pb):
                    np_ = min(pb, P - p0)
                    engine.propagate(plan, self._probes[p0:p0 + np_], t, wf_out=store, frame0=b0, probe0=p0,
                                     layer_every=self.layer_every, work=work)
        # (L, P, T, nx, ny) storage exposed in the reference's (P, T, nx, ny, L) index order
        self.wavefunction_data = store.permute(1, 2, 3, 4, 0)
        logger.info(f"Simulation completed in {time.time() - t_start:.2f}s ({T_loc} frames computed)")

        # axis labels exactly as the reference builds them (calculators.py:218-221): float32, from
        # `sampling` rather than the true pixel size
        kxs = torch.fft.fftshift(torch.fft.fftfreq(nx, self.sampling))
        kys = torch.fft.fftshift(torch.fft.fftfreq(ny, self.sampling))
        time_array = np.arange(self.n_frames) * self.trajectory.timestep
        if self.layer_every > 0:
            taps = [z for z in range(self.nz - 1) if (z + 1) % self.layer_every == 0] + [self.nz - 1]
            layer_array = np.array(taps)
        else:
            layer_array = np.array([0])
        wf = WFData(probe_positions=self.probe_positions, time=time_array, kxs=kxs, kys=kys, layer=layer_array,
                    wavefunction_data=self.wavefunction_data, probe=self.base_probe)
        wf.shard = self.shard
        return wf
