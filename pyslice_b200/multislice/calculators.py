"""`MultisliceCalculator` with the reference's setup()/run() API
(reference src/multislice/calculators.py:39-250), driving the CUDA engine.

Differences from the reference that a caller can observe (all documented in DESIGN.md):
  * results are complex64 CUDA tensors instead of complex128 CPU tensors;
  * frames are processed in batches on the GPU; with `torch.distributed` initialised (one process
    per GPU) each rank propagates a contiguous block of frames and `WFData.shard` records it;
  * the `psi_data/torch_<key>/frame_<i>.npy` frame cache (reference calculators.py:81-92,140,259-260,311) is
    opt-in (`frame_cache=`) instead of a side effect of every run, and its default key also covers the atom
    positions (the reference's key ignores them, so a new trajectory of the same shape silently reads the old
    frames); `cache_key="reference"` reproduces the reference's key to read caches it wrote;
  * new optional kwarg `layer_every`: also record the wave function after every n-th slice
    (layer axis of WFData); the default 0 reproduces the reference's single layer.
"""
from __future__ import annotations

import hashlib
import logging
import queue
import threading
import time
from dataclasses import dataclass
from pathlib import Path
from typing import List, Optional, Tuple

import numpy as np
import torch

from .. import engine, hostmath
from ..postprocessing.wf_data import SlabStore, WFData
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


_COPY_STREAMS = {}


def _copy_stream(device) -> "torch.cuda.Stream":
    """one side stream per device for host -> device copies that overlap the kernels of the main stream"""
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COPY_STREAMS[key]


class _FrameWriter:
    """Background writer of the per-frame cache files: (P, nx, ny) complex64 host tensors in, `.npy` files of shape
    (P, nx, ny, 1, 1) complex128 out (the reference's wire format)."""

    def __init__(self):
        self.q = queue.Queue(maxsize=64)
        self.err = None
        self.th = threading.Thread(target=self._run, daemon=True)
        self.th.start()

    def _run(self):
        while True:
            item = self.q.get()
            if item is None:
                return
            path, frame = item
            try:
                np.save(path, frame.numpy().astype(np.complex128)[:, :, :, None, None])
            except Exception as e:          # surfaced by close()
                self.err = e

    def put(self, path, frame):
        self.q.put((path, frame))

    def close(self):
        self.q.put(None)
        self.th.join()
        if self.err is not None:
            raise self.err


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
        frame_cache=False,
        cache_key: str = "positions",
        adf_collection_angle: Optional[float] = None,
    ):
        """Same keyword arguments and defaults as the reference (calculators.py:96-109).
        `defocus`, `batch_size`, `save_path`, `cleanup_temp_files` are accepted and, as in the
        reference, have no effect on the result.

        frame_cache: False (default) / True (`./psi_data`, the reference's location) / a directory.  When on, every
        frame's exit waves are written to `<dir>/torch_<key>/frame_<i>.npy` in the reference's wire format
        ((P, nx, ny, 1, 1) complex128, calculators.py:276,311) by a writer thread, and frames whose file exists
        are loaded instead of computed (calculators.py:259-260).  cache_key: "positions" (default: the reference's
        parameters plus a digest of the positions) or "reference" (exactly calculators.py:81-92).

        adf_collection_angle (mrad): imaging runs that only want `HAADFData(wf).calculateADF(angle)`.  run() then keeps
        the detector sums sum_k |psi_k| * (|k| > angle/lambda) per (layer, probe, frame) -- the reduction of reference
        haadf_data.py:43-65, taken right after each exit FFT -- and never allocates the (P, T, nx, ny) exit-wave cube
        (54 GB at config C3); `wavefunction_data` is None on such a WFData."""
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
        self.adf_collection_angle = adf_collection_angle
        if adf_collection_angle is not None and frame_cache:
            raise ValueError("adf_collection_angle keeps no exit waves: not available with the frame cache")
        if cache_key not in ("positions", "reference"):
            raise ValueError("cache_key must be 'positions' or 'reference'")
        if frame_cache and self.layer_every > 0:
            raise ValueError("the frame cache holds exit waves only (reference format): not available with layer_every")
        self.output_dir = None
        if frame_cache:
            key = self._generate_cache_key(trajectory, aperture, voltage_eV, slice_thickness, sampling, probe_positions,
                                           with_positions=(cache_key == "positions"))
            root = Path("psi_data") if frame_cache is True else Path(frame_cache)
            self.output_dir = root / f"torch_{key}"
            self.output_dir.mkdir(parents=True, exist_ok=True)

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
    def _generate_cache_key(self, trajectory, aperture, voltage_eV, slice_thickness, sampling, probe_positions,
                            with_positions: bool = False) -> str:
        """The reference's key (calculators.py:78-92: md5 of the sorted parameter dict, 12 hex digits; 'backend'
        is 'pytorch' as in a reference run with torch installed); `with_positions` appends a digest of the positions."""
        params = {
            'n_frames': trajectory.n_frames,
            'n_atoms': trajectory.n_atoms,
            'box_matrix': trajectory.box_matrix.tolist(),
            'atom_types': trajectory.atom_types.tolist(),
            'aperture': aperture,
            'voltage_eV': voltage_eV,
            'slice_thickness': slice_thickness,
            'sampling': sampling,
            'probe_positions': probe_positions,
            'backend': 'pytorch',
        }
        if with_positions:
            pos = trajectory.positions
            pos = pos.detach().cpu().numpy() if isinstance(pos, torch.Tensor) else np.asarray(pos)
            params['positions_sha1'] = hashlib.sha1(np.ascontiguousarray(pos, dtype=np.float64).tobytes()).hexdigest()
        return hashlib.md5(str(sorted(params.items())).encode()).hexdigest()[:12]

    def release_workspace(self) -> None:
        """drop the per-batch device buffers kept between run() calls (the transmission stack of one frame batch, psi work
        area, binning scratch) of this calculator's device"""
        engine.WORKSPACES.pop(str(self.device), None)

    def _cache_file(self, frame: int) -> Path:
        return self.output_dir / f"frame_{frame}.npy"

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
        fb, pb = engine.batch_sizes(plan, P, max(T_loc, 1))
        # one probe per frame: the stack is written once and read once, so it is kept as float32 phases (half the HBM
        # traffic, exp(i*phase) evaluated inside the fused slice step); shared by several probes it stays complex64
        use_phase = P == 1 and engine.phase_format_supported(plan)
        # Per-batch workspaces are reused by every batch, every later run() and the next calculator of the same geometry on
        # this device (one set per device is kept, see engine.WORKSPACES): besides saving the allocations, stable addresses
        # let libpsb replay a batch's ~2000 launches as recorded CUDA graphs (graph_cache.cu).  release_workspace() frees them.
        A = int(self.trajectory.positions.shape[1])
        ws_key = (str(self.device), fb, pb, P, nx, ny, plan.nz, use_phase, A)
        ws = engine.WORKSPACES.get(str(self.device))
        if ws is None or ws["key"] != ws_key:
            engine.WORKSPACES.pop(str(self.device), None)
            ws = None
            ws = dict(key=ws_key,
                      tbuf=torch.empty((fb, plan.nz, nx, ny), dtype=torch.float32 if use_phase else torch.complex64, device=self.device),
                      work=torch.empty((fb * min(pb, P), nx, ny), dtype=torch.complex64, device=self.device),
                      t0=torch.empty((fb, nx, ny), dtype=torch.complex64, device=self.device) if use_phase else None,
                      scratch=torch.empty((nx * ny * engine.chunk_images(plan, fb),), dtype=torch.complex64, device=self.device),
                      bins=engine.bin_buffers(plan, fb, A))
            engine.WORKSPACES[str(self.device)] = ws
        tbuf, work = ws["tbuf"], ws["work"]
        # the result store is allocated AFTER the multi-GB stack: while a previous result is still alive the caching
        # allocator would otherwise carve the new store out of the cached stack block and then cudaMalloc a fresh stack
        # (tens of milliseconds of idle GPU, seen as random gaps between the phases of repeated runs)
        # Result store.  One process: (L, P, T, nx, ny), exposed as the reference's (P, T, nx, ny, L).  Frame-sharded run:
        # the per-destination slab layout of the frames -> kx-rows all-to-all (SlabStore), written by the exit FFT itself.
        # Detector-only run: (L, P, T) float64 sums and one batch of k-space scratch.
        world = self.shard.world if self.shard is not None else 1
        slabs = world > 1 and self.output_dir is None and self.adf_collection_angle is None
        det = None
        if self.adf_collection_angle is not None:
            kxs32 = torch.fft.fftshift(torch.fft.fftfreq(nx, self.sampling))
            kys32 = torch.fft.fftshift(torch.fft.fftfreq(ny, self.sampling))
            q = torch.sqrt(kxs32[:, None] ** 2 + kys32[None, :] ** 2)
            radius = (self.adf_collection_angle * 1e-3) / self.base_probe.wavelength
            mask = (q > radius).to(torch.float32).to(self.device).contiguous()       # haadf_data.py:52-57
            sums = torch.zeros((self.n_layers, P, max(T_loc, 1)), dtype=torch.float64, device=self.device)
            det = (mask, sums, torch.empty((fb * min(pb, P), nx, ny), dtype=torch.complex64, device=self.device))
            store = None
        elif slabs:
            store = torch.empty((self.n_layers * P * T_loc * nx * ny,), dtype=torch.complex64, device=self.device)
        else:
            store = torch.empty((self.n_layers, P, T_loc, nx, ny), dtype=torch.complex64, device=self.device)
        positions = self.trajectory.positions
        # frame cache (opt-in): cached frames are loaded, the others are computed in contiguous runs and handed to
        # a writer thread (D2H on a side stream would buy nothing here: the files are written by the host anyway)
        cached = [False] * T_loc
        writer = None
        if self.output_dir is not None:
            cached = [self._cache_file(f_lo + i).exists() for i in range(T_loc)]
            for i in range(T_loc):
                if cached[i]:
                    data = np.load(self._cache_file(f_lo + i))            # (P, nx, ny, 1, 1), reference wire format
                    if data.shape != (P, nx, ny, 1, 1):
                        raise ValueError(f"{self._cache_file(f_lo + i)}: shape {data.shape}, expected {(P, nx, ny, 1, 1)}")
                    store[0, :, i] = torch.from_numpy(np.ascontiguousarray(data[:, :, :, 0, 0])).to(self.device, torch.complex64)
            writer = _FrameWriter()
        self.frames_cached = sum(cached)
        self.frames_computed = T_loc - self.frames_cached
        runs, i = [], 0
        while i < T_loc:                                  # contiguous runs of frames to compute
            if cached[i]:
                i += 1
                continue
            j = i
            while j < T_loc and not cached[j]:
                j += 1
            runs.append((i, j))
            i = j
        batches = [(b0, min(fb, r1 - b0)) for r0, r1 in runs for b0 in range(r0, r1, fb)]
        # Host positions go up on a copy stream, every batch queued before the first kernel: only the first batch's
        # copy (a few MB) is exposed, the rest overlap the compute of earlier batches (pinned host memory; pageable
        # memory degrades to a staged copy but stays correct)
        uploads = {}
        if batches and not isinstance(positions, torch.Tensor) and self.device.type == "cuda":
            main = torch.cuda.current_stream(self.device)
            side = _copy_stream(self.device)
            dev_pos = torch.empty((T_loc,) + tuple(positions.shape[1:]), dtype=torch.float64, device=self.device)
            ready = torch.cuda.Event()
            ready.record(main)                              # the buffer's previous life on the main stream is over
            side.wait_event(ready)
            with torch.cuda.stream(side):
                for b0, nb in batches:
                    pos = np.ascontiguousarray(positions[f_lo + b0:f_lo + b0 + nb], dtype=np.float64)
                    dev_pos[b0:b0 + nb].copy_(torch.from_numpy(pos), non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    uploads[b0] = (dev_pos[b0:b0 + nb], ev)
        for b0, nb in batches:
            block = positions[f_lo + b0:f_lo + b0 + nb]
            if b0 in uploads:
                pos_d, ev = uploads.pop(b0)
                torch.cuda.current_stream(self.device).wait_event(ev)
            elif isinstance(block, torch.Tensor):      # already resident (device arm of bench.py)
                pos_d = block.to(device=self.device, dtype=torch.float64).contiguous()
            else:
                pos = np.ascontiguousarray(block, dtype=np.float64)
                pos_d = torch.from_numpy(pos).to(self.device, non_blocking=True)
            with timer.phase("potential"):
                t = engine.build_transmission(plan, pos_d, out=tbuf[:nb], phase=use_phase, scratch=ws["scratch"], bins=ws["bins"])
            with timer.phase("propagate"):
                for p0 in range(0, P, pb):
                    np_ = min(pb, P - p0)
                    engine.propagate(plan, self._probes[p0:p0 + np_], t, wf_out=store, frame0=b0, probe0=p0,
                                     layer_every=self.layer_every, work=work, detector=det, t0=ws["t0"],
                                     slabs=(world, self.n_layers, T_loc, P) if slabs else None)
            if writer is not None:
                host = store[0, :, b0:b0 + nb].to("cpu")                 # (P, nb, nx, ny); synchronises this batch
                for k in range(nb):
                    writer.put(self._cache_file(f_lo + b0 + k), host[:, k])
        if writer is not None:
            writer.close()
            if self.cleanup_temp_files:                   # reference calculators.py:235-245
                for i in range(T_loc):
                    self._cache_file(f_lo + i).unlink(missing_ok=True)
                try:
                    self.output_dir.rmdir()
                except OSError:
                    pass
        if det is not None:
            self.wavefunction_data = None
        elif slabs:
            self.wavefunction_data = SlabStore(store, world, self.n_layers, T_loc, P, nx, ny)
        else:
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
        if det is not None:
            wf.adf_sums = det[1][:, :, :T_loc]          # (L, P, T_local) float64 on the device
            wf.adf_collection_angle = self.adf_collection_angle
        return wf
