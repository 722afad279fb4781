"""`TACAWData(wfdata)`: time -> frequency FFT of the exit waves and the reducers over the
|Psi(omega, k)|^2 cube (reference src/postprocessing/tacaw_data.py:38-353), on the GPU.

Multi-GPU: when the WFData came from a frame-sharded run (one process per GPU), the exit waves
are re-sharded once with an NCCL all-to-all from frames to kx rows; each rank then holds the
intensity of its kx rows and the reducers combine ranks with a small all-reduce.
"""
from __future__ import annotations

import logging
from typing import List, Optional

import numpy as np
import torch

from .. import engine
from .wf_data import WFData

logger = logging.getLogger(__name__)


def _row_split(n: int, world: int):
    starts, counts = engine.row_split(n, world)
    return starts, counts


def _exchange(send, recv, rank, group=None):
    """one all-to-all of contiguous blocks: send[h] -> rank h, recv[g] <- rank g (views, nothing is packed)"""
    import torch.distributed as dist
    send_r = [torch.view_as_real(x) for x in send]
    recv_r = [torch.view_as_real(x) for x in recv]
    if dist.get_backend(group) == "nccl":
        dist.all_to_all(recv_r, send_r, group=group)          # grouped send/recv over NVLink 5 / NVSwitch
        return
    # gloo (CPU tests of the host logic) has no all_to_all: pairwise exchange
    reqs = []
    for h in range(len(send)):
        if h == rank:
            recv_r[h].copy_(send_r[h])
            continue
        reqs.append(dist.isend(send_r[h], h, group=group))
        reqs.append(dist.irecv(recv_r[h], h, group=group))
    for r in reqs:
        r.wait()


def slabs_to_rows_all_to_all(store, layer: int, shard, group=None) -> torch.Tensor:
    """SlabStore (frame shard, slab layout) -> (P, T_total, nx_local, ny) kx-row shard of one layer.

    The one collective of the pipeline: rank g sends rank h the slab psi[:, F_g, K_h, :], which the exit
    FFT already wrote as one contiguous block per (destination, layer).  With a single layer the send side is the whole
    store and the exchange is one `all_to_all_single`; with several layers the blocks of the requested layer are sent as
    views.  The result is a (P, T, rows, ny) view of a (T, P, rows, ny) buffer (every source's frames contiguous)."""
    import torch.distributed as dist
    world, rank = shard.world, shard.rank
    P, ny = store.P, store.ny
    rows = store.row_counts[rank]
    full = torch.empty((shard.total, P, rows, ny), dtype=store.dtype, device=store.device)
    if store.L == 1 and dist.get_backend(group) == "nccl":
        in_split = [2 * store.T * P * store.row_counts[h] * ny for h in range(world)]
        out_split = [2 * shard.counts[g] * P * rows * ny for g in range(world)]
        dist.all_to_all_single(torch.view_as_real(full).reshape(-1), torch.view_as_real(store.buf).reshape(-1),
                               output_split_sizes=out_split, input_split_sizes=in_split, group=group)
        return full.permute(1, 0, 2, 3)
    send = [store.block(h, layer) for h in range(world)]                       # (T_loc, P, rows_h, ny) contiguous views
    recv, t0 = [], 0
    for g in range(world):
        recv.append(full[t0:t0 + shard.counts[g]])
        t0 += shard.counts[g]
    _exchange(send, recv, rank, group)
    return full.permute(1, 0, 2, 3)


def frames_to_rows_all_to_all(wf_layer: torch.Tensor, shard, group=None) -> torch.Tensor:
    """(P, T_local, nx, ny) dense frame shard -> (P, T_total, nx_local, ny) kx-row shard (packs one block per destination;
    used when the exit waves are not in the slab layout: frame-cache runs, WFData assembled by hand)."""
    P, T_loc, nx, ny = wf_layer.shape
    world, rank = shard.world, shard.rank
    starts, counts = _row_split(nx, world)
    send = [wf_layer[:, :, starts[h]:starts[h] + counts[h], :].permute(1, 0, 2, 3).contiguous() for h in range(world)]
    rows = counts[rank]
    full = torch.empty((shard.total, P, rows, ny), dtype=wf_layer.dtype, device=wf_layer.device)
    t0 = 0
    recv = []
    for g in range(world):
        recv.append(full[t0:t0 + shard.counts[g]])
        t0 += shard.counts[g]
    _exchange(send, recv, rank, group)
    return full.permute(1, 0, 2, 3)


class TACAWData(WFData):
    """Drop-in for the reference class: `TACAWData(wfdata, layer_index=None)` gains `.frequencies`
    (THz for ps time steps) and `.intensity` (probe, frequency, kx, ky) and shares the WFData's
    attribute dict exactly as the reference does (tacaw_data.py:38-43)."""

    def __init__(self, WFData, layer_index: int = None):
        self.__class__ = type(WFData.__class__.__name__, (self.__class__, WFData.__class__), {})
        self.__dict__ = WFData.__dict__
        self.fft_from_wf_data(layer_index)

    def fft_from_wf_data(self, layer_index: int = None):
        """intensity = |fftshift_t FFT_t(psi - mean_t psi)|^2 for one layer (tacaw_data.py:61-106)."""
        if layer_index is None:
            layer_index = len(self.layer) - 1
        if layer_index < 0 or layer_index >= len(self.layer):
            raise ValueError(f"layer_index {layer_index} out of range [0, {len(self.layer)-1}]")
        n_freq = len(self.time)
        dt = self.time[1] - self.time[0]
        self.frequencies = np.fft.fftshift(np.fft.fftfreq(n_freq, d=dt))

        from .wf_data import SlabStore
        shard = getattr(self, "shard", None)
        timer = getattr(self, "_timer", None) or engine.NO_TIMER      # bench.py: CUDA-event phases
        if isinstance(self.wavefunction_data, SlabStore):              # frame-sharded run: blocks are ready to send
            store = self.wavefunction_data
            self.row_range = (store.row_starts[shard.rank], store.row_starts[shard.rank] + store.row_counts[shard.rank])
            with timer.phase("all_to_all"):
                wf_layer = slabs_to_rows_all_to_all(store, layer_index, shard)
            with timer.phase("tacaw"):
                self.intensity = engine.tacaw_intensity(wf_layer)
            return
        wf_layer = self.wavefunction_data[:, :, :, :, layer_index]
        if not hasattr(wf_layer, "dim"):
            wf_layer = torch.from_numpy(np.asarray(wf_layer))
        dev = engine._device(wf_layer.device if wf_layer.device.type == "cuda" else None)
        wf_layer = wf_layer.to(device=dev, dtype=torch.complex64)
        if wf_layer.stride(3) != 1 or wf_layer.stride(2) != wf_layer.shape[3]:
            wf_layer = wf_layer.contiguous()
        self.row_range = (0, wf_layer.shape[2])
        if shard is not None and shard.world > 1:
            starts, counts = _row_split(wf_layer.shape[2], shard.world)
            self.row_range = (starts[shard.rank], starts[shard.rank] + counts[shard.rank])
            with timer.phase("all_to_all"):
                wf_layer = frames_to_rows_all_to_all(wf_layer, shard)
        with timer.phase("tacaw"):
            self.intensity = engine.tacaw_intensity(wf_layer)

    # ---- helpers ---------------------------------------------------------------------------
    def _n_probes(self):
        return len(self.probe_positions)

    def _check_probe(self, probe_index):
        if probe_index is not None and probe_index >= self._n_probes():
            raise ValueError(f"Probe index {probe_index} out of range")

    def _allreduce(self, x: torch.Tensor) -> torch.Tensor:
        shard = getattr(self, "shard", None)
        if shard is not None and shard.world > 1:
            import torch.distributed as dist
            dist.all_reduce(x)
        return x

    def _full_rows(self, local: torch.Tensor, nx: int) -> torch.Tensor:
        """place this rank's kx rows (dim -2) into a zero full-size array and sum over ranks"""
        shard = getattr(self, "shard", None)
        if shard is None or shard.world == 1:
            return local
        full = torch.zeros(local.shape[:-2] + (nx, local.shape[-1]), dtype=local.dtype, device=local.device)
        full[..., self.row_range[0]:self.row_range[1], :] = local
        return self._allreduce(full)

    def _sum_k(self, mask: Optional[torch.Tensor] = None) -> np.ndarray:
        """(P, T) float64: sum over (kx, ky) of intensity (* mask)"""
        P, T, rows, ny = self.intensity.shape
        s = engine.sum_pixels(self.intensity.reshape(P * T, rows * ny), mask)
        return self._allreduce(s).reshape(P, T).cpu().numpy()

    # ---- reducers (reference tacaw_data.py:109-353) -----------------------------------------
    def spectrum(self, probe_index: int = None) -> np.ndarray:
        """Sum over all k for one probe, or the mean over probes of those sums (:109-143)."""
        self._check_probe(probe_index)
        s = self._sum_k()
        return s.mean(axis=0) if probe_index is None else s[probe_index]

    def spectrum_image(self, frequency: float, probe_indices: Optional[List[int]] = None) -> np.ndarray:
        """k-summed intensity at the nearest frequency, one value per probe (:145-179)."""
        fi = int(np.argmin(np.abs(self.frequencies - frequency)))
        if probe_indices is None:
            probe_indices = list(range(self._n_probes()))
        return self._sum_k()[list(probe_indices), fi]

    def diffraction(self, probe_index: int = None) -> np.ndarray:
        """Sum over frequency -> (kx, ky), one probe or the mean over probes (:183-217)."""
        self._check_probe(probe_index)
        P, T, rows, ny = self.intensity.shape
        d = engine.sum_frames(self.intensity.reshape(P, T, rows * ny)).reshape(P, rows, ny)
        d = self._full_rows(d, len(self.kxs)).cpu().numpy().astype(np.float64)
        return d.mean(axis=0) if probe_index is None else d[probe_index]

    def spectral_diffraction(self, frequency: float, probe_index: int = None) -> np.ndarray:
        """(kx, ky) plane at the nearest frequency (:219-255)."""
        self._check_probe(probe_index)
        fi = int(np.argmin(np.abs(self.frequencies - frequency)))
        plane = self._full_rows(self.intensity[:, fi].contiguous(), len(self.kxs)).cpu().numpy().astype(np.float64)
        return plane.mean(axis=0) if probe_index is None else plane[probe_index]

    def masked_spectrum(self, mask: np.ndarray, probe_index: int = None) -> np.ndarray:
        """Spectrum with a (kx, ky) weight mask (:257-299).  The reference's shape check reads
        attributes that do not exist (`self.kx`); here it checks against kxs/kys."""
        mask = np.asarray(mask)
        if mask.shape != (len(self.kxs), len(self.kys)):
            raise ValueError(f"Mask shape {mask.shape} doesn't match k-space shape ({len(self.kxs)}, {len(self.kys)})")
        self._check_probe(probe_index)
        m = torch.from_numpy(np.ascontiguousarray(mask[self.row_range[0]:self.row_range[1]], dtype=np.float32))
        s = self._sum_k(m.to(self.intensity.device).reshape(-1))
        return s.mean(axis=0) if probe_index is None else s[probe_index]

    def dispersion(self, kx_path: np.ndarray, ky_path: np.ndarray, probe_index: int = None) -> np.ndarray:
        """Intensity at the grid points nearest to (kx_path, ky_path) -> (n_freq, n_k) (:301-353)."""
        self._check_probe(probe_index)
        kxs = np.asarray(self.kxs)
        kys = np.asarray(self.kys)
        ix = np.array([int(np.argmin(np.abs(kxs - v))) for v in kx_path])
        iy = np.array([int(np.argmin(np.abs(kys - v))) for v in ky_path])
        r0, r1 = self.row_range
        mine = (ix >= r0) & (ix < r1)
        P, T = self.intensity.shape[:2]
        out = torch.zeros((P, T, len(ix)), dtype=torch.float32, device=self.intensity.device)
        if mine.any():
            sel = torch.from_numpy(np.nonzero(mine)[0]).to(out.device)
            lx = torch.from_numpy(ix[mine] - r0).to(out.device)
            ly = torch.from_numpy(iy[mine]).to(out.device)
            out[:, :, sel] = self.intensity[:, :, lx, ly]
        d = self._allreduce(out).cpu().numpy().astype(np.float64)
        return d.mean(axis=0) if probe_index is None else d[probe_index]
