"""`WFData`: the output container (reference src/postprocessing/wf_data.py:9-28)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, List, Tuple

import numpy as np


class SlabStore:
    """Exit waves of a frame-sharded (multi-GPU) run, stored the way the frames -> kx-rows all-to-all wants them
    (SURVEY.md 8e): `buf` is flat, laid out [destination h][layer][local frame][probe][kx row of h's block][ky], written
    directly by the exit FFT (psb_propagate_ex, slab layout), so every (destination, layer) block is one contiguous message
    and no pack pass runs before the exchange.  `shape` is the logical (P, T_local, nx, ny, L) of the reference's
    wavefunction_data; indexing (`store[...]`, `store[:, :, :, :, l]`) materialises a dense copy for code that wants one."""

    def __init__(self, buf, world, n_layers, n_frames, n_probes, nx, ny):
        self.buf, self.world = buf, world
        self.L, self.T, self.P, self.nx, self.ny = n_layers, n_frames, n_probes, nx, ny
        counts = [nx // world + (1 if r < nx % world else 0) for r in range(world)]
        self.row_counts = counts
        self.row_starts = [sum(counts[:r]) for r in range(world)]
        self.shape = (n_probes, n_frames, nx, ny, n_layers)
        self.dtype, self.device = buf.dtype, buf.device

    def block(self, h, layer=None):
        """(L, T, P, rows_h, ny) view of destination h's slab, or (T, P, rows_h, ny) of one layer"""
        planes = self.L * self.T * self.P
        o = planes * self.row_starts[h] * self.ny
        v = self.buf[o:o + planes * self.row_counts[h] * self.ny].view(self.L, self.T, self.P, self.row_counts[h], self.ny)
        return v if layer is None else v[layer]

    def layer(self, layer):
        """dense (P, T, nx, ny) copy of one layer"""
        import torch
        return torch.cat([self.block(h, layer).permute(1, 0, 2, 3) for h in range(self.world)], dim=2)

    def dense(self):
        """dense (P, T, nx, ny, L) copy (the reference's index order)"""
        import torch
        return torch.stack([self.layer(l) for l in range(self.L)], dim=-1)

    def __getitem__(self, idx):
        if isinstance(idx, tuple) and len(idx) == 5 and isinstance(idx[4], int) and all(i == slice(None) for i in idx[:4]):
            return self.layer(idx[4] % self.L)
        return self.dense()[idx]

    def cpu(self):
        return self.dense().cpu()


@dataclass
class WFData:
    """probe_positions, time (ps), kxs, kys, layer, wavefunction_data (probe, time, kx, ky, layer), probe.

    `wavefunction_data[p, f, :, :, l]` is the fftshifted, unnormalised forward FFT of the wave
    function of probe p / frame f after layer l (reciprocal space), a complex64 CUDA tensor here.

    Multi-GPU runs add `shard` (calculators.FrameShard): `wavefunction_data` then holds only this rank's block of frames
    (`shard.counts[shard.rank]` of them, starting at `shard.start`) while `time` describes the whole run, and it is a
    `SlabStore` (same logical shape, laid out for the all-to-all; index it or call `.dense()` for a tensor).  TACAWData and
    HAADFData know about both (their reducers are collectives then: call them on every rank); code that reads
    `wavefunction_data` directly must index frames relative to `shard.start`.
    """
    probe_positions: List[Tuple[float, float]]
    time: np.ndarray
    kxs: Any
    kys: Any
    layer: np.ndarray
    wavefunction_data: Any
    probe: Any
