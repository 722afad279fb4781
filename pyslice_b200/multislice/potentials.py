"""`gridFromTrajectory` and `Potential` with the reference's signatures
(reference src/multislice/potentials.py:113-131, 187-348), computed by the CUDA engine.

`Potential.array` has the reference's (nx, ny, nz) shape (a permuted view of the engine's
(nz, nx, ny) float32 buffer, so `array[:, :, z]` is a contiguous plane here).
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import engine, hostmath
from ..hostmath import atomic_number as getZfromElementName  # noqa: F401  (reference name)

logger = logging.getLogger(__name__)


def gridFromTrajectory(trajectory, sampling=0.1, slice_thickness=0.5):
    """xs, ys, zs, lx, ly, lz from the box diagonal (reference potentials.py:113-131)."""
    return hostmath.grid_from_box(trajectory.box_matrix, sampling, slice_thickness)


def kirkland(qsq, Z):
    """Kirkland form factor f_e(q^2) on a (nx, ny) grid of q^2 values (reference potentials.py:50-96).
    Host float64 table helper; returns a NumPy array."""
    qsq = np.asarray(qsq.cpu() if hasattr(qsq, "cpu") else qsq, dtype=np.float64)
    abcd = hostmath.kirkland_table()[hostmath.atomic_number(Z) - 1]
    out = np.zeros_like(qsq)
    for a, b, c, d in abcd:
        out += a / (qsq + b) + c * np.exp(-d * qsq)
    return out


class Potential:
    """Projected potential slices of one atomic configuration.

    Same constructor as the reference (potentials.py:188).  Atoms are binned into slices with the
    reference's float64 interval rule (bit-exact), the structure-factor sum and per-slice inverse
    FFT run on the GPU in float32.  Only `slice_axis=2` is supported (the reference's other values
    reuse z-built grids inconsistently).
    """

    def __init__(self, xs, ys, zs, positions, atomTypes, kind="kirkland", device=None, slice_axis=2):
        if slice_axis != 2:
            raise NotImplementedError("pyslice_b200 supports slice_axis=2 only")
        if kind not in ("kirkland", "gauss"):
            raise ValueError(f"unknown potential kind {kind!r}")
        xs = np.asarray(xs.cpu() if hasattr(xs, "cpu") else xs, dtype=np.float64)
        ys = np.asarray(ys.cpu() if hasattr(ys, "cpu") else ys, dtype=np.float64)
        zs = np.asarray(zs.cpu() if hasattr(zs, "cpu") else zs, dtype=np.float64)
        self._plan = engine.make_plan(xs, ys, zs, list(atomTypes), eV=100e3, device=device)
        plan = self._plan
        if kind == "gauss":   # reference potentials.py:279-280: exp(-q^2/2) for every type
            qsq = plan.kxs[:, None] ** 2 + plan.kys[None, :] ** 2
            g = np.broadcast_to(np.exp(-qsq / 2), (plan.ntypes,) + qsq.shape).astype(np.float32)
            plan.formfactors = torch.from_numpy(np.ascontiguousarray(g)).to(plan.device)
        self.device = plan.device
        self.use_torch = True
        self.dtype = torch.float32
        self.complex_dtype = torch.complex64
        self.xs = torch.from_numpy(xs)
        self.ys = torch.from_numpy(ys)
        self.zs = torch.from_numpy(zs)
        self.kxs = torch.from_numpy(plan.kxs)
        self.kys = torch.from_numpy(plan.kys)
        self.slice_axis = slice_axis
        self.inplane_axis1, self.inplane_axis2 = 0, 1
        self.slice_coords = zs
        self.slice_spacing = plan.dz
        self.n_slices = plan.nz
        pos = np.ascontiguousarray(np.asarray(positions.cpu() if hasattr(positions, "cpu") else positions,
                                              dtype=np.float64)).reshape(1, -1, 3)
        pos_d = torch.from_numpy(pos).to(plan.device)
        _, V = engine.build_transmission(plan, pos_d, want_potential=True)
        self._V = V[0]                                  # (nz, nx, ny) float32
        self.array = self._V.permute(1, 2, 0)           # reference layout (nx, ny, nz)

    def to_cpu(self):
        return self.array.cpu().numpy()

    def to_device(self, device):
        self.device = device
        return self

    def plot(self):
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots()
        img = torch.sum(torch.absolute(self.array), axis=2).T.cpu()
        ax.imshow(img, cmap="inferno", extent=(float(self.xs.min()), float(self.xs.max()),
                                                float(self.ys.min()), float(self.ys.max())))
        plt.show()
