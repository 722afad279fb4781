"""`Probe`, `probe_grid`, `create_batched_probes`, `Propagate` with the reference's signatures
(reference src/multislice/multislice.py), executed by the CUDA engine in complex64.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from .. import engine, hostmath
from ..hostmath import C_LIGHT as c_light, H_PLANCK as h_planck, M_ELECTRON as m_electron, Q_ELECTRON as q_electron  # noqa: F401
from ..hostmath import wavelength  # noqa: F401

logger = logging.getLogger(__name__)
complex_dtype = torch.complex64
float_dtype = torch.float32


def m_effective(eV):
    """Relativistic electron mass in kg (reference multislice.py:37-39)."""
    return m_electron + eV * q_electron / c_light ** 2


def _np(a):
    return np.asarray(a.cpu() if hasattr(a, "cpu") else a, dtype=np.float64)


class Probe:
    """Electron probe on the (xs, ys) grid (reference multislice.py:44-190).

    `mrad == 0` gives a plane wave (ones); otherwise `ifftshift(ifft2(|k| < mrad*1e-3/lambda))`,
    unnormalised and with the strict `<` of the reference.  `array` is a complex64 CUDA tensor.
    """

    def __init__(self, xs, ys, mrad, eV, array=None, device=None):
        self.device = engine._device(device)
        self.use_torch = True
        self.dtype = torch.float32
        self.complex_dtype = torch.complex64
        self.xs = xs
        self.ys = ys
        self.mrad = mrad
        self.eV = eV
        self.wavelength = hostmath.wavelength(eV)
        kxs, kys = hostmath.kgrid(_np(xs), _np(ys))
        self._kxs, self._kys = kxs, kys
        self.kxs = torch.from_numpy(kxs)
        self.kys = torch.from_numpy(kys)
        nx, ny = len(kxs), len(kys)
        if array is not None:
            if not hasattr(array, "to"):
                array = torch.from_numpy(np.ascontiguousarray(np.asarray(array)))
            self.array = array.to(device=self.device, dtype=torch.complex64).contiguous()
            return
        if mrad == 0:
            self.array = torch.ones((nx, ny), dtype=torch.complex64, device=self.device)
            return
        mask = hostmath.aperture_mask(kxs, kys, mrad, self.wavelength)
        recip = torch.from_numpy(mask.astype(np.complex64)).to(self.device)
        real = engine.fft2(recip, inverse=True, scale=1.0 / (nx * ny))
        self.array = torch.fft.ifftshift(real).contiguous()     # index permutation only

    def copy(self):
        return Probe(self.xs, self.ys, self.mrad, self.eV, array=self.array.clone(), device=self.device)

    def to_cpu(self):
        return self.array.cpu().numpy()

    def to_device(self, device):
        self.device = engine._device(device)
        self.array = self.array.to(self.device)
        return self

    def plot(self):
        import matplotlib.pyplot as plt
        fig, ax = plt.subplots()
        img = (torch.absolute(self.array.T) ** .25).cpu()
        ax.imshow(img, cmap="inferno", extent=(np.amin(_np(self.xs)), np.amax(_np(self.xs)),
                                                np.amin(_np(self.ys)), np.amax(_np(self.ys))))
        plt.show()

    def defocus(self, dz):
        """Fresnel-propagate the probe by dz Angstrom (positive: waist above the sample),
        reference multislice.py:183-190: multiply (dz>0) or divide (dz<0) the spectrum by
        exp(-i pi lambda dz k^2)."""
        if dz == 0:
            return
        px, py = hostmath.propagator_tables(self._kxs, self._kys, self.wavelength, dz)
        if dz < 0:
            px, py = 1.0 / px, 1.0 / py
        nx, ny = len(px), len(py)
        a = self.array.reshape(-1, nx, ny).contiguous()
        spec = engine.fft2(a)
        ramp_x = engine._c64(np.broadcast_to(px, (a.shape[0], nx)), self.device)
        ramp_y = engine._c64(np.broadcast_to(py, (a.shape[0], ny)), self.device)
        out = torch.empty_like(a)
        for i in range(a.shape[0]):
            out[i] = engine.shift_probes(spec[i], ramp_x[i:i + 1], ramp_y[i:i + 1])[0]
        self.array = out.reshape(self.array.shape)


def probe_grid(xlims, ylims, n, m):
    """(n*m, 2) scan positions (reference multislice.py:193-195)."""
    x, y = np.meshgrid(np.linspace(*xlims, n), np.linspace(*ylims, m))
    return np.reshape([x, y], (2, len(x.flat))).T


def create_batched_probes(base_probe, probe_positions, device=None):
    """Probe with array (P, nx, ny): ifft2(fft2(base) * exp(+2 pi i kx px) * exp(+2 pi i ky py))
    per position (reference multislice.py:198-235, sign convention included)."""
    base = base_probe.array
    if base.dim() == 3:
        base = base[0]
    base_k = engine.fft2(base.contiguous())
    rx, ry = hostmath.shift_ramps(base_probe._kxs, base_probe._kys, probe_positions)
    arr = engine.shift_probes(base_k, engine._c64(rx, base.device), engine._c64(ry, base.device))
    return Probe(base_probe.xs, base_probe.ys, base_probe.mrad, base_probe.eV, array=arr, device=base_probe.device)


def _plan_for_potential(probe, potential):
    cached = getattr(potential, "_plan", None)
    xs, ys, zs = _np(potential.xs), _np(potential.ys), _np(potential.zs)
    if cached is not None and cached.eV == probe.eV:
        return cached
    plan = engine.make_plan(xs, ys, zs, [1], eV=probe.eV, device=probe.device)
    if cached is None and hasattr(potential, "kxs"):      # user object: honour its k axes (reference :273)
        kxs, kys = _np(potential.kxs), _np(potential.kys)
        px, py = hostmath.propagator_tables(kxs, kys, plan.wavelength, plan.dz)
        plan.prop_x = engine._c64(px / (plan.nx * plan.ny), plan.device)
        plan.prop_y = engine._c64(py, plan.device)
    return plan


def Propagate(probe, potential, device=None):
    """Multislice propagation of one probe or a (P, nx, ny) batch through `potential`
    (reference multislice.py:237-299): every slice psi *= exp(i sigma V_z); between slices
    psi = ifft2(P * fft2(psi)).  Returns the real-space exit wave(s) as a complex64 CUDA tensor,
    squeezed to (nx, ny) for a single probe like the reference."""
    if probe.array.dim() == 2:
        probe.array = probe.array[None, :, :]
    plan = _plan_for_potential(probe, potential)
    V = getattr(potential, "_V", None)
    if V is None:
        arr = potential.array
        if not hasattr(arr, "to"):
            arr = torch.from_numpy(np.asarray(arr))
        V = arr.to(device=plan.device, dtype=torch.float32).permute(2, 0, 1)
    V = V.contiguous()
    t = engine.transmission_from_potential(V, plan.sigma)[None]          # (1, nz, nx, ny)
    out = engine.propagate(plan, probe.array.contiguous(), t)[0]
    if out.shape[0] == 1:
        return out[0]
    return out
