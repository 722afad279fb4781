"""CPU oracle for the PySlice multislice + TACAW hot path.  TEST INFRASTRUCTURE ONLY.

This module is a float64 / complex128 NumPy restatement of what the reference
(h-walk/PySlice, torch path, `/root/reference` in the build container) computes on the
path `Potential -> create_batched_probes -> Propagate -> exit FFT -> TACAWData`.
It exists to check the CUDA engine in `pyslice_b200/`; nothing in the product imports
it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may use it.

Parity pinning: the reference ships no usable golden vectors for this path (its
`.npy` goldens are absent from the checkout), so the oracle is
pinned by running the reference itself in the build container
(`tests/golden/make_golden.py`) and committing its outputs as `tests/golden/*.npz`;
`tests/test_oracle_golden.py` compares this module against those files.

Every function cites the reference lines it restates (paths relative to the reference
checkout).  FFTs go through `scipy.fft` (pocketfft, the same algorithm family torch's
CPU path uses) so that `workers=` can use all host threads for the CPU baseline.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.fft as sfft

# --- physical constants, src/multislice/multislice.py:31-34 -------------------------
M_ELECTRON = 9.109383e-31
Q_ELECTRON = 1.602177e-19
C_LIGHT = 299792458.0
H_PLANCK = 6.62607015e-34

_KIRKLAND = None
_HERE = os.path.dirname(os.path.abspath(__file__))
_TABLE_CSV = os.path.join(_HERE, "..", "pyslice_b200", "data", "kirkland_abcd.csv")

_ELEMENTS = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P",
             "S", "Cl", "Ar", "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn",
             "Ga", "Ge", "As", "Se", "Br", "Kr"]


def atomic_number(kind) -> int:
    """Element name or number -> Z (src/multislice/potentials.py:98-111, Z<=36 subset that
    src/multislice/calculators.py:67-75 can produce)."""
    if isinstance(kind, str):
        return _ELEMENTS.index(kind) + 1
    return int(kind)


def kirkland_table() -> np.ndarray:
    """(103,3,4) table, rows [a_i,b_i,c_i,d_i] -- the parse result of
    src/multislice/potentials.py:158-172 (stored pre-parsed in the package's CSV)."""
    global _KIRKLAND
    if _KIRKLAND is None:
        raw = np.loadtxt(_TABLE_CSV, delimiter=",", comments="#")
        _KIRKLAND = raw[:, 1:].reshape(103, 3, 4)
    return _KIRKLAND


def wavelength(eV: float) -> float:
    """Relativistic electron wavelength in Angstrom, src/multislice/multislice.py:41-42."""
    return H_PLANCK * C_LIGHT / ((eV * Q_ELECTRON) ** 2
                                 + 2 * eV * Q_ELECTRON * M_ELECTRON * C_LIGHT ** 2) ** 0.5 * 1e10


def interaction_sigma(eV: float) -> float:
    """Interaction parameter (Kirkland eq. 5.6), src/multislice/multislice.py:258-260."""
    e0 = M_ELECTRON * C_LIGHT ** 2 / Q_ELECTRON
    return (2 * np.pi) / (wavelength(eV) * eV) * (e0 + eV) / (2 * e0 + eV)


def grid_from_box(box_matrix, sampling=0.1, slice_thickness=0.5):
    """Grid rule, src/multislice/potentials.py:113-131 (orthogonal box, diagonal only)."""
    lx, ly, lz = box_matrix[0, 0], box_matrix[1, 1], box_matrix[2, 2]
    nx = int(lx / sampling) + 1
    ny = int(ly / sampling) + 1
    nz = int(lz / slice_thickness) + 1
    xs = np.linspace(0, lx, nx, endpoint=False)
    ys = np.linspace(0, ly, ny, endpoint=False)
    zs = np.linspace(0, lz, nz, endpoint=False)
    return xs, ys, zs, lx, ly, lz


def form_factor(qsq: np.ndarray, Z: int) -> np.ndarray:
    """Kirkland f_e(q^2) = sum a/(q^2+b) + sum c exp(-d q^2), src/multislice/potentials.py:50-96."""
    abcd = kirkland_table()[Z - 1]
    a, b, c, d = (abcd[:, i][:, None, None] for i in range(4))
    q = qsq[None, :, :]
    return np.sum(a / (q + b), axis=0) + np.sum(c * np.exp(-d * q), axis=0)


def slice_bounds(zs: np.ndarray):
    """Per-slice z intervals [lo, hi), src/multislice/potentials.py:230,304-305.

    lo[i] = zs[i]-dz/2 (0 for the first slice); hi[i] = zs[i]+dz/2 (zs[-1]+dz for the last).
    Evaluated with the identical float64 expressions, so adjacent slices can leave 1-ulp
    gaps or overlaps exactly as the reference does."""
    nz = len(zs)
    dz = zs[1] - zs[0] if nz > 1 else 0.5
    lo = np.empty(nz)
    hi = np.empty(nz)
    for i in range(nz):
        lo[i] = zs[i] - dz / 2 if i > 0 else 0
        hi[i] = zs[i] + dz / 2 if i < nz - 1 else zs[-1] + dz
    return lo, hi


def bin_atoms(z: np.ndarray, zs: np.ndarray) -> np.ndarray:
    """Membership matrix M[s, a] = (z[a] >= lo[s]) & (z[a] < hi[s]),
    src/multislice/potentials.py:307.  An atom can be in 0, 1 or 2 slices."""
    lo, hi = slice_bounds(zs)
    z = np.asarray(z, dtype=np.float64)
    return (z[None, :] >= lo[:, None]) & (z[None, :] < hi[:, None])


def potential(xs, ys, zs, positions, atom_kinds, workers=1) -> np.ndarray:
    """Projected potential slices V[x, y, s] (float64), src/multislice/potentials.py:226-342.

    V[:,:,s] = Re IFFT2( sum_types f_Z(k) * sum_{a in slice s} e^{-2 pi i kx x_a} e^{-2 pi i ky y_a} )
               / (dx^2 dy^2)   -- no physical prefactor, exactly as the reference."""
    nx, ny, nz = len(xs), len(ys), len(zs)
    dx = xs[1] - xs[0]
    dy = ys[1] - ys[0]
    kxs = np.fft.fftfreq(nx, d=dx)
    kys = np.fft.fftfreq(ny, d=dy)
    qsq = kxs[:, None] ** 2 + kys[None, :] ** 2
    pos = np.asarray(positions, dtype=np.float64)
    Zs = np.array([atomic_number(k) for k in atom_kinds])
    member = bin_atoms(pos[:, 2], zs)                       # (nz, A)
    recip = np.zeros((nz, nx, ny), dtype=np.complex128)
    for Z in sorted(set(Zs.tolist())):
        ff = form_factor(qsq, Z)
        of_type = Zs == Z
        for s in range(nz):
            sel = of_type & member[s]
            if not sel.any():
                continue
            ex = np.exp(-1j * 2 * np.pi * kxs[None, :] * pos[sel, 0][:, None])   # (A_s, nx)
            ey = np.exp(-1j * 2 * np.pi * kys[None, :] * pos[sel, 1][:, None])   # (A_s, ny)
            recip[s] += (ex.T @ ey) * ff                      # einsum('ax,ay->xy'), :328-330
    real = sfft.ifft2(recip, axes=(1, 2), workers=workers).real
    real /= dx ** 2 * dy ** 2
    return np.ascontiguousarray(np.moveaxis(real, 0, 2))     # (nx, ny, nz) like the reference


def probe_array(xs, ys, mrad, eV) -> np.ndarray:
    """Base probe, src/multislice/multislice.py:95-124: ones for a plane wave, else
    ifftshift(ifft2(|k| < alpha/lambda)) (unnormalised, strict '<')."""
    nx, ny = len(xs), len(ys)
    if mrad == 0:
        return np.ones((nx, ny), dtype=np.float64)
    dx = xs[1] - xs[0]
    dy = ys[1] - ys[0]
    kxs = np.fft.fftfreq(nx, d=dx)
    kys = np.fft.fftfreq(ny, d=dy)
    radius = (mrad * 1e-3) / wavelength(eV)
    radii = np.sqrt(kxs[:, None] ** 2 + kys[None, :] ** 2)
    recip = np.zeros((nx, ny))
    recip[radii < radius] = 1.0
    return np.fft.ifftshift(sfft.ifft2(recip))


def shifted_probes(base: np.ndarray, xs, ys, probe_positions) -> np.ndarray:
    """(P, nx, ny) probes, src/multislice/multislice.py:216-231:
    ifft2( fft2(base) * exp(+2 pi i kx px) * exp(+2 pi i ky py) )  (sign as in the reference)."""
    nx, ny = len(xs), len(ys)
    kxs = np.fft.fftfreq(nx, d=xs[1] - xs[0])
    kys = np.fft.fftfreq(ny, d=ys[1] - ys[0])
    base_k = sfft.fft2(np.asarray(base, dtype=np.complex128))
    out = np.empty((len(probe_positions), nx, ny), dtype=np.complex128)
    for i, (px, py) in enumerate(probe_positions):
        ramp_x = np.exp(2j * np.pi * kxs[:, None] * px)
        ramp_y = np.exp(2j * np.pi * kys[None, :] * py)
        out[i] = sfft.ifft2(base_k * ramp_x * ramp_y)
    return out


def fresnel_propagator(xs, ys, zs, eV) -> np.ndarray:
    """P = exp(-i pi lambda dz (kx^2+ky^2)), src/multislice/multislice.py:266-275."""
    nx, ny = len(xs), len(ys)
    kxs = np.fft.fftfreq(nx, d=xs[1] - xs[0])
    kys = np.fft.fftfreq(ny, d=ys[1] - ys[0])
    dz = zs[1] - zs[0] if len(zs) > 1 else 0.5
    ksq = kxs[:, None] ** 2 + kys[None, :] ** 2
    return np.exp(-1j * np.pi * wavelength(eV) * dz * ksq)


def propagate(psi: np.ndarray, V: np.ndarray, xs, ys, zs, eV, workers=1, n_slices=None) -> np.ndarray:
    """Multislice loop, src/multislice/multislice.py:278-294: for every slice psi *= exp(i sigma V_s);
    between slices psi = ifft2(P * fft2(psi)).  `n_slices` truncates the stack (the layer-resolved
    oracle of SURVEY.md section 8c: transmission of slice n_slices-1 applied, not propagated)."""
    psi = np.array(psi, dtype=np.complex128, ndmin=3)
    sigma = interaction_sigma(eV)
    P = fresnel_propagator(xs, ys, zs, eV)
    nz = V.shape[2] if n_slices is None else n_slices
    for s in range(nz):
        psi = np.exp(1j * sigma * V[:, :, s])[None] * psi
        if s < nz - 1:
            psi = sfft.ifft2(P[None] * sfft.fft2(psi, axes=(-2, -1), workers=workers),
                             axes=(-2, -1), workers=workers)
    return psi


def exit_to_kspace(psi: np.ndarray, workers=1) -> np.ndarray:
    """fftshift(fft2(psi)) over the last two axes, src/multislice/calculators.py:285-287."""
    return np.fft.fftshift(sfft.fft2(psi, axes=(-2, -1), workers=workers), axes=(-2, -1))


def frame_exit_waves(xs, ys, zs, positions, atom_kinds, probes, eV, workers=1, layer_slices=None):
    """One MD frame: potential -> propagate -> shifted k-space exit waves (P, nx, ny[, L]),
    the per-frame worker src/multislice/calculators.py:256-290."""
    V = potential(xs, ys, zs, positions, atom_kinds, workers=workers)
    if layer_slices is None:
        return exit_to_kspace(propagate(probes, V, xs, ys, zs, eV, workers=workers), workers=workers)
    outs = [exit_to_kspace(propagate(probes, V, xs, ys, zs, eV, workers=workers, n_slices=n), workers)
            for n in layer_slices]
    return np.stack(outs, axis=-1)


def multislice_run(positions, atom_kinds, box_matrix, aperture=0.0, voltage_eV=60e3,
                   slice_thickness=0.5, sampling=0.1, probe_positions=None, workers=1,
                   frame_threads=1):
    """MultisliceCalculator.setup()+run(), src/multislice/calculators.py:96-232.
    Returns (wavefunction_data (P,T,nx,ny,1) complex128, grid dict)."""
    xs, ys, zs, lx, ly, lz = grid_from_box(np.asarray(box_matrix), sampling, slice_thickness)
    if probe_positions is None:
        probe_positions = [(lx / 2, ly / 2)]
    base = probe_array(xs, ys, aperture, voltage_eV)
    probes = shifted_probes(base, xs, ys, probe_positions)
    T = positions.shape[0]
    out = np.zeros((len(probe_positions), T, len(xs), len(ys), 1), dtype=np.complex128)

    def one(f):
        out[:, f, :, :, 0] = frame_exit_waves(xs, ys, zs, positions[f], atom_kinds, probes,
                                              voltage_eV, workers=workers)

    if frame_threads > 1:
        with ThreadPoolExecutor(frame_threads) as pool:
            list(pool.map(one, range(T)))
    else:
        for f in range(T):
            one(f)
    grid = dict(xs=xs, ys=ys, zs=zs, lx=lx, ly=ly, lz=lz, base_probe=base, probes=probes)
    return out, grid


def wf_axes(nx, ny, sampling, n_frames, timestep):
    """kxs/kys labels (float32, from `sampling`, shifted) and time axis,
    src/multislice/calculators.py:218-221."""
    kxs = np.fft.fftshift(np.fft.fftfreq(nx, sampling)).astype(np.float32)
    kys = np.fft.fftshift(np.fft.fftfreq(ny, sampling)).astype(np.float32)
    return kxs, kys, np.arange(n_frames) * timestep


def tacaw_intensity(wf_layer: np.ndarray, time: np.ndarray, workers=1):
    """|fftshift_t FFT_t(psi - <psi>_t)|^2 and the THz axis, src/postprocessing/tacaw_data.py:82-104.
    wf_layer: (P, T, nx, ny) complex."""
    dt = time[1] - time[0]
    freqs = np.fft.fftshift(np.fft.fftfreq(len(time), d=dt))
    mean = np.mean(wf_layer, axis=1, keepdims=True)
    spec = np.fft.fftshift(sfft.fft(wf_layer - mean, axis=1, workers=workers), axes=1)
    return np.abs(spec) ** 2, freqs


# --- reducers, src/postprocessing/tacaw_data.py:109-353 ------------------------------
def spectrum(intensity, probe_index=None):
    """Sum over (kx,ky); mean over probes if probe_index is None (:109-143)."""
    per_probe = intensity.sum(axis=(2, 3))
    return per_probe.mean(axis=0) if probe_index is None else per_probe[probe_index]


def spectrum_image(intensity, freqs, frequency, probe_indices=None):
    """Nearest-frequency plane summed over k, per probe (:145-179)."""
    fi = int(np.argmin(np.abs(freqs - frequency)))
    idx = range(intensity.shape[0]) if probe_indices is None else probe_indices
    return np.array([intensity[p, fi].sum() for p in idx])


def diffraction(intensity, probe_index=None):
    """Sum over frequency (:183-217)."""
    per_probe = intensity.sum(axis=1)
    return per_probe.mean(axis=0) if probe_index is None else per_probe[probe_index]


def spectral_diffraction(intensity, freqs, frequency, probe_index=None):
    """Nearest-frequency (kx,ky) plane (:219-255)."""
    fi = int(np.argmin(np.abs(freqs - frequency)))
    return intensity[:, fi].mean(axis=0) if probe_index is None else intensity[probe_index, fi]


def masked_spectrum(intensity, mask, probe_index=None):
    """Sum over k of intensity*mask (:257-299; the reference's shape check reads attributes that do
    not exist, the arithmetic below is what it does once past that check)."""
    per_probe = (intensity * mask[None, None]).sum(axis=(2, 3))
    return per_probe.mean(axis=0) if probe_index is None else per_probe[probe_index]


def dispersion(intensity, kxs, kys, kx_path, ky_path, probe_index=None):
    """Nearest-(kx,ky) gather -> (n_freq, n_k) (:301-353)."""
    ix = np.array([int(np.argmin(np.abs(kxs - v))) for v in kx_path])
    iy = np.array([int(np.argmin(np.abs(kys - v))) for v in ky_path])
    sel = intensity[:, :, ix, iy]                            # (P, T, n_k)
    return sel.mean(axis=0) if probe_index is None else sel[probe_index]


def haadf_adf(wf, kxs, kys, probe_positions, eV, collection_angle=45):
    """HAADFData.calculateADF, src/postprocessing/haadf_data.py:43-65: mean over frames of
    sum_k |psi * (q > radius)| for the probe nearest each unique (x, y)."""
    probe_positions = np.asarray(probe_positions)
    ux = np.asarray(sorted(set(probe_positions[:, 0])))
    uy = np.asarray(sorted(set(probe_positions[:, 1])))
    q = np.sqrt(kxs[:, None] ** 2 + kys[None, :] ** 2)
    radius = (collection_angle * 1e-3) / wavelength(eV)
    mask = np.zeros(q.shape)
    mask[q > radius] = 1
    adf = np.zeros((len(ux), len(uy)))
    for i, x in enumerate(ux):
        for j, y in enumerate(uy):
            p = int(np.argmin(np.sqrt(np.sum((probe_positions - np.array([x, y])[None]) ** 2, axis=1))))
            exits = wf[p, :, :, :, -1]
            adf[i, j] = np.mean(np.sum(np.abs(exits * mask[None]), axis=(1, 2)))
    return adf, ux, uy
