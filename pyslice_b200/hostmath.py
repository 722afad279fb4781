"""Frame-invariant host tables of the engine, evaluated once per setup in float64 (NumPy).

These are the pieces of the reference that are *not* per-frame work: the grid rule, the slice
bounds used by the bit-exact binning, electron-optical constants, Kirkland form factors on the
k grid, the separable Fresnel propagator and the aperture mask.  They are computed with the same
expressions as the reference so that rounding decisions (grid sizes, `<` comparisons) are identical,
then handed to the CUDA kernels as float32 / complex64 tables.
"""
from __future__ import annotations

import os

import numpy as np

# reference src/multislice/multislice.py:31-34
M_ELECTRON = 9.109383e-31
Q_ELECTRON = 1.602177e-19
C_LIGHT = 299792458.0
H_PLANCK = 6.62607015e-34

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "kirkland_abcd.csv")
_TABLE = None

ELEMENTS = ["H", "He", "Li", "Be", "B", "C", "N", "O", "F", "Ne", "Na", "Mg", "Al", "Si", "P", "S", "Cl", "Ar",
            "K", "Ca", "Sc", "Ti", "V", "Cr", "Mn", "Fe", "Co", "Ni", "Cu", "Zn", "Ga", "Ge", "As", "Se", "Br", "Kr",
            "Rb", "Sr", "Y", "Zr", "Nb", "Mo", "Tc", "Ru", "Rh", "Pd", "Ag", "Cd", "In", "Sn", "Sb", "Te", "I", "Xe",
            "Cs", "Ba", "La", "Ce", "Pr", "Nd", "Pm", "Sm", "Eu", "Gd", "Tb", "Dy", "Ho", "Er", "Tm", "Yb",
            "Lu", "Hf", "Ta", "W", "Re", "Os", "Ir", "Pt", "Au", "Hg", "Tl", "Pb", "Bi", "Po", "At", "Rn",
            "Fr", "Ra", "Ac", "Th", "Pa", "U", "Np", "Pu", "Am", "Cm", "Bk", "Cf", "Es", "Fm", "Md", "No", "Lr"]


def atomic_number(kind) -> int:
    """Element symbol or number -> Z (reference potentials.py:98-111)."""
    if isinstance(kind, str):
        return ELEMENTS.index(kind) + 1
    return int(kind)


def kirkland_table() -> np.ndarray:
    """(103, 3, 4) rows [a_i, b_i, c_i, d_i]; FileNotFoundError if the data file is missing
    (reference potentials.py:155-156)."""
    global _TABLE
    if _TABLE is None:
        if not os.path.exists(_DATA):
            raise FileNotFoundError("Could not find kirkland parameter table " + _DATA)
        _TABLE = np.loadtxt(_DATA, delimiter=",", comments="#")[:, 1:].reshape(103, 3, 4)
    return _TABLE


def wavelength(eV: float) -> float:
    """Relativistic wavelength in Angstrom (reference multislice.py:41-42)."""
    return H_PLANCK * C_LIGHT / ((eV * Q_ELECTRON) ** 2 + 2 * eV * Q_ELECTRON * M_ELECTRON * C_LIGHT ** 2) ** 0.5 * 1e10


def interaction_sigma(eV: float) -> float:
    """sigma = 2 pi/(lambda eV) (E0+eV)/(2E0+eV) (reference multislice.py:258-260)."""
    e0 = M_ELECTRON * C_LIGHT ** 2 / Q_ELECTRON
    return (2 * np.pi) / (wavelength(eV) * eV) * (e0 + eV) / (2 * e0 + eV)


def grid_from_box(box_matrix, sampling=0.1, slice_thickness=0.5):
    """Reference grid rule potentials.py:113-131, evaluated with the identical expressions."""
    lx = box_matrix[0, 0]
    ly = box_matrix[1, 1]
    lz = box_matrix[2, 2]
    nx = int(lx / sampling) + 1
    ny = int(ly / sampling) + 1
    nz = int(lz / slice_thickness) + 1
    return (np.linspace(0, lx, nx, endpoint=False), np.linspace(0, ly, ny, endpoint=False),
            np.linspace(0, lz, nz, endpoint=False), lx, ly, lz)


def slice_bounds(zs):
    """lo/hi of every slice with the reference's float64 expressions (potentials.py:230,304-305)."""
    nz = len(zs)
    dz = zs[1] - zs[0] if nz > 1 else 0.5
    lo = np.array([zs[i] - dz / 2 if i > 0 else 0 for i in range(nz)], dtype=np.float64)
    hi = np.array([zs[i] + dz / 2 if i < nz - 1 else zs[-1] + dz for i in range(nz)], dtype=np.float64)
    return lo, hi, float(dz)


def kgrid(xs, ys):
    """fftfreq axes on the true pixel size (reference potentials.py:251-252)."""
    return np.fft.fftfreq(len(xs), d=xs[1] - xs[0]), np.fft.fftfreq(len(ys), d=ys[1] - ys[0])


def form_factor_table(kxs, kys, Zs) -> np.ndarray:
    """(len(Zs), nx, ny) float64 Kirkland f_e(q^2) (reference potentials.py:86-96)."""
    qsq = kxs[:, None] ** 2 + kys[None, :] ** 2
    out = np.empty((len(Zs),) + qsq.shape)
    tab = kirkland_table()
    for i, Z in enumerate(Zs):
        abcd = tab[Z - 1]
        acc = np.zeros_like(qsq)
        for a, b, c, d in abcd:
            acc += a / (qsq + b) + c * np.exp(-d * qsq)
        out[i] = acc
    return out


def propagator_tables(kxs, kys, wavelength_A, dz):
    """Separable Fresnel propagator exp(-i pi lambda dz k^2) = px[kx] * py[ky] (reference multislice.py:273-275)."""
    return (np.exp(-1j * np.pi * wavelength_A * dz * kxs ** 2), np.exp(-1j * np.pi * wavelength_A * dz * kys ** 2))


def aperture_mask(kxs, kys, mrad, wavelength_A) -> np.ndarray:
    """|k| < alpha/lambda with the reference's strict comparison (multislice.py:116-122)."""
    radius = (mrad * 1e-3) / wavelength_A
    return (np.sqrt(kxs[:, None] ** 2 + kys[None, :] ** 2) < radius)


def shift_ramps(kxs, kys, probe_positions):
    """exp(+2 pi i k p) ramps per probe (reference multislice.py:221-222; sign as in the reference)."""
    pos = np.asarray(probe_positions, dtype=np.float64).reshape(-1, 2)
    return (np.exp(2j * np.pi * kxs[None, :] * pos[:, 0:1]), np.exp(2j * np.pi * kys[None, :] * pos[:, 1:2]))
