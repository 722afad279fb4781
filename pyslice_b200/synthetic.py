"""Seeded synthetic MD trajectories for parity tests and the benchmark.

Boxes are engineered as ``L = N*sampling - sampling/2`` so that the reference grid rule
``n = int(L/sampling)+1`` (reference src/multislice/potentials.py:123-125) yields exactly N
pixels; crystals are offset by a quarter cell in z so thermally displaced atoms stay inside
``[0, lz)``.  All generators are pure NumPy (host side) and deterministic for a given seed.
"""
from __future__ import annotations

import numpy as np

from .multislice.trajectory import Trajectory

_DIAMOND = np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0],
                     [.25, .25, .25], [.25, .75, .75], [.75, .25, .75], [.75, .75, .25]])


def _phonon_displacements(rng, n_frames, n_atoms, timestep, n_modes=4, amp=0.03, noise=0.01,
                          frames=None):
    """Sum of `n_modes` sinusoids (3-16 THz) with random per-atom polarisation and phase plus
    white noise -> a trajectory with a real THz spectrum.  `frames` selects a window of the
    infinite time series (used to give each GPU rank its own block of the same trajectory)."""
    freqs = rng.uniform(3.0, 16.0, size=n_modes)                       # THz (1/ps)
    pol = rng.normal(size=(n_modes, n_atoms, 3))
    pol /= np.linalg.norm(pol, axis=2, keepdims=True)
    phase = rng.uniform(0, 2 * np.pi, size=(n_modes, n_atoms))
    f0 = 0 if frames is None else frames[0]
    t = (np.arange(n_frames) + f0) * timestep
    disp = np.zeros((n_frames, n_atoms, 3))
    for m in range(n_modes):
        disp += amp * np.sin(2 * np.pi * freqs[m] * t[:, None] + phase[m][None, :])[:, :, None] * pol[m][None]
    if noise > 0:
        nrng = np.random.default_rng([int(rng.integers(1 << 31)), f0])
        disp += nrng.normal(scale=noise, size=disp.shape)
    return disp


def silicon_trajectory(cells=(5, 5, 10), a=5.11, n_frames=20, seed=0, timestep=0.01,
                       displacement="iid", sigma=0.05, frames=None) -> Trajectory:
    """Diamond-cubic Si supercell (8 atoms per cell, Z=14).

    cells=(5,5,10), a=5.11  -> 2 000 atoms, 256 x 256 x 103 grid   (config C1)
    cells=(5,5,50), a=5.11  -> 10 000 atoms, 256 x 256 x 512 grid  (config C2)
    cells=(20,20,12), a=5.1175 -> 38 400 atoms, 1024 x 1024 x 123  (config C4)
    """
    rng = np.random.default_rng(seed)
    cx, cy, cz = cells
    ijk = np.stack(np.meshgrid(np.arange(cx), np.arange(cy), np.arange(cz), indexing="ij"), -1).reshape(-1, 3)
    base = ((ijk[:, None, :] + _DIAMOND[None]) * a).reshape(-1, 3)
    base[:, 2] += 0.125 * a                         # keep displaced atoms inside [0, lz)
    n_atoms = base.shape[0]
    if displacement == "iid":
        disp = rng.normal(scale=sigma, size=(n_frames, n_atoms, 3))
    else:
        disp = _phonon_displacements(rng, n_frames, n_atoms, timestep, frames=frames)
    positions = base[None] + disp
    box = np.diag([cx * a, cy * a, cz * a]).astype(np.float64)
    return Trajectory(atom_types=np.full(n_atoms, 14, dtype=np.int64), positions=positions,
                      velocities=np.zeros_like(positions), box_matrix=box, timestep=timestep)


def hbn_graphene_trajectory(cells=(20, 12), n_layers=10, n_frames=100, seed=2, timestep=0.01,
                            a=2.5575, b=4.2625, spacing=3.33, frames=None) -> Trajectory:
    """hBN / graphene stack on a rectangular 4-atom cell (configs C3 / C5):
    20 x 12 cells -> 51.15 x 51.15 A -> 512 x 512 pixels; 10 layers x 3.33 A -> 67 slices.
    Even layers are hBN (B, N alternating), odd layers graphene (C)."""
    rng = np.random.default_rng(seed)
    cx, cy = cells
    frac = np.array([[0, 0], [.5, .5], [0, 1 / 3], [.5, 5 / 6]])
    ij = np.stack(np.meshgrid(np.arange(cx), np.arange(cy), indexing="ij"), -1).reshape(-1, 2)
    xy = ((ij[:, None, :] + frac[None]) * np.array([a, b])).reshape(-1, 2)
    sub = np.tile(np.array([0, 0, 1, 1]), cx * cy)              # sublattice A / B
    pos, types = [], []
    for layer in range(n_layers):
        z = (layer + 0.5) * spacing
        pos.append(np.concatenate([xy, np.full((len(xy), 1), z)], axis=1))
        types.append(np.where(sub == 0, 5, 7) if layer % 2 == 0 else np.full(len(xy), 6))
    base = np.concatenate(pos)
    atom_types = np.concatenate(types).astype(np.int64)
    disp = _phonon_displacements(rng, n_frames, len(base), timestep, frames=frames)
    box = np.diag([cx * a, cy * b, n_layers * spacing]).astype(np.float64)
    positions = base[None] + disp
    return Trajectory(atom_types=atom_types, positions=positions, velocities=np.zeros_like(positions),
                      box_matrix=box, timestep=timestep)


def random_trajectory(n_atoms=200, box=(6.35, 6.35, 4.1), n_frames=3, seed=0, types=(6, 14, 31),
                      timestep=0.01, stray=False) -> Trajectory:
    """Small random-gas trajectory for fast parity tests.  `stray=True` adds atoms with z<0,
    z>=lz and z exactly on slice bounds to exercise the drop / gap / overlap rules of the
    reference binning (reference src/multislice/potentials.py:304-310)."""
    rng = np.random.default_rng(seed)
    box = np.asarray(box, dtype=np.float64)
    positions = rng.uniform(0, 1, size=(n_frames, n_atoms, 3)) * box[None, None, :]
    if stray:
        nz = int(box[2] / 0.5) + 1
        zs = np.linspace(0, box[2], nz, endpoint=False)
        dz = zs[1] - zs[0]
        edges = np.concatenate([zs - dz / 2, zs + dz / 2, [zs[-1] + dz, 0.0, -0.01, box[2] + 0.3]])
        k = min(len(edges), n_atoms // 2)
        positions[:, :k, 2] = edges[:k][None, :]
        positions[:, k:k + 4, 2] = np.nextafter(edges[:4], np.inf)[None, :]
        positions[:, k + 4:k + 8, 2] = np.nextafter(edges[4:8], -np.inf)[None, :]
    atom_types = rng.choice(np.asarray(types), size=n_atoms).astype(np.int64)
    return Trajectory(atom_types=atom_types, positions=positions, velocities=np.zeros_like(positions),
                      box_matrix=np.diag(box), timestep=timestep)
