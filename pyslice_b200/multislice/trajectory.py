"""`Trajectory`: the input container of the hot path (host side, NumPy).

Mirrors the reference dataclass (reference src/multislice/trajectory.py:9-237): same field
names, same validation errors, same helper methods.  It never touches the GPU; the engine
uploads `positions` frame blocks itself.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np


@dataclass
class Trajectory:
    atom_types: np.ndarray      # (A,)
    positions: np.ndarray       # (T, A, 3) Angstrom
    velocities: np.ndarray      # (T, A, 3)
    box_matrix: np.ndarray      # (3, 3), orthogonal boxes use the diagonal only
    timestep: float             # ps

    def __post_init__(self):
        self._validate_shapes()

    # reference trajectory.py:20-40 -- same messages so callers' error handling is unchanged
    def _validate_shapes(self):
        for name in ("positions", "velocities"):
            arr = getattr(self, name)
            if arr.ndim != 3 or arr.shape[2] != 3:
                raise ValueError(f"{name} must be (frames, atoms, 3), got {arr.shape}")
        if self.atom_types.ndim != 1:
            raise ValueError(f"atom_types must be 1D, got {self.atom_types.ndim}D")
        if self.box_matrix.shape != (3, 3):
            raise ValueError(f"box_matrix must be (3, 3), got {self.box_matrix.shape}")
        tp, ap = self.positions.shape[:2]
        tv, av = self.velocities.shape[:2]
        if tp != tv:
            raise ValueError(f"Frame count mismatch: {tp} vs {tv}")
        if not (ap == av == len(self.atom_types)):
            raise ValueError(f"Atom count mismatch: {ap}, {av}, {len(self.atom_types)}")

    @property
    def n_frames(self) -> int:
        return self.positions.shape[0]

    @property
    def n_atoms(self) -> int:
        return len(self.atom_types)

    @property
    def box_tilts(self) -> np.ndarray:
        m = self.box_matrix
        return np.array([m[0, 1], m[0, 2], m[1, 2]])

    def get_mean_positions(self) -> np.ndarray:
        if self.n_frames == 0:
            return np.empty((0, 3), dtype=self.positions.dtype)
        return self.positions.mean(axis=0)

    def _like(self, **changes) -> "Trajectory":
        kw = dict(atom_types=self.atom_types, positions=self.positions, velocities=self.velocities,
                  box_matrix=self.box_matrix, timestep=self.timestep)
        kw.update(changes)
        return Trajectory(**kw)

    def tile_positions(self, repeats: Tuple[int, int, int]) -> "Trajectory":
        """Repeat the cell (rx, ry, rz) times (reference trajectory.py:63-111)."""
        rx, ry, rz = repeats
        shifts = [self.box_matrix @ np.array([i, j, k])
                  for i in range(rx) for j in range(ry) for k in range(rz)]
        box = self.box_matrix.copy()
        box[:, 0] *= rx
        box[:, 1] *= ry
        box[:, 2] *= rz
        return self._like(
            atom_types=np.concatenate([self.atom_types] * len(shifts)),
            positions=np.concatenate([self.positions + s for s in shifts], axis=1),
            velocities=np.concatenate([self.velocities] * len(shifts), axis=1),
            box_matrix=box)

    def slice_positions(self, x_range: Optional[Sequence[float]] = None,
                        y_range: Optional[Sequence[float]] = None,
                        z_range: Optional[Sequence[float]] = None) -> "Trajectory":
        """Keep atoms whose MEAN position lies inside the closed ranges; the box diagonal of each
        filtered axis becomes (max-min) (reference trajectory.py:124-194)."""
        ranges = (x_range, y_range, z_range)
        if self.n_atoms == 0 or all(r is None for r in ranges):
            return self
        for r, name in zip(ranges, "XYZ"):
            if r is not None and r[0] > r[1]:
                raise ValueError(f"{name} range invalid: min={r[0]} > max={r[1]}")
        mean = self.get_mean_positions()
        keep = np.ones(self.n_atoms, dtype=bool)
        box = self.box_matrix.copy()
        for ax, r in enumerate(ranges):
            if r is None:
                continue
            keep &= (mean[:, ax] >= r[0]) & (mean[:, ax] <= r[1])
            box[ax, ax] = r[1] - r[0]
        if not keep.any():
            desc = " AND ".join(f"{n}∈[{r[0]:.2f},{r[1]:.2f}]" for n, r in zip("XYZ", ranges) if r)
            raise ValueError(f"Filter {desc} resulted in 0 atoms")
        if keep.all():
            return self
        return self._like(atom_types=self.atom_types[keep], positions=self.positions[:, keep, :],
                          velocities=self.velocities[:, keep, :], box_matrix=box)

    def slice_timesteps(self, frame_indices: List[int]) -> "Trajectory":
        """Select frames by index (reference trajectory.py:196-224)."""
        if len(frame_indices) == 0:
            raise ValueError("frame_indices cannot be empty")
        if max(frame_indices) >= self.n_frames:
            raise ValueError(f"Frame index {max(frame_indices)} out of range [0, {self.n_frames-1}]")
        return self._like(positions=self.positions[frame_indices, :, :],
                          velocities=self.velocities[frame_indices, :, :])

    def generate_random_displacements(self, n_displacements, sigma):
        """Frozen-phonon style copies of frame 0 with uniform [0, sigma) offsets
        (reference trajectory.py:226-237; uses the global NumPy RNG like the reference)."""
        offsets = np.random.random(size=(n_displacements, self.n_atoms, 3)) * sigma
        return self._like(positions=self.positions[0] + offsets,
                          velocities=np.ones(n_displacements)[:, None, None] * self.velocities[0])
