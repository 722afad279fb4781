"""`TrajectoryLoader` with the reference's constructor, cache-file naming and `.npy` wire format
(src/io/loader.py:24-182), restricted to what this engine needs: trajectories come in through the
`<stem>.positions.npy / .velocities.npy / .atom_types.npy / .box_matrix.npy` cache the reference
writes after its first parse of a LAMMPS / CIF file.  Parsing those text formats needs OVITO / ASE
and is out of scope: without a cache `load()` raises ImportError naming the four
files, exactly where the reference would have called OVITO.

`save()` writes the same four files from a `Trajectory`, so synthetic or externally parsed data can be
handed to either implementation."""
from __future__ import annotations

import logging
from pathlib import Path
from typing import Dict, Optional, Union

import numpy as np

from .. import hostmath
from ..multislice.trajectory import Trajectory

logger = logging.getLogger(__name__)


class TrajectoryLoader:
    def __init__(self, filename: str, timestep: Optional[float] = None,
                 atom_mapping: Optional[Dict[int, Union[int, str]]] = None,
                 atomic_numbers: Optional[Dict[int, int]] = None,
                 element_names: Optional[Dict[int, str]] = None):
        """Same arguments, defaults and errors as the reference (loader.py:25-61); the trajectory file
        itself must exist even when only its cache is read, as there."""
        if timestep is not None and timestep <= 0:
            raise ValueError("timestep must be positive if specified.")
        self.filepath = Path(filename)
        if not self.filepath.exists():
            raise FileNotFoundError(f"Trajectory file not found: {filename}")
        self.timestep = timestep if timestep is not None else 1.0
        if atomic_numbers is not None:
            logger.warning("atomic_numbers parameter is deprecated. Use atom_mapping instead.")
            atom_mapping = atomic_numbers
        elif element_names is not None:
            logger.warning("element_names parameter is deprecated. Use atom_mapping instead.")
            atom_mapping = element_names
        self.atomic_numbers = self._process_atom_mapping(atom_mapping)

    @staticmethod
    def _process_atom_mapping(mapping):
        """loader.py:63-82: element names -> atomic numbers, range check for integers."""
        if mapping is None:
            return None
        result = {}
        for atom_type, value in mapping.items():
            if isinstance(value, str):
                result[atom_type] = hostmath.atomic_number(value)
            elif isinstance(value, (int, np.integer)):
                if not (1 <= value <= 118):
                    raise ValueError(f"Invalid atomic number {value} for type {atom_type}. Must be between 1 and 118.")
                result[atom_type] = int(value)
            else:
                raise ValueError(f"Invalid mapping value {value} for type {atom_type}. Must be int (atomic number) or str (element name).")
        return result

    def _apply_atomic_mapping(self, atom_types: np.ndarray) -> np.ndarray:
        """loader.py:84-102 (the reference applies it while parsing; cached arrays are stored mapped)."""
        if self.atomic_numbers is None:
            return atom_types
        mapped = atom_types.copy()
        unmapped = []
        for t in np.unique(atom_types):
            if t in self.atomic_numbers:
                mapped[atom_types == t] = self.atomic_numbers[t]
            else:
                unmapped.append(t)
        if unmapped:
            logger.warning(f"No mapping provided for atom types {unmapped}.")
        return mapped

    def _get_cache_files(self) -> Dict[str, Path]:
        """loader.py:104-112."""
        stem = self.filepath.parent / self.filepath.stem
        return {k: stem.with_suffix(f".{k}.npy") for k in ("positions", "velocities", "atom_types", "box_matrix")}

    def _load_from_cache(self) -> Optional[Trajectory]:
        """loader.py:114-145: None if a file is missing or the arrays do not form a valid Trajectory."""
        files = self._get_cache_files()
        if not all(f.exists() for f in files.values()):
            return None
        try:
            pos = np.load(files["positions"])
            vel = np.load(files["velocities"])
            atom_types = np.load(files["atom_types"])
            box = np.load(files["box_matrix"])
            if box.shape != (3, 3):
                raise ValueError(f"Invalid box_matrix shape: {box.shape}")
            traj = Trajectory(atom_types=atom_types, positions=pos, velocities=vel, box_matrix=box, timestep=self.timestep)
            logger.info(f"Loaded: {traj.n_frames} frames, {traj.n_atoms} atoms")
            return traj
        except Exception as e:          # as the reference: a bad cache is not fatal there, it re-parses
            logger.warning(f"Cache loading failed: {e}")
            return None

    def load(self) -> Trajectory:
        traj = self._load_from_cache()
        if traj is not None:
            return traj
        names = ", ".join(f.name for f in self._get_cache_files().values())
        raise ImportError(
            f"{self.filepath.name}: no usable .npy cache ({names}) next to the file, and parsing LAMMPS/CIF text "
            "needs OVITO/ASE, which pyslice_b200 does not bundle. Parse once with the reference (it writes this "
            "cache) or write it with TrajectoryLoader.save().")

    def save(self, trajectory: Trajectory) -> None:
        """loader.py:147-157: the cache writer."""
        files = self._get_cache_files()
        files["positions"].parent.mkdir(parents=True, exist_ok=True)
        np.save(files["positions"], trajectory.positions)
        np.save(files["velocities"], trajectory.velocities)
        np.save(files["atom_types"], trajectory.atom_types)
        np.save(files["box_matrix"], trajectory.box_matrix)
