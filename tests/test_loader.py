"""TrajectoryLoader .npy cache (reference src/io/loader.py:104-182): same file names and arrays as the
reference writes, so a cache produced by either side loads in the other."""
import sys

import numpy as np
import pytest

from pyslice_b200 import synthetic
from pyslice_b200.io.loader import TrajectoryLoader


def test_cache_round_trip_and_names(tmp_path):
    traj = synthetic.random_trajectory(n_atoms=30, box=(6.35, 6.35, 4.1), n_frames=4, seed=2)
    dump = tmp_path / "run.lammpstrj"
    dump.write_text("placeholder: only the cache is read\n")
    loader = TrajectoryLoader(str(dump), timestep=0.02)
    loader.save(traj)
    names = sorted(p.name for p in tmp_path.iterdir())
    assert names == ["run.atom_types.npy", "run.box_matrix.npy", "run.lammpstrj", "run.positions.npy", "run.velocities.npy"]
    back = TrajectoryLoader(str(dump), timestep=0.02).load()
    assert np.array_equal(back.positions, traj.positions) and np.array_equal(back.atom_types, traj.atom_types)
    assert np.array_equal(back.box_matrix, traj.box_matrix) and back.timestep == 0.02
    assert back.n_frames == 4 and back.n_atoms == 30


def test_reference_loader_reads_our_cache(tmp_path):
    """the reference's own TrajectoryLoader (when the checkout is present) accepts the files we write"""
    import os
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference checkout not present on this box")
    sys.path.insert(0, "/root/reference")
    try:
        from src.io.loader import TrajectoryLoader as RefLoader
    except Exception as e:                      # tqdm / optional imports
        pytest.skip(f"reference loader not importable here: {e}")
    finally:
        sys.path.remove("/root/reference")
    traj = synthetic.random_trajectory(n_atoms=12, box=(5.0, 5.0, 3.0), n_frames=3, seed=4)
    dump = tmp_path / "md.lammpstrj"
    dump.write_text("x\n")
    TrajectoryLoader(str(dump)).save(traj)
    ref = RefLoader(str(dump), timestep=0.5).load()
    assert np.array_equal(ref.positions, traj.positions) and np.array_equal(ref.atom_types, traj.atom_types)
    assert np.array_equal(ref.box_matrix, traj.box_matrix)


def test_errors_match_the_reference(tmp_path):
    with pytest.raises(FileNotFoundError):
        TrajectoryLoader(str(tmp_path / "missing.lammpstrj"))
    f = tmp_path / "a.lammpstrj"
    f.write_text("x\n")
    with pytest.raises(ValueError):
        TrajectoryLoader(str(f), timestep=0.0)
    with pytest.raises(ValueError):
        TrajectoryLoader(str(f), atom_mapping={1: 300})
    assert TrajectoryLoader(str(f), atom_mapping={1: "Si", 2: 8}).atomic_numbers == {1: 14, 2: 8}
    with pytest.raises(ImportError):
        TrajectoryLoader(str(f)).load()         # no cache, no parser
    # a corrupt cache is reported like a missing one
    for k in ("positions", "velocities", "atom_types"):
        np.save(tmp_path / f"a.{k}.npy", np.zeros((2, 3, 3)))
    np.save(tmp_path / "a.box_matrix.npy", np.zeros((2, 2)))
    with pytest.raises(ImportError):
        TrajectoryLoader(str(f)).load()
