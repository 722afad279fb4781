"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 (b is the reference)."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def small64_traj():
    from pyslice_b200 import synthetic
    return synthetic.random_trajectory(n_atoms=200, box=(6.35, 6.35, 4.1), n_frames=3, seed=11, stray=True)


def tacaw48_traj():
    from pyslice_b200 import synthetic
    return synthetic.random_trajectory(n_atoms=120, box=(4.75, 4.75, 3.2), n_frames=12, seed=5, types=(5, 7))


def si_c1_traj(n_frames=1):
    from pyslice_b200 import synthetic
    return synthetic.silicon_trajectory(cells=(5, 5, 10), a=5.11, n_frames=n_frames, seed=0)
