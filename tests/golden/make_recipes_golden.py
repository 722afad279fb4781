"""Replay the reference's own unit-test recipes (src/unittests/01_potentials.py ... 07_defocus.py) WITH THE REFERENCE'S
CODE on a seeded synthetic hBN/graphene trajectory (the recipes' input files hBN_truncated.lammpstrj / their stored
.npy results are absent from the checkout, .MISSING_LARGE_BLOBS) and store what each recipe saves or plots:

    python tests/golden/make_recipes_golden.py        ->  tests/golden/recipes.npz

Build-container only (needs /root/reference).  tests/test_reference_recipes.py replays the same recipes with
pyslice_b200 on the GPU and applies the recipes' own residual, sum((|F|-|D|)^2)/sum(|F|^2) <= 1e-6.
"""
import os
import sys
import tempfile

import numpy as np

REF = os.environ.get("PYSLICE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, REF)

from src.multislice.calculators import MultisliceCalculator  # noqa: E402
from src.multislice.multislice import Probe, Propagate, create_batched_probes, probe_grid  # noqa: E402
from src.multislice.potentials import Potential, gridFromTrajectory  # noqa: E402
from src.multislice.trajectory import Trajectory as RefTrajectory  # noqa: E402
from src.postprocessing.haadf_data import HAADFData  # noqa: E402
from src.postprocessing.tacaw_data import TACAWData  # noqa: E402

from tests.recipes_input import A, B, NAMES, recipe_trajectory  # noqa: E402


def to_ref(traj):
    return RefTrajectory(atom_types=traj.atom_types, positions=traj.positions, velocities=traj.velocities,
                         box_matrix=traj.box_matrix, timestep=traj.timestep)


def npy(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


def main():
    os.chdir(tempfile.mkdtemp(prefix="pyslice_ref_"))
    out = {}
    trajectory = to_ref(recipe_trajectory())
    names = [NAMES[int(z)] for z in trajectory.atom_types]

    # 01_potentials.py
    xs, ys, zs, lx, ly, lz = gridFromTrajectory(trajectory, sampling=0.1, slice_thickness=0.5)
    potential = Potential(xs, ys, zs, trajectory.positions[0], names, kind="kirkland")
    out["r01_potential"] = npy(potential.to_cpu())[::3, ::3, :].astype(np.float32)

    # 02_propagate.py
    probe = Probe(xs, ys, mrad=5, eV=100e3)
    out["r02_exit"] = npy(Propagate(probe, potential)).astype(np.complex64)[::2, ::2]

    # 03_manyprobes.py
    cut = trajectory.slice_positions([0, 4 * A], [0, 3 * B])
    cxs, cys, czs, *_ = gridFromTrajectory(cut, sampling=0.1, slice_thickness=0.5)
    probe30 = Probe(cxs, cys, mrad=30, eV=100e3)
    x, y = np.meshgrid(np.linspace(A, 3 * A, 16), np.linspace(B, 2 * B, 16))
    xy = np.reshape([x, y], (2, len(x.flat))).T
    many = create_batched_probes(probe30, xy)
    cnames = [NAMES[int(z)] for z in cut.atom_types]
    cpot = Potential(cxs, cys, czs, cut.positions[0], cnames, kind="kirkland")
    res = npy(Propagate(many, cpot))
    out["r03_exit_sum"] = np.sum(np.absolute(res), axis=0).astype(np.float32)        # what the recipe plots and saves
    out["r03_exit_some"] = res[::37].astype(np.complex64)[:, ::2, ::2]

    # 04_haadf.py
    order = np.arange(cut.n_frames)
    np.random.seed(5)
    np.random.shuffle(order)
    three = cut.slice_timesteps(order[:3])
    hxy = probe_grid([A, 3 * A], [B, 2 * B], 14, 16)
    calc = MultisliceCalculator(force_cpu=True)
    calc.setup(three, aperture=30, voltage_eV=100e3, sampling=.1, slice_thickness=.5, probe_positions=hxy)
    haadf = HAADFData(calc.run())
    out["r04_adf"] = npy(haadf.calculateADF(preview=False)).astype(np.float64)
    out["r04_frames"] = order[:3]

    # 05_tacaw.py
    os.chdir(tempfile.mkdtemp(prefix="pyslice_ref_"))
    calc = MultisliceCalculator(force_cpu=True)
    calc.setup(trajectory, aperture=0, voltage_eV=100e3, sampling=.1, slice_thickness=.5)
    tacaw = TACAWData(calc.run())
    out["r05_frequencies"] = np.asarray(tacaw.frequencies)
    out["r05_slice"] = np.asarray(npy(tacaw.intensity)[0, 7, :, :] ** .1).astype(np.float32)
    out["r05_spectrum"] = np.asarray(tacaw.spectrum())

    # 07_defocus.py
    one = trajectory.slice_timesteps([0])
    dprobe = Probe(xs, ys, mrad=30, eV=100e3)
    dprobe.defocus(10 * 1e2)
    dpot = Potential(xs, ys, zs, one.positions[0], names, kind="kirkland")
    out["r07_exit_abs"] = np.absolute(npy(Propagate(dprobe, dpot))).astype(np.float32)

    path = os.path.join(HERE, "recipes.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
