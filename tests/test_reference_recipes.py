"""The reference's own unit-test recipes (src/unittests/01_potentials.py, 02_propagate.py, 03_manyprobes.py,
04_haadf.py, 05_tacaw.py, 07_defocus.py) replayed call for call with pyslice_b200 on the GPU -- same functions, same
arguments, same residual `sum((|F|-|D|)^2)/sum(|F|^2) <= 1e-6` -- against tests/golden/recipes.npz, which
tests/golden/make_recipes_golden.py produced by running the recipes with the REFERENCE'S code on the same seeded
trajectory (the recipes' own input files are absent from the checkout, .MISSING_LARGE_BLOBS)."""
import numpy as np
import pytest
import torch

from tests.helpers import golden
from tests.recipes_input import A, B, NAMES, recipe_trajectory

pytestmark = pytest.mark.gpu


def dz(ary, previous):
    """the recipes' scaling-resistant residual (e.g. 01_potentials.py:31-33)"""
    F, D = np.absolute(np.asarray(ary, dtype=np.complex128)), np.absolute(np.asarray(previous, dtype=np.complex128))
    return np.sum((F - D) ** 2) / np.sum(F ** 2)


def npy(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


@pytest.fixture(scope="module")
def g():
    return golden("recipes.npz")


@pytest.fixture(scope="module")
def trajectory():
    return recipe_trajectory()


def test_01_potentials(g, trajectory):
    from pyslice_b200.multislice.potentials import Potential, gridFromTrajectory
    positions = trajectory.positions[0]
    atom_types = [NAMES[int(z)] for z in trajectory.atom_types]
    xs, ys, zs, lx, ly, lz = gridFromTrajectory(trajectory, sampling=0.1, slice_thickness=0.5)
    potential = Potential(xs, ys, zs, positions, atom_types, kind="kirkland")
    ary = potential.to_cpu()
    assert ary.shape == (154, 171, 14)
    assert dz(npy(ary)[::3, ::3, :], g["r01_potential"]) < 1e-6


def test_02_propagate(g, trajectory):
    from pyslice_b200.multislice.multislice import Probe, Propagate
    from pyslice_b200.multislice.potentials import Potential, gridFromTrajectory
    xs, ys, zs, lx, ly, lz = gridFromTrajectory(trajectory, sampling=0.1, slice_thickness=0.5)
    probe = Probe(xs, ys, mrad=5, eV=100e3)
    potential = Potential(xs, ys, zs, trajectory.positions[0], [NAMES[int(z)] for z in trajectory.atom_types], kind="kirkland")
    result = Propagate(probe, potential)
    ary = result.cpu().numpy() if hasattr(result, "cpu") else np.asarray(result)
    assert dz(ary[::2, ::2], g["r02_exit"]) < 1e-6
    assert np.linalg.norm(ary[::2, ::2] - g["r02_exit"]) / np.linalg.norm(g["r02_exit"]) < 1e-4     # phases too


def test_03_manyprobes(g, trajectory):
    from pyslice_b200.multislice.multislice import Probe, Propagate, create_batched_probes
    from pyslice_b200.multislice.potentials import Potential, gridFromTrajectory
    cut = trajectory.slice_positions([0, 4 * A], [0, 3 * B])
    xs, ys, zs, lx, ly, lz = gridFromTrajectory(cut, sampling=0.1, slice_thickness=0.5)
    probe = Probe(xs, ys, mrad=30, eV=100e3)
    x, y = np.meshgrid(np.linspace(A, 3 * A, 16), np.linspace(B, 2 * B, 16))
    xy = np.reshape([x, y], (2, len(x.flat))).T
    probes_many = create_batched_probes(probe, xy)
    potential = Potential(xs, ys, zs, cut.positions[0], [NAMES[int(z)] for z in cut.atom_types], kind="kirkland")
    result = npy(Propagate(probes_many, potential))
    assert result.shape[0] == 256
    assert dz(np.sum(np.absolute(result), axis=0), g["r03_exit_sum"]) < 1e-6
    assert dz(result[::37][:, ::2, ::2], g["r03_exit_some"]) < 1e-6


def test_04_haadf(g, trajectory):
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.multislice.multislice import probe_grid
    from pyslice_b200.postprocessing.haadf_data import HAADFData
    cut = trajectory.slice_positions([0, 4 * A], [0, 3 * B])
    slice_timesteps = np.arange(cut.n_frames)
    np.random.seed(5)
    np.random.shuffle(slice_timesteps)
    slice_timesteps = slice_timesteps[:3]
    assert np.array_equal(slice_timesteps, g["r04_frames"])
    three = cut.slice_timesteps(slice_timesteps)
    xy = probe_grid([A, 3 * A], [B, 2 * B], 14, 16)
    calculator = MultisliceCalculator()
    calculator.setup(three, aperture=30, voltage_eV=100e3, sampling=.1, slice_thickness=.5, probe_positions=xy)
    exitwaves = calculator.run()
    haadf = HAADFData(exitwaves)
    ary = np.asarray(npy(haadf.calculateADF(preview=False)))
    assert ary.shape == (14, 16)
    assert dz(ary, g["r04_adf"]) < 1e-6
    assert np.abs(ary / g["r04_adf"] - 1).max() < 1e-4
    # the same scan as a detector-only run (no exit-wave cube): 224 probes x 3 frames reduced after each exit FFT
    calculator.setup(three, aperture=30, voltage_eV=100e3, sampling=.1, slice_thickness=.5, probe_positions=xy,
                     adf_collection_angle=45)
    lean = np.asarray(npy(HAADFData(calculator.run()).calculateADF(preview=False)))
    assert dz(lean, g["r04_adf"]) < 1e-6 and np.abs(lean / ary - 1).max() < 1e-5


def test_05_tacaw(g, trajectory):
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    calculator = MultisliceCalculator()
    calculator.setup(trajectory, aperture=0, voltage_eV=100e3, sampling=.1, slice_thickness=.5)
    exitwaves = calculator.run()
    tacaw = TACAWData(exitwaves)
    assert np.allclose(tacaw.frequencies, g["r05_frequencies"], rtol=0, atol=1e-9)
    ary = np.asarray(npy(tacaw.intensity[0, 7, :, :] ** .1))
    assert dz(ary, g["r05_slice"]) < 1e-6
    spec = np.asarray(tacaw.spectrum())
    keep = [i for i in range(12) if i != 6]                 # the zero-frequency bin is rounding noise in both
    assert np.abs(spec[keep] / g["r05_spectrum"][keep] - 1).max() < 1e-3


def test_07_defocus(g, trajectory):
    from pyslice_b200.multislice.multislice import Probe, Propagate
    from pyslice_b200.multislice.potentials import Potential, gridFromTrajectory
    one = trajectory.slice_timesteps([0])
    xs, ys, zs, lx, ly, lz = gridFromTrajectory(one, sampling=0.1, slice_thickness=0.5)
    potential = Potential(xs, ys, zs, one.positions[0], [NAMES[int(z)] for z in one.atom_types], kind="kirkland")
    probe = Probe(xs, ys, mrad=30, eV=100e3)
    probe.defocus(10 * 1e2)
    result = Propagate(probe, potential)
    plot_result = np.absolute(npy(result))
    assert dz(plot_result, g["r07_exit_abs"]) < 1e-6
