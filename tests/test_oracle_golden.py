"""Pin the CPU oracle (oracle/pyslice_oracle.py) to outputs of the reference itself.

The golden files were produced by tests/golden/make_golden.py, which runs the unmodified
reference (torch CPU, complex128) in the build container.  Tolerances are float64 round-off
(different FFT/BLAS summation orders), far below the 1e-4 / 1e-3 parity budget of the engine.
"""
import numpy as np
import pytest

from oracle import pyslice_oracle as orc
from tests.helpers import golden, rel_l2, si_c1_traj, small64_traj, tacaw48_traj


def test_constants():
    # reference src/multislice/multislice.py:41-42 evaluated at 100 kV
    assert abs(orc.wavelength(100e3) - 0.0370143628) < 1e-9


def test_grid_rule_quirk():
    # n = int(L/sampling)+1 must be evaluated with Python float division: 0.3/0.1 == 2.9999999999999996
    xs, ys, zs, *_ = orc.grid_from_box(np.diag([0.3, 25.55, 51.1]), 0.1, 0.5)
    assert (len(xs), len(ys), len(zs)) == (3, 256, 103)
    assert xs[1] == 0.3 / 3 and zs[1] == 51.1 / 103


def test_small64_grid_and_potential():
    g = golden("small64.npz")
    traj = small64_traj()
    xs, ys, zs, *_ = orc.grid_from_box(traj.box_matrix)
    assert np.array_equal(xs, g["xs"]) and np.array_equal(ys, g["ys"]) and np.array_equal(zs, g["zs"])
    V = orc.potential(xs, ys, zs, traj.positions[0], traj.atom_types)
    assert V.shape == g["potential0"].shape
    assert rel_l2(V, g["potential0"]) < 1e-12


def test_small64_binning_has_edge_cases():
    traj = small64_traj()
    xs, ys, zs, *_ = orc.grid_from_box(traj.box_matrix)
    M = orc.bin_atoms(traj.positions[0][:, 2], zs)
    per_atom = M.sum(axis=0)
    assert (per_atom == 0).any(), "fixture must contain dropped atoms"
    assert (per_atom >= 1).any()


def test_small64_plane_wave_run():
    g = golden("small64.npz")
    traj = small64_traj()
    wf, grid = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=0.0,
                                  voltage_eV=100e3)
    assert wf.shape == g["wf_plane"].shape
    for f in range(wf.shape[1]):
        assert rel_l2(wf[0, f], g["wf_plane"][0, f]) < 1e-10
    kxs, kys, time = orc.wf_axes(len(grid["xs"]), len(grid["ys"]), 0.1, traj.n_frames, traj.timestep)
    assert np.array_equal(kxs, g["kxs"]) and np.array_equal(kys, g["kys"])
    assert np.allclose(time, g["time"], rtol=0, atol=0)


def test_small64_probe_run_and_haadf():
    g = golden("small64.npz")
    traj = small64_traj()
    xy = g["probe_xy"]
    wf, grid = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=30.0,
                                  voltage_eV=100e3, probe_positions=xy)
    assert rel_l2(grid["base_probe"], g["base_probe"]) < 1e-12
    for p in range(wf.shape[0]):
        for f in range(wf.shape[1]):
            assert rel_l2(wf[p, f], g["wf_probes"][p, f]) < 1e-10
    adf, ux, uy = orc.haadf_adf(wf, g["kxs"], g["kys"], xy, 100e3, 45)
    assert rel_l2(adf, g["adf"]) < 1e-5           # reference accumulates with float32 kxs labels


def test_tacaw48():
    g = golden("tacaw48.npz")
    traj = tacaw48_traj()
    wf, grid = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=0.0,
                                  voltage_eV=100e3)
    assert wf.shape == g["wf"].shape and wf.shape[2] == 48
    assert rel_l2(wf, g["wf"]) < 1e-10
    kxs, kys, time = orc.wf_axes(48, 48, 0.1, traj.n_frames, traj.timestep)
    inten, freqs = orc.tacaw_intensity(wf[..., 0], time)
    assert np.allclose(freqs, g["frequencies"], rtol=1e-15, atol=0)
    # DC bin is ~1e-19 in the reference (mean subtracted): compare it absolutely
    dc = inten.shape[1] // 2
    nz = [i for i in range(inten.shape[1]) if i != dc]
    assert rel_l2(inten[:, nz], g["intensity"][:, nz]) < 1e-9
    assert np.abs(inten[:, dc]).max() < 1e-12 * inten.max()
    assert rel_l2(orc.spectrum(inten)[nz], g["spectrum"][nz]) < 1e-9
    assert rel_l2(orc.spectrum(inten, 0)[nz], g["spectrum0"][nz]) < 1e-9
    assert rel_l2(orc.diffraction(inten), g["diffraction"]) < 1e-9
    assert rel_l2(orc.spectral_diffraction(inten, freqs, 20.0), g["spectral_diffraction"]) < 1e-9
    assert rel_l2(orc.spectrum_image(inten, freqs, 20.0), g["spectrum_image"]) < 1e-9
    disp = orc.dispersion(inten, kxs, kys, g["kx_path"], g["ky_path"])
    assert rel_l2(disp[nz], g["dispersion"][nz]) < 1e-9


@pytest.mark.parametrize("mrad", [1, 3, 5, 15, 30])
def test_probe_kat(mrad):
    # the reference's own known-answer recipe, src/unittests/00_probe.py:7-18
    g = golden("probe_kat.npz")
    xs = np.linspace(0, 50, 501)
    ys = np.linspace(0, 49, 491)
    p = orc.probe_array(xs, ys, mrad, 100e3)[::5, ::5]
    assert rel_l2(p, g[f"mrad{mrad}"]) < 5e-7      # golden stored as complex64


def test_si_c1_frame0():
    g = golden("si_c1.npz")
    traj = si_c1_traj()
    xs, ys, zs, *_ = orc.grid_from_box(traj.box_matrix)
    assert (len(xs), len(ys), len(zs)) == (256, 256, 103)
    V = orc.potential(xs, ys, zs, traj.positions[0], traj.atom_types, workers=4)
    assert rel_l2(V.sum(axis=2), g["pot_sum_z"]) < 1e-11
    assert rel_l2(V.sum(axis=(0, 1)), g["pot_sum_xy"]) < 1e-11
    assert rel_l2(V[:, :, [0, 1, 51, 102]], g["pot_planes"]) < 1e-6   # float32 golden
    psi = orc.propagate(np.ones((256, 256)), V, xs, ys, zs, 100e3, workers=4)
    wf = orc.exit_to_kspace(psi)[0]
    assert rel_l2(wf, g["wf"]) < 1e-6                                   # complex64 golden


# ---- the oracle against the reference's unit-test recipes (tests/golden/recipes.npz, produced by the reference's code) --

def _dz(ary, previous):
    F, D = np.absolute(np.asarray(ary, dtype=np.complex128)), np.absolute(np.asarray(previous, dtype=np.complex128))
    return np.sum((F - D) ** 2) / np.sum(F ** 2)


def test_oracle_replays_reference_recipes_01_02_05():
    """01_potentials.py, 02_propagate.py, 05_tacaw.py with the oracle on the recipes' seeded trajectory, under the
    recipes' own residual (float32 storage of the goldens bounds it at ~1e-14)"""
    from tests.recipes_input import recipe_trajectory
    g = golden("recipes.npz")
    traj = recipe_trajectory()
    xs, ys, zs, lx, ly, lz = orc.grid_from_box(traj.box_matrix)
    V = orc.potential(xs, ys, zs, traj.positions[0], traj.atom_types, workers=4)
    assert V.shape == (154, 171, 14)
    assert _dz(V[::3, ::3, :], g["r01_potential"]) < 1e-12
    probe = orc.probe_array(xs, ys, 5, 100e3)
    exit_wave = orc.propagate(probe, V, xs, ys, zs, 100e3, workers=4)[0]
    assert _dz(exit_wave[::2, ::2], g["r02_exit"]) < 1e-12
    assert np.linalg.norm(exit_wave[::2, ::2] - g["r02_exit"]) / np.linalg.norm(g["r02_exit"]) < 1e-6
    wf, _ = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=0.0, voltage_eV=100e3, workers=4)
    inten, freqs = orc.tacaw_intensity(wf[..., 0], np.arange(traj.n_frames) * traj.timestep)
    assert np.allclose(freqs, g["r05_frequencies"], rtol=0, atol=1e-9)
    assert _dz(inten[0, 7] ** .1, g["r05_slice"]) < 1e-10
    keep = [i for i in range(12) if i != 6]
    assert np.abs(orc.spectrum(inten)[keep] / g["r05_spectrum"][keep] - 1).max() < 1e-9


def test_oracle_replays_reference_recipe_04_haadf():
    """04_haadf.py: cropped trajectory, 3 shuffled frames, 14 x 16 probe grid at 30 mrad, calculateADF"""
    from pyslice_b200.multislice.multislice import probe_grid
    from tests.recipes_input import A, B, recipe_trajectory
    g = golden("recipes.npz")
    cut = recipe_trajectory().slice_positions([0, 4 * A], [0, 3 * B])
    three = cut.slice_timesteps(g["r04_frames"])
    xy = probe_grid([A, 3 * A], [B, 2 * B], 14, 16)
    wf, grid = orc.multislice_run(three.positions, three.atom_types, three.box_matrix, aperture=30.0, voltage_eV=100e3,
                                  probe_positions=[tuple(p) for p in xy], workers=4)
    kxs, kys, _ = orc.wf_axes(wf.shape[2], wf.shape[3], 0.1, three.n_frames, three.timestep)
    adf, ux, uy = orc.haadf_adf(wf, kxs, kys, xy, 100e3)
    assert adf.shape == (14, 16)
    assert np.abs(adf / g["r04_adf"] - 1).max() < 1e-5
