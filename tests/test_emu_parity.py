"""CPU parity tests of the kernel SOURCE through the emulator (tests/emu): the CUDA kernels of
pyslice_b200/csrc compiled by g++ with threads emulated, driven through the same Python host code
and C ABI as on the GPU, compared with the reference goldens.  Tolerances as on the GPU:
binning exact, exit wave rel-L2 <= 1e-4, TACAW intensity <= 1e-3."""
import numpy as np
import pytest
import torch

from oracle import pyslice_oracle as orc
from tests import emu
from tests.helpers import golden, rel_l2, small64_traj, tacaw48_traj


@pytest.fixture(scope="module", autouse=True)
def emulator():
    if torch.cuda.is_available():
        pytest.skip("emulator tests are for CPU-only boxes; the GPU suite covers the real library")
    from pyslice_b200 import _lib
    emu.activate()
    yield
    _lib._reset()


def membership(plan, offsets, atom_list, n_atoms):
    off, al = offsets.numpy(), atom_list.numpy()
    m = np.zeros((plan.nz, n_atoms), dtype=bool)
    for s in range(plan.nz):
        for t in range(plan.ntypes):
            seg = s * plan.ntypes + t
            ids = al[off[seg]:off[seg + 1]]
            assert np.all(np.diff(ids) > 0), "segment must be sorted by atom index"
            m[s, ids] = True
    return m


def test_binning_of_crowded_segments():
    """(slice, type) segments with hundreds of atoms (a 2-D material, a thick slice): the CTA-wide ranking of
    BinOrderLarge orders them like the one-warp kernel orders small ones -- ascending atom index, memberships exact"""
    from pyslice_b200 import engine, hostmath, synthetic
    traj = synthetic.random_trajectory(n_atoms=700, box=(1.55, 1.55, 1.2), n_frames=2, seed=19, types=(6,), stray=True)
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    offsets, atom_list, ux, uy = engine.bin_atoms(plan, torch.from_numpy(traj.positions.copy()))
    assert int((offsets[0, 1:] - offsets[0, :-1]).max()) > 64
    for f in range(2):
        want = orc.bin_atoms(traj.positions[f][:, 2], zs)
        assert np.array_equal(membership(plan, offsets[f], atom_list[f], traj.n_atoms), want)
        n = int(offsets[f, -1])
        ids = atom_list[f, :n].numpy()
        u = traj.positions[f][ids, 0] / (plan.nx * plan.dx)
        want_u = ((np.floor((u - np.floor(u)) * 4294967296.0 + 0.5)).astype(np.uint64) & 0xffffffff).astype(np.uint32)
        assert np.array_equal(ux[f, :n].numpy().view(np.uint32), want_u)


@pytest.mark.parametrize("shape", [(16, 16), (32, 64), (48, 40), (20, 36), (128, 16)])
def test_fft2_kernels(shape):
    from pyslice_b200 import engine
    rng = np.random.default_rng(1)
    x = (rng.normal(size=(2,) + shape) + 1j * rng.normal(size=(2,) + shape)).astype(np.complex64)
    y = engine.fft2(torch.from_numpy(x)).numpy()
    assert rel_l2(y, np.fft.fft2(x.astype(np.complex128))) < 1e-6
    yi = engine.fft2(torch.from_numpy(x), inverse=True, scale=1.0 / (shape[0] * shape[1])).numpy()
    assert rel_l2(yi, np.fft.ifft2(x.astype(np.complex128))) < 1e-6


def test_binning_bit_exact_and_potential():
    from pyslice_b200 import engine, hostmath
    traj = small64_traj()
    g = golden("small64.npz")
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    pos = torch.from_numpy(traj.positions[:2].copy())
    offsets, atom_list, _, _ = engine.bin_atoms(plan, pos)
    for f in range(2):
        want = orc.bin_atoms(traj.positions[f][:, 2], zs)
        assert np.array_equal(membership(plan, offsets[f], atom_list[f], traj.n_atoms), want)
    t, V = engine.build_transmission(plan, pos[:1], want_potential=True)
    assert rel_l2(V[0].permute(1, 2, 0).numpy(), g["potential0"]) < 1e-5
    assert np.allclose(np.abs(t.numpy()), 1.0, atol=1e-6)


def test_plane_wave_run_and_probe_run():
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.haadf_data import HAADFData
    traj = small64_traj()
    g = golden("small64.npz")
    calc = MultisliceCalculator()
    calc.setup(traj.slice_timesteps([0, 1]), aperture=0.0, voltage_eV=100e3)
    wf = calc.run()
    for f in range(2):
        assert rel_l2(wf.wavefunction_data[0, f, :, :, 0].numpy(), g["wf_plane"][0, f, :, :, 0]) < 1e-4
    assert np.array_equal(wf.kxs.numpy(), g["kxs"])
    calc.setup(traj.slice_timesteps([0]), aperture=30.0, voltage_eV=100e3, probe_positions=g["probe_xy"])
    assert rel_l2(calc.base_probe.array.numpy(), g["base_probe"]) < 1e-5
    wf2 = calc.run()
    for p in range(4):
        assert rel_l2(wf2.wavefunction_data[p, 0, :, :, 0].numpy(), g["wf_probes"][p, 0, :, :, 0]) < 1e-4
    adf = HAADFData(wf2).calculateADF(45)
    assert adf.shape == (2, 2)


def test_tacaw_non_power_of_two():
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    traj = tacaw48_traj()
    g = golden("tacaw48.npz")
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    wf = calc.run()
    assert rel_l2(wf.wavefunction_data.numpy(), g["wf"]) < 1e-4
    tac = TACAWData(wf)
    inten = tac.intensity.numpy()
    dc = inten.shape[1] // 2
    keep = [i for i in range(inten.shape[1]) if i != dc]
    assert np.allclose(tac.frequencies, g["frequencies"])
    assert rel_l2(inten[:, keep], g["intensity"][:, keep]) < 1e-3
    assert inten[:, dc].max() < 1e-9 * inten.max()
    assert rel_l2(tac.spectrum()[keep], g["spectrum"][keep]) < 1e-3
    assert rel_l2(tac.diffraction(), g["diffraction"]) < 1e-3
    assert rel_l2(tac.spectral_diffraction(20.0), g["spectral_diffraction"]) < 1e-3
    assert rel_l2(tac.spectrum_image(20.0), g["spectrum_image"]) < 1e-3
    assert rel_l2(tac.dispersion(g["kx_path"], g["ky_path"])[keep], g["dispersion"][keep]) < 1e-3
    with pytest.raises(ValueError):
        TACAWData(wf, layer_index=3)
    with pytest.raises(ValueError):
        tac.spectrum(5)


def test_layers_and_propagate_api():
    """Layer-resolved output against the truncated-stack oracle and the low-level
    Potential / Propagate API against the oracle's real-space exit wave."""
    from pyslice_b200 import hostmath
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.multislice.multislice import Probe, Propagate
    from pyslice_b200.multislice.potentials import Potential
    traj = small64_traj().slice_timesteps([0])
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3, layer_every=4)
    wf = calc.run()
    assert list(wf.layer) == [3, 7, 8] and wf.wavefunction_data.shape[-1] == 3
    V = orc.potential(xs, ys, zs, traj.positions[0], traj.atom_types)
    for li, last in enumerate([3, 7, 8]):
        ref = orc.exit_to_kspace(orc.propagate(np.ones((64, 64)), V, xs, ys, zs, 100e3, n_slices=last + 1))[0]
        assert rel_l2(wf.wavefunction_data[0, 0, :, :, li].numpy(), ref) < 1e-4
    pot = Potential(xs, ys, zs, traj.positions[0], traj.atom_types.tolist())
    assert tuple(pot.array.shape) == (64, 64, 9)
    probe = Probe(xs, ys, 0.0, 100e3)
    psi = Propagate(probe, pot)
    ref = orc.propagate(np.ones((64, 64)), V, xs, ys, zs, 100e3)[0]
    assert rel_l2(psi.numpy(), ref) < 1e-4


def test_potential_chunking_is_bitwise_invariant():
    """the potential build streams slice-pair images through an L2-sized scratch; any chunking
    (pairs split within a frame, several frames per chunk) must give identical bits"""
    from pyslice_b200 import engine, hostmath
    traj = small64_traj()
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    pos = torch.from_numpy(traj.positions.copy())
    ref = engine.build_transmission(plan, pos)
    old = engine.SCRATCH_BYTES
    try:
        for nimg in (1, 3, 11):
            engine.SCRATCH_BYTES = 64 * 64 * 8 * nimg
            assert torch.equal(engine.build_transmission(plan, pos), ref)
    finally:
        engine.SCRATCH_BYTES = old
