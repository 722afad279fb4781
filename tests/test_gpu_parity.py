"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path through the C ABI against
the CPU oracle and the reference goldens.

Tolerances (BASELINE.json north_star): atom->slice binning bit-exact; exit wave relative L2 <= 1e-4
per (probe, frame); TACAW intensity <= 1e-3 relative; potential <= 1e-5 (own budget)."""
import numpy as np
import pytest
import torch

from oracle import pyslice_oracle as orc
from tests.helpers import golden, rel_l2, si_c1_traj, small64_traj, tacaw48_traj

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def real_library():
    from pyslice_b200 import _lib
    _lib._reset()
    assert not _lib.is_emulated()
    yield


def dev(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def membership(plan, offsets, atom_list, n_atoms):
    off, al = offsets.cpu().numpy(), atom_list.cpu().numpy()
    m = np.zeros((plan.nz, n_atoms), dtype=bool)
    for s in range(plan.nz):
        for t in range(plan.ntypes):
            seg = s * plan.ntypes + t
            ids = al[off[seg]:off[seg + 1]]
            assert np.all(np.diff(ids) > 0)
            m[s, ids] = True
    return m


FFT_SHAPES = [(16, 16), (32, 64), (64, 64), (128, 256), (256, 256), (512, 512), (1024, 1024), (2048, 64),
              (64, 4096), (8192, 16), (48, 40), (20, 36), (272, 272), (100, 500), (1000, 24), (24, 2000), (4000, 16),
              (1, 64), (64, 1), (7, 9)]


@pytest.mark.parametrize("shape", FFT_SHAPES)
def test_fft2_every_size(shape):
    from pyslice_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((3,) + shape, dtype=torch.complex64, device="cuda", generator=g)
    ref = torch.fft.fft2(x.to(torch.complex128))
    got = engine.fft2(x)
    assert float((got - ref).norm() / ref.norm()) < 2e-6
    refi = torch.fft.ifft2(x.to(torch.complex128))
    goti = engine.fft2(x, inverse=True, scale=1.0 / (shape[0] * shape[1]))
    assert float((goti - refi).norm() / refi.norm()) < 2e-6


def test_binning_bit_exact_small_and_large():
    from pyslice_b200 import engine, hostmath, synthetic
    traj = small64_traj()
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    offsets, atom_list, _, _ = engine.bin_atoms(plan, dev(traj.positions))
    for f in range(traj.n_frames):
        assert np.array_equal(membership(plan, offsets[f], atom_list[f], traj.n_atoms),
                              orc.bin_atoms(traj.positions[f][:, 2], zs))
    # 611 slices with 1-ulp gaps/overlaps (lz = 305.3), 20k atoms incl. atoms on every slice bound
    big = synthetic.random_trajectory(n_atoms=20000, box=(6.35, 6.35, 305.3), n_frames=2, seed=9, stray=True)
    xs, ys, zs, *_ = hostmath.grid_from_box(big.box_matrix)
    plan = engine.make_plan(xs, ys, zs, big.atom_types.tolist(), 100e3)
    offsets, atom_list, _, _ = engine.bin_atoms(plan, dev(big.positions))
    for f in range(2):
        want = orc.bin_atoms(big.positions[f][:, 2], zs)
        got = membership(plan, offsets[f], atom_list[f], big.n_atoms)
        assert np.array_equal(got, want)
        assert (want.sum(axis=0) == 2).any() or (want.sum(axis=0) == 0).any()


def test_potential_against_reference_golden():
    from pyslice_b200 import engine, hostmath
    traj = small64_traj()
    g = golden("small64.npz")
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    t, V = engine.build_transmission(plan, dev(traj.positions[:1]), want_potential=True)
    assert rel_l2(V[0].permute(1, 2, 0).cpu().numpy(), g["potential0"]) < 1e-5
    assert float((t.abs() - 1).abs().max()) < 1e-6            # unit-modulus transmission


def test_small64_runs_against_reference_golden():
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.haadf_data import HAADFData
    traj = small64_traj()
    g = golden("small64.npz")
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    wf = calc.run()
    assert tuple(wf.wavefunction_data.shape) == g["wf_plane"].shape
    for f in range(3):
        assert rel_l2(wf.wavefunction_data[0, f, :, :, 0].cpu().numpy(), g["wf_plane"][0, f, :, :, 0]) < 1e-4
    assert np.array_equal(wf.kxs.numpy(), g["kxs"]) and np.array_equal(wf.time, g["time"])
    calc.setup(traj, aperture=30.0, voltage_eV=100e3, probe_positions=g["probe_xy"])
    assert rel_l2(calc.base_probe.array.cpu().numpy(), g["base_probe"]) < 1e-5
    wf2 = calc.run()
    for p in range(4):
        for f in range(3):
            assert rel_l2(wf2.wavefunction_data[p, f, :, :, 0].cpu().numpy(), g["wf_probes"][p, f, :, :, 0]) < 1e-4
    adf = HAADFData(wf2).calculateADF(45)
    assert rel_l2(adf.numpy(), g["adf"]) < 1e-4
    # detector-only run: the same image without ever allocating the exit-wave cube (sums taken after each exit FFT)
    calc.setup(traj, aperture=30.0, voltage_eV=100e3, probe_positions=g["probe_xy"], adf_collection_angle=45)
    wf3 = calc.run()
    assert wf3.wavefunction_data is None and tuple(wf3.adf_sums.shape) == (1, 4, 3)
    adf3 = HAADFData(wf3).calculateADF(45)
    assert rel_l2(adf3.numpy(), g["adf"]) < 1e-4 and rel_l2(adf3.numpy(), adf.numpy()) < 1e-6
    with pytest.raises(ValueError):
        HAADFData(wf3).calculateADF(30)


def test_tacaw48_against_reference_golden():
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    traj = tacaw48_traj()
    g = golden("tacaw48.npz")
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    wf = calc.run()
    assert rel_l2(wf.wavefunction_data.cpu().numpy(), g["wf"]) < 1e-4
    tac = TACAWData(wf)
    inten = tac.intensity.cpu().numpy()
    dc = inten.shape[1] // 2
    keep = [i for i in range(inten.shape[1]) if i != dc]
    assert rel_l2(inten[:, keep], g["intensity"][:, keep]) < 1e-3
    assert inten[:, dc].max() < 1e-9 * inten.max()             # reference DC bin ~1e-19: absolute check
    assert rel_l2(tac.spectrum()[keep], g["spectrum"][keep]) < 1e-3
    assert rel_l2(tac.spectrum(0)[keep], g["spectrum0"][keep]) < 1e-3
    assert rel_l2(tac.diffraction(), g["diffraction"]) < 1e-3
    assert rel_l2(tac.spectral_diffraction(20.0), g["spectral_diffraction"]) < 1e-3
    assert rel_l2(tac.spectrum_image(20.0), g["spectrum_image"]) < 1e-3
    assert rel_l2(tac.dispersion(g["kx_path"], g["ky_path"])[keep], g["dispersion"][keep]) < 1e-3
    ref_masked = (g["intensity"] * g["mask"][None, None]).sum(axis=(2, 3)).mean(axis=0)
    assert rel_l2(tac.masked_spectrum(g["mask"])[keep], ref_masked[keep]) < 1e-3


def test_si_c1_against_reference_golden():
    """config C1 geometry: 256 x 256 x 103, 2000 Si atoms, plane wave."""
    from pyslice_b200 import engine
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    g = golden("si_c1.npz")
    traj = si_c1_traj()
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    assert (calc.nx, calc.ny, calc.nz) == (256, 256, 103)
    wf = calc.run()
    assert rel_l2(wf.wavefunction_data[0, 0, :, :, 0].cpu().numpy(), g["wf"]) < 1e-4
    _, V = engine.build_transmission(calc._plan, dev(traj.positions[:1]), want_potential=True)
    V = V[0].permute(1, 2, 0).cpu().numpy()
    assert rel_l2(V[:, :, [0, 1, 51, 102]], g["pot_planes"]) < 1e-5
    assert rel_l2(V.sum(axis=2), g["pot_sum_z"]) < 1e-5


@pytest.mark.parametrize("aperture", [0.0, 30.0])
def test_long_stack_611_slices_vs_oracle(aperture):
    """fp32 round-off accumulates over slices: check at nz = 611."""
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.random_trajectory(n_atoms=6000, box=(12.75, 12.75, 305.3), n_frames=1, seed=21, types=(14,))
    pp = None if aperture == 0 else [(3.0, 4.0), (7.7, 9.1)]
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=aperture, voltage_eV=100e3, probe_positions=pp)
    assert (calc.nx, calc.nz) == (128, 611)
    wf = calc.run().wavefunction_data.cpu().numpy()
    ref, _ = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=aperture,
                                voltage_eV=100e3, probe_positions=pp, workers=4)
    for p in range(wf.shape[0]):
        assert rel_l2(wf[p, 0], ref[p, 0]) < 1e-4


def test_non_power_of_two_grid_272_vs_oracle():
    """natural Si lattice constant -> 272 x 272 grid (Bluestein path), config C1 variant."""
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.silicon_trajectory(cells=(5, 5, 4), a=5.431, n_frames=2, seed=4)
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    assert (calc.nx, calc.ny) == (272, 272)
    wf = calc.run().wavefunction_data.cpu().numpy()
    ref, _ = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, voltage_eV=100e3, workers=4)
    for f in range(2):
        assert rel_l2(wf[0, f], ref[0, f]) < 1e-4


def test_hbn_multi_type_probes_vs_oracle():
    """three element types, empty slices, convergent probes (config C3 physics at 128 x 128)."""
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.multislice.multislice import probe_grid
    traj = synthetic.hbn_graphene_trajectory(cells=(5, 3), n_layers=4, n_frames=2, seed=2)
    xy = probe_grid([2, 9], [3, 10], 3, 2)
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=30.0, voltage_eV=100e3, probe_positions=xy)
    wf = calc.run().wavefunction_data.cpu().numpy()
    ref, _ = orc.multislice_run(traj.positions, traj.atom_types, traj.box_matrix, aperture=30.0, voltage_eV=100e3,
                                probe_positions=xy, workers=4)
    assert wf.shape == ref.shape
    for p in range(6):
        for f in range(2):
            assert rel_l2(wf[p, f], ref[p, f]) < 1e-4


@pytest.mark.parametrize("T", [20, 100, 500, 2000, 4000, 64, 97, 14, 331, 45, 6, 2, 40, 60, 200, 250, 300, 1000])
def test_tacaw_time_fft_lengths(T):
    # 2^a 3^b 5^c lengths run the tiled mixed-radix kernel (tacaw_fast.cu; 40 ... 1000: plans mixing the prime-factor radices
    # 10 and 20 with 2, 3, 4, 5), 97 / 14 / 331 the Bluestein line pass
    from pyslice_b200 import engine
    rng = np.random.default_rng(T)
    P, nx, ny = 2, 8, 16
    base = rng.normal(size=(P, 1, nx, ny)) + 1j * rng.normal(size=(P, 1, nx, ny))
    x = (5.0 * base + 0.1 * (rng.normal(size=(P, T, nx, ny)) + 1j * rng.normal(size=(P, T, nx, ny)))).astype(np.complex64)
    got = engine.tacaw_intensity(dev(x)).cpu().numpy()
    ref, _ = orc.tacaw_intensity(x.astype(np.complex128), np.arange(T) * 0.01)
    dc = T // 2
    keep = [i for i in range(T) if i != dc]
    assert rel_l2(got[:, keep], ref[:, keep]) < 1e-3
    # Parseval: sum_w I = T * sum_t |psi - mean|^2
    d = x.astype(np.complex128) - x.astype(np.complex128).mean(axis=1, keepdims=True)
    assert abs(got.sum() / (T * (np.abs(d) ** 2).sum()) - 1) < 1e-4


def test_full_size_properties_c2_grid():
    """Size-independent properties at the benchmark grid (256 x 256 x 512, 10k atoms, 2 frames):
    |t| = 1, the propagator is unitary so sum|wf|^2 = nx*ny*sum|probe|^2, and runs are deterministic."""
    from pyslice_b200 import engine, synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.silicon_trajectory(cells=(5, 5, 50), a=5.11, n_frames=2, seed=1, displacement="phonon")
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    assert (calc.nx, calc.ny, calc.nz) == (256, 256, 512)
    wf1 = calc.run().wavefunction_data.clone()
    wf2 = calc.run().wavefunction_data
    assert torch.equal(wf1, wf2)
    power = (wf1.abs().double() ** 2).sum(dim=(2, 3, 4)) / (256 * 256)
    assert float((power / (256 * 256) - 1).abs().max()) < 1e-4
    # batching must not change a single bit
    old = engine.PSI_BATCH_BYTES
    try:
        engine.PSI_BATCH_BYTES = 600 << 10
        wf3 = calc.run().wavefunction_data
    finally:
        engine.PSI_BATCH_BYTES = old
    assert torch.equal(wf1, wf3)
    # one frame against the oracle at full depth
    ref, _ = orc.multislice_run(traj.positions[:1], traj.atom_types, traj.box_matrix, voltage_eV=100e3, workers=8)
    assert rel_l2(wf1[0, 0, :, :, 0].cpu().numpy(), ref[0, 0, :, :, 0]) < 1e-4


def test_layers_and_low_level_api():
    from pyslice_b200 import hostmath
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.multislice.multislice import Probe, Propagate, create_batched_probes
    from pyslice_b200.multislice.potentials import Potential
    traj = small64_traj().slice_timesteps([0])
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3, layer_every=4)
    wf = calc.run()
    assert list(wf.layer) == [3, 7, 8]
    V = orc.potential(xs, ys, zs, traj.positions[0], traj.atom_types)
    for li, last in enumerate([3, 7, 8]):
        ref = orc.exit_to_kspace(orc.propagate(np.ones((64, 64)), V, xs, ys, zs, 100e3, n_slices=last + 1))[0]
        assert rel_l2(wf.wavefunction_data[0, 0, :, :, li].cpu().numpy(), ref) < 1e-4
    pot = Potential(xs, ys, zs, traj.positions[0], traj.atom_types.tolist())
    assert rel_l2(pot.array.cpu().numpy(), V) < 1e-5
    base = Probe(xs, ys, 30.0, 100e3)
    probes = create_batched_probes(base, [(1.0, 2.0), (3.3, 4.4)])
    want = orc.shifted_probes(orc.probe_array(xs, ys, 30.0, 100e3), xs, ys, [(1.0, 2.0), (3.3, 4.4)])
    assert rel_l2(probes.array.cpu().numpy(), want) < 1e-5
    psi = Propagate(probes, pot)
    ref = orc.propagate(want, V, xs, ys, zs, 100e3)
    assert rel_l2(psi.cpu().numpy(), ref) < 1e-4
    # defocus (reference multislice.py:183-190)
    p2 = Probe(xs, ys, 30.0, 100e3)
    p2.defocus(200.0)
    k = np.fft.fftfreq(64, xs[1] - xs[0])
    Pd = np.exp(-1j * np.pi * orc.wavelength(100e3) * 200.0 * (k[:, None] ** 2 + k[None, :] ** 2))
    wantd = np.fft.ifft2(Pd * np.fft.fft2(orc.probe_array(xs, ys, 30.0, 100e3)))
    assert rel_l2(p2.array.cpu().numpy(), wantd) < 1e-5


FAST_CASES = [
    # box (A), probes, aperture  ->  grid; all four (column length, row length) kernel instantiations
    ((25.55, 25.55, 6.1), 3, 30.0, (256, 256)),
    ((51.15, 51.15, 3.1), 2, 30.0, (512, 512)),
    ((25.55, 51.15, 4.1), 1, 0.0, (256, 512)),
    ((51.15, 25.55, 4.1), 2, 20.0, (512, 256)),
    # 1024-point lines (warp-pair row pipelines, radix 16 x 4 x 16 with the in-place middle stage, 8-column tiles)
    ((102.35, 102.35, 2.6), 2, 30.0, (1024, 1024)),
    ((102.35, 25.55, 3.1), 1, 0.0, (1024, 256)),
    ((25.55, 102.35, 3.1), 2, 20.0, (256, 1024)),
    ((51.15, 102.35, 2.1), 1, 0.0, (512, 1024)),
]


@pytest.mark.parametrize("box,n_probes,aperture,grid", FAST_CASES)
def test_fused_slice_step_vs_generic_and_oracle(box, n_probes, aperture, grid):
    """The persistent TMA/packed-fp32 slice-step kernels (fast_path.cu) against the generic line passes
    (same inputs, both CUDA) and against the oracle, for every supported (nx, ny) pairing; 5 frames x
    n_probes images so that tiles wrap over CTAs/warps more than once in the persistent loops."""
    from pyslice_b200 import engine, synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    rng = np.random.default_rng(5)
    traj = synthetic.random_trajectory(n_atoms=600, box=box, n_frames=5, seed=31, types=(6, 14))
    pp = None if n_probes == 1 and aperture == 0 else [(float(rng.uniform(1, box[0] - 1)), float(rng.uniform(1, box[1] - 1)))
                                                        for _ in range(n_probes)]
    out = {}
    for fast in (True, False):
        engine.set_fast_path(fast)
        try:
            calc = MultisliceCalculator()
            calc.setup(traj, aperture=aperture, voltage_eV=100e3, probe_positions=pp)
            assert (calc.nx, calc.ny) == grid
            out[fast] = calc.run().wavefunction_data.cpu().numpy()
        finally:
            engine.set_fast_path(True)
    assert rel_l2(out[True], out[False]) < 5e-6
    ref, _ = orc.multislice_run(traj.positions[:2], traj.atom_types, traj.box_matrix, aperture=aperture, voltage_eV=100e3,
                                probe_positions=pp, workers=4)
    for p in range(ref.shape[0]):
        for f in range(2):
            assert rel_l2(out[True][p, f], ref[p, f]) < 1e-4


@pytest.mark.parametrize("box,n_atoms,types,n_frames", [
    ((25.55, 25.55, 9.3), 900, (6, 14, 31), 3),   # 256 x 256, 19 slices (odd: last pair is half empty), 3 types
    ((51.15, 25.55, 3.2), 500, (14,), 3),         # 512 x 256
    ((25.55, 51.15, 2.2), 700, (5, 7), 2),        # 256 x 512
    ((51.15, 51.15, 1.2), 400, (6,), 2),          # 512 x 512
    ((102.35, 102.35, 1.7), 1500, (14, 6), 2),    # 1024 x 1024 (3 slices: last pair half empty)
    ((102.35, 51.15, 1.2), 600, (14,), 2),        # 1024 x 512
    ((25.55, 102.35, 1.2), 500, (5, 7), 2),       # 256 x 1024
    ((25.55, 25.55, 1.6), 6000, (14, 79), 2),     # 256 x 256, ~1500 atoms per slice: segments of many ring blocks
    ((25.55, 25.55, 40.3), 300, (6, 14), 4),      # 256 x 256, 81 slices, ~2 atoms per (slice, type): empty segments
    ((6.35, 6.35, 4.1), 200, (6, 14, 31), 3),     # 64 x 64: pipelined structure factor + generic transforms
    ((4.75, 3.95, 3.2), 150, (5, 7), 3),          # 48 x 40 (odd half sizes, Bluestein transforms)
    ((2.45, 2.45, 30.2), 2500, (14,), 3),         # 25 x 25 odd grid, ~40 atoms per slice and > 32 in many (multi-block segments)
])
def test_potential_pipelined_vs_generic_and_oracle(box, n_atoms, types, n_frames):
    """psb_build_transmission through both kernel generations -- 1: pipelined structure factor (sf_fast.cu) + fused
    inverse transforms, 0: generic kernels -- on the same inputs, and against the oracle's potential."""
    from pyslice_b200 import engine, hostmath, synthetic
    traj = synthetic.random_trajectory(n_atoms=n_atoms, box=box, n_frames=n_frames, seed=13, types=types, stray=True)
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    pos = dev(traj.positions)
    out = {}
    for level in (1, 0):
        engine.set_fast_path(level)
        try:
            t, V = engine.build_transmission(plan, pos, want_potential=True)
            out[level] = (t.cpu().numpy(), V.cpu().numpy())
        finally:
            engine.set_fast_path(True)
    for level in (1,):
        assert rel_l2(out[level][1], out[0][1]) < 2e-6, level
        assert rel_l2(out[level][0], out[0][0]) < 2e-6, level
    Vref = orc.potential(xs, ys, zs, traj.positions[1], traj.atom_types)      # (nx, ny, nz)
    for level in (1,):
        err = rel_l2(np.moveaxis(out[level][1][1], 0, 2), Vref)
        print(f"potential rel-L2 vs oracle, level {level}: {err:.3e}")
        assert err < 1e-5, level
    assert np.allclose(np.abs(out[1][0]), 1.0, atol=1e-6)


def test_potential_small_scratch_chunks():
    """chunks smaller than a frame's pair count (several launches per frame, partial last chunk) through the fused
    structure-factor + column kernel equal the one-chunk result"""
    from pyslice_b200 import engine, hostmath, synthetic
    traj = synthetic.random_trajectory(n_atoms=800, box=(25.55, 25.55, 12.3), n_frames=2, seed=21, types=(6, 14), stray=True)
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    pos = dev(traj.positions)
    t_ref = engine.build_transmission(plan, pos)
    keep = engine.SCRATCH_BYTES
    try:
        engine.SCRATCH_BYTES = 5 * 256 * 256 * 8          # 5 pair images per chunk; 25 slices = 13 pairs per frame
        t_small = engine.build_transmission(plan, pos)
    finally:
        engine.SCRATCH_BYTES = keep
    assert torch.equal(t_ref, t_small)


def test_fused_slice_step_many_images_ragged():
    """image counts that do not divide the persistent grids (148 SMs x 16 warps / x 2 CTAs): 7 frames x 3 probes
    at 256 x 256, every (probe, frame) equal to the generic path within round-off"""
    from pyslice_b200 import engine, synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.random_trajectory(n_atoms=300, box=(25.55, 25.55, 2.1), n_frames=7, seed=8, types=(14,))
    pp = [(3.0, 4.0), (12.2, 20.1), (21.0, 7.7)]
    out = {}
    for fast in (True, False):
        engine.set_fast_path(fast)
        try:
            calc = MultisliceCalculator()
            calc.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp)
            out[fast] = calc.run().wavefunction_data.cpu().numpy()
        finally:
            engine.set_fast_path(True)
    for p in range(3):
        for f in range(7):
            assert rel_l2(out[True][p, f], out[False][p, f]) < 5e-6


def test_probe_kat():
    """the reference's own probe recipe (src/unittests/00_probe.py:7-18), 501 x 491 grid"""
    from pyslice_b200.multislice.multislice import Probe
    g = golden("probe_kat.npz")
    xs = np.linspace(0, 50, 501)
    ys = np.linspace(0, 49, 491)
    for mrad in (1, 5, 30):
        p = Probe(xs, ys, mrad, 100e3).array.cpu().numpy()[::5, ::5]
        assert rel_l2(p, g[f"mrad{mrad}"]) < 1e-5


def test_errors_are_loud():
    from pyslice_b200 import _lib, engine
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    with pytest.raises(RuntimeError):
        MultisliceCalculator(force_cpu=True)
    x = torch.zeros((1, 3, 9000), dtype=torch.complex64, device="cuda")
    with pytest.raises(_lib.PsbError):
        engine.fft2(x)                      # 9000 > 4096 and not a power of two -> unsupported


def test_frame_cache_on_device(tmp_path):
    """SURVEY 8f-3 on the CUDA path: files in the reference's wire format, reload equals the computed run"""
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.random_trajectory(n_atoms=120, box=(6.35, 6.35, 3.1), n_frames=5, seed=4, types=(6, 14))
    pp = [(1.0, 2.0), (4.0, 5.0), (3.0, 3.0)]
    c = MultisliceCalculator()
    c.setup(traj, aperture=20.0, voltage_eV=100e3, probe_positions=pp, frame_cache=tmp_path)
    wf = c.run().wavefunction_data.clone()
    assert (c.frames_computed, c.frames_cached) == (5, 0)
    a = np.load(c.output_dir / "frame_3.npy")
    assert a.shape == (3, 64, 64, 1, 1) and a.dtype == np.complex128
    assert np.array_equal(a[:, :, :, 0, 0].astype(np.complex64), wf[:, 3, :, :, 0].cpu().numpy())
    (c.output_dir / "frame_1.npy").unlink()
    d = MultisliceCalculator()
    d.setup(traj, aperture=20.0, voltage_eV=100e3, probe_positions=pp, frame_cache=tmp_path)
    wf2 = d.run().wavefunction_data
    assert (d.frames_computed, d.frames_cached) == (1, 4)
    assert torch.equal(wf, wf2)


def test_linearity_in_the_probe_and_time_parseval():
    """Properties at a fused-kernel grid (256 x 256, 40 slices): Propagate is linear in the probe, and the TACAW
    cube obeys Parseval along time: sum_w I(w, k) = T * sum_t |psi_t(k) - <psi(k)>|^2."""
    from pyslice_b200 import engine, hostmath, synthetic
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.random_trajectory(n_atoms=1500, box=(25.55, 25.55, 20.1), n_frames=6, seed=17, types=(14, 31))
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    t = engine.build_transmission(plan, dev(traj.positions[:2]))
    g = torch.Generator(device="cuda").manual_seed(3)
    p = torch.randn((2, 256, 256), generator=g, device="cuda", dtype=torch.float32) + 0j
    p = p.to(torch.complex64)
    both = torch.stack([p[0], p[1], (0.3 - 1.1j) * p[0] + (2.0 + 0.5j) * p[1]])
    out = engine.propagate(plan, both, t)                     # (F, P, nx, ny) real-space exit waves
    lin = (0.3 - 1.1j) * out[:, 0] + (2.0 + 0.5j) * out[:, 1]
    assert rel_l2(out[:, 2].cpu().numpy(), lin.cpu().numpy()) < 2e-6

    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3)
    wf = calc.run()
    tac = TACAWData(wf)
    psi = wf.wavefunction_data[0, :, :, :, 0].to(torch.complex128)
    want = psi.shape[0] * ((psi - psi.mean(dim=0, keepdim=True)).abs() ** 2).sum(dim=0)
    got = tac.intensity[0].double().sum(dim=0)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-5


@pytest.mark.parametrize("config", ["c3", "c4"])
def test_full_size_properties_other_grids(config):
    """Unitarity and determinism at the other benchmark grids: C3 512 x 512 x 67 (9 600 atoms, three types, 4 convergent
    probes, fused kernels) and C4 1024 x 1024 x 123 (38 400 atoms, plane wave, generic line passes), one frame each."""
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.multislice.multislice import probe_grid
    if config == "c3":
        traj = synthetic.hbn_graphene_trajectory(n_frames=1, seed=2)
        pp, ap, grid = [tuple(q) for q in probe_grid([10, 40], [12, 38], 2, 2)], 30.0, (512, 512, 67)
    else:
        traj = synthetic.silicon_trajectory(cells=(20, 20, 12), a=5.1175, n_frames=1, seed=3, displacement="phonon")
        pp, ap, grid = None, 0.0, (1024, 1024, 123)
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=ap, voltage_eV=100e3, probe_positions=pp)
    assert (calc.nx, calc.ny, calc.nz) == grid
    wf1 = calc.run().wavefunction_data.clone()
    wf2 = calc.run().wavefunction_data
    assert torch.equal(wf1, wf2)
    n = grid[0] * grid[1]
    p0 = (calc._probes.abs().double() ** 2).sum(dim=(1, 2))                     # real-space probe power
    power = (wf1.abs().double() ** 2).sum(dim=(2, 3, 4))[:, 0] / n             # Parseval: sum|FFT|^2 = n * sum|psi|^2
    assert float((power / p0 - 1).abs().max()) < 1e-4


@pytest.mark.parametrize("T,nx,ny,P", [(100, 33, 7, 3), (500, 16, 24, 1), (20, 5, 5, 2), (2000, 4, 10, 1), (4000, 3, 3, 1)])
def test_tacaw_tiled_kernel_vs_generic(T, nx, ny, P):
    """tacaw_fast.cu against the generic line pass on the same input: pixel counts that do not fill the last tile, and
    a layer view of a (L, P, T, nx, ny) store (frame stride != pixels per image is not needed, probe stride is)."""
    from pyslice_b200 import engine
    rng = np.random.default_rng(T + nx)
    store = torch.from_numpy((rng.normal(size=(2, P, T, nx, ny)) + 1j * rng.normal(size=(2, P, T, nx, ny))
                              + 3.0).astype(np.complex64)).cuda()
    x = store[1]
    out = {}
    for level in (1, 0):
        engine.set_fast_path(level)
        try:
            out[level] = engine.tacaw_intensity(x).cpu().numpy()
        finally:
            engine.set_fast_path(True)
    dc = T // 2
    keep = [i for i in range(T) if i != dc]
    assert rel_l2(out[1][:, keep], out[0][:, keep]) < 2e-5
    ref, _ = orc.tacaw_intensity(x.cpu().numpy().astype(np.complex128), np.arange(T) * 0.01)
    assert rel_l2(out[1][:, keep], ref[:, keep]) < 1e-4
    assert np.abs(out[1][:, dc]).max() <= 1e-6 * np.abs(ref).max() * T      # DC bin: rounding noise of the mean only
    assert np.array_equal(out[1], engine.tacaw_intensity(x).cpu().numpy())  # deterministic


@pytest.mark.parametrize("box,grid", [((25.55, 25.55, 12.2), (256, 256)), ((51.15, 25.55, 3.1), (512, 256)),
                                      ((102.35, 102.35, 2.6), (1024, 1024))])
def test_phase_stack_equals_complex_stack(box, grid):
    """The float32 phase format of the transmission stack (psb_build_phase / psb_propagate_phase: sigma*V stored,
    exp(i*phase) evaluated inside the fused row pass) against the complex64 stack: exit waves of the same probes."""
    from pyslice_b200 import _lib, engine, hostmath, synthetic
    traj = synthetic.random_trajectory(n_atoms=900, box=box, n_frames=5, seed=23, types=(6, 14, 31), stray=True)
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    assert (plan.nx, plan.ny) == grid and engine.phase_format_supported(plan)
    pos = dev(traj.positions)
    t = engine.build_transmission(plan, pos)
    ph = engine.build_transmission(plan, pos, phase=True)
    assert ph.dtype == torch.float32 and ph.shape == t.shape
    assert rel_l2(torch.polar(torch.ones_like(ph), ph).cpu().numpy(), t.cpu().numpy()) < 1e-6
    g = torch.Generator(device="cuda").manual_seed(1)
    probes = torch.randn((3,) + grid, generator=g, device="cuda", dtype=torch.float32).to(torch.complex64)
    a = engine.propagate(plan, probes, t).clone()
    b = engine.propagate(plan, probes, ph)
    assert rel_l2(b.cpu().numpy(), a.cpu().numpy()) < 5e-6
    # k-space output with layer taps goes through the same path
    L = engine.layer_count(plan.nz, 4)
    wa = torch.empty((L, 3, 5) + grid, dtype=torch.complex64, device="cuda")
    wb = torch.empty_like(wa)
    engine.propagate(plan, probes, t, wf_out=wa, layer_every=4)
    engine.propagate(plan, probes, ph, wf_out=wb, layer_every=4)
    assert rel_l2(wb.cpu().numpy(), wa.cpu().numpy()) < 5e-6
    # grids without fused kernels refuse the format loudly
    small = synthetic.random_trajectory(n_atoms=50, box=(6.35, 6.35, 2.1), n_frames=1, seed=1)
    sxs, sys_, szs, *_ = hostmath.grid_from_box(small.box_matrix)
    splan = engine.make_plan(sxs, sys_, szs, small.atom_types.tolist(), 100e3)
    assert not engine.phase_format_supported(splan)
    with pytest.raises(_lib.PsbError):
        engine.build_transmission(splan, dev(small.positions), phase=True)


def test_graph_replay_is_bit_identical_and_survives_workspace_growth():
    """CUDA-graph replay inside libpsb (graph_cache.cu): the third and later calls with the same buffers replay the recorded
    launch sequence.  Results must equal direct launches bit for bit -- phase stack, exit waves, layer taps -- also after a
    larger job in between made the structure-factor workspace grow (recorded graphs pointing at the old block are dropped)."""
    from pyslice_b200 import engine, synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    traj = synthetic.random_trajectory(n_atoms=500, box=(25.55, 25.55, 5.1), n_frames=6, seed=77, types=(6, 14))
    big = synthetic.random_trajectory(n_atoms=4000, box=(25.55, 25.55, 5.1), n_frames=6, seed=78, types=(6, 14, 31))
    pp = [(3.0, 4.0), (12.5, 20.0)]

    def runs(calc, n):
        return [calc.run().wavefunction_data.clone() for _ in range(n)]

    engine.set_graph_mode(False)
    try:
        ref = MultisliceCalculator()
        ref.setup(traj, aperture=0.0, voltage_eV=100e3, layer_every=4)
        want = runs(ref, 1)[0]
        ref2 = MultisliceCalculator()
        ref2.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp)
        want2 = runs(ref2, 1)[0]
    finally:
        engine.set_graph_mode(True)
    l0 = engine.launch_count()
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3, layer_every=4)
    got = runs(calc, 4)                      # eager, capture, replay, replay
    for g in got:
        assert torch.equal(g, want)
    per_run = (engine.launch_count() - l0) // 4
    assert per_run > 20                      # replayed launches are still counted
    calc2 = MultisliceCalculator()
    calc2.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp)
    for g in runs(calc2, 3):
        assert torch.equal(g, want2)
    other = MultisliceCalculator()           # more atoms per frame: the phase-table workspace grows
    other.setup(big, aperture=0.0, voltage_eV=100e3)
    runs(other, 3)
    for g in runs(calc, 3):                  # the small job again, through graphs recorded afresh
        assert torch.equal(g, want)


@pytest.mark.parametrize("box,n_atoms,types,n_frames", [
    ((102.35, 102.35, 1.7), 6000, (14,), 2),          # 1024 x 1024, 3 slices (last pair half empty), ~4000 atoms per pair
    ((102.35, 25.55, 1.2), 2500, (6, 14), 2),         # 1024 x 256, two types
    ((51.15, 51.15, 2.2), 3000, (5, 6, 7), 2),        # 512 x 512 (1024-point fine grid), three types
    ((51.15, 102.35, 0.7), 1500, (14,), 1),           # 512 x 1024, one slice pair
])
def test_potential_nufft_vs_direct_sum_and_oracle(box, n_atoms, types, n_frames):
    """Structure factor of dense slices through the 1-D NUFFT along x (sf_nufft.cu: ES-kernel spreading onto a twofold
    oversampled grid in shared memory, fused 2 nx-point transform, deconvolution) against the direct sum (sf_fast.cu) on
    the same inputs -- atoms on slice bounds and outside the box included -- and against the oracle's potential."""
    from pyslice_b200 import engine, hostmath, synthetic
    traj = synthetic.random_trajectory(n_atoms=n_atoms, box=box, n_frames=n_frames, seed=29, types=types, stray=True)
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    pos = dev(traj.positions)
    out = {}
    try:
        for mode in (2, 1):
            engine.set_sf_mode(mode)
            t, V = engine.build_transmission(plan, pos, want_potential=True)
            out[mode] = (t.clone(), V.clone())
        engine.set_sf_mode(2)
        t2, V2 = engine.build_transmission(plan, pos, want_potential=True)
        assert torch.equal(V2, out[2][1]) and torch.equal(t2, out[2][0])          # reproducible bit for bit
        ph = engine.build_transmission(plan, pos, phase=True)
        assert rel_l2(ph.cpu().numpy(), (plan.sigma * out[2][1]).cpu().numpy()) < 1e-6
    finally:
        engine.set_sf_mode(0)
    err = rel_l2(out[2][1].cpu().numpy(), out[1][1].cpu().numpy())
    print(f"NUFFT vs direct sum, potential rel-L2: {err:.3e}")
    assert err < 3e-6
    Vref = orc.potential(xs, ys, zs, traj.positions[0], traj.atom_types, workers=8)      # (nx, ny, nz)
    for mode in (2, 1):
        e = rel_l2(np.moveaxis(out[mode][1][0].cpu().numpy(), 0, 2), Vref)
        print(f"potential rel-L2 vs oracle, sf mode {mode}: {e:.3e}")
        assert e < 1e-5, mode
