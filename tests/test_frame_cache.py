"""Row f-3 of SURVEY.md section 8: the per-frame `.npy` cache of the reference (src/multislice/calculators.py
:78-92 key, :140 directory, :259-260 read, :276/:311 wire format), opt-in here.  Host logic on CPU, kernels through
the emulator.  The two golden keys were produced by the reference's own `_generate_cache_key` on these inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.skipif(torch.cuda.is_available(), reason="CPU host-logic test (kernel emulator)")


@pytest.fixture(scope="module")
def api():
    from tests import emu
    emu.build()
    emu.activate()
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    yield synthetic, MultisliceCalculator
    from pyslice_b200 import _lib
    _lib._reset()


def _traj(synthetic, seed=8):
    return synthetic.random_trajectory(n_atoms=60, box=(3.15, 2.35, 2.2), n_frames=4, seed=seed, types=(6, 14))


def test_reference_key_is_reproduced(api):
    synthetic, Calc = api
    traj = _traj(synthetic)
    c = Calc()
    assert c._generate_cache_key(traj, 0.0, 100e3, 0.5, 0.1, None) == "110b3a3a8537"
    assert c._generate_cache_key(traj, 25.0, 100e3, 0.5, 0.1, [(1.0, 1.5), (2.0, 0.5)]) == "ef6f46c19ed3"
    from pyslice_b200.multislice.trajectory import Trajectory
    other = Trajectory(traj.atom_types, traj.positions + 0.01, traj.velocities, traj.box_matrix, traj.timestep)   # same shapes, other positions
    assert c._generate_cache_key(other, 0.0, 100e3, 0.5, 0.1, None) == "110b3a3a8537"
    k1 = c._generate_cache_key(traj, 0.0, 100e3, 0.5, 0.1, None, with_positions=True)
    k2 = c._generate_cache_key(other, 0.0, 100e3, 0.5, 0.1, None, with_positions=True)
    assert k1 != k2 and len(k1) == 12


def test_cache_write_read_partial_and_cleanup(api, tmp_path):
    synthetic, Calc = api
    traj = _traj(synthetic)
    pp = [(1.0, 1.5), (2.0, 0.5)]
    plain = Calc()
    plain.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp)
    assert plain.output_dir is None
    ref = plain.run().wavefunction_data.clone()

    c = Calc()
    c.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp, frame_cache=tmp_path, cache_key="reference")
    assert c.output_dir == tmp_path / "torch_ef6f46c19ed3"
    wf = c.run()
    assert (c.frames_computed, c.frames_cached) == (4, 0)
    assert torch.equal(wf.wavefunction_data, ref)
    for f in range(4):
        a = np.load(c.output_dir / f"frame_{f}.npy")
        assert a.shape == (2, c.nx, c.ny, 1, 1) and a.dtype == np.complex128          # the reference's wire format
        assert np.array_equal(a[:, :, :, 0, 0].astype(np.complex64), ref[:, f, :, :, 0].cpu().numpy())

    # second run: everything comes from the files (one of them is doctored to prove it)
    np.save(c.output_dir / "frame_2.npy", 2.0 * np.load(c.output_dir / "frame_2.npy"))
    again = Calc()
    again.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp, frame_cache=tmp_path, cache_key="reference")
    wf2 = again.run()
    assert (again.frames_computed, again.frames_cached) == (0, 4)
    assert torch.equal(wf2.wavefunction_data[:, [0, 1, 3]], ref[:, [0, 1, 3]])
    assert torch.equal(wf2.wavefunction_data[:, 2], 2.0 * ref[:, 2])

    # partial cache: the missing frame in the middle is recomputed, the rest loaded; then clean up
    (c.output_dir / "frame_2.npy").unlink()
    (c.output_dir / "frame_0.npy").unlink()
    part = Calc()
    part.setup(traj, aperture=25.0, voltage_eV=100e3, probe_positions=pp, frame_cache=tmp_path, cache_key="reference",
               cleanup_temp_files=True)
    wf3 = part.run()
    assert (part.frames_computed, part.frames_cached) == (2, 2)
    assert torch.equal(wf3.wavefunction_data, ref)
    assert not part.output_dir.exists()


def test_cache_rejects_layers_and_bad_shapes(api, tmp_path):
    synthetic, Calc = api
    traj = _traj(synthetic)
    with pytest.raises(ValueError):
        Calc().setup(traj, voltage_eV=100e3, frame_cache=tmp_path, layer_every=2)
    with pytest.raises(ValueError):
        Calc().setup(traj, voltage_eV=100e3, frame_cache=tmp_path, cache_key="md5")
    c = Calc()
    c.setup(traj, voltage_eV=100e3, frame_cache=tmp_path)
    np.save(c.output_dir / "frame_1.npy", np.zeros((1, 3, 3, 1, 1), np.complex128))
    with pytest.raises(ValueError):
        c.run()
