"""N > 1 host logic on CPU: two processes over gloo, kernels through the emulator.

Frames are sharded by rank for propagation, re-sharded to kx rows with the pipeline's single
all-to-all, and the reducers combine ranks -- results must equal the single-process run bit for bit
(intensity rows) / to float64 round-off (reducers)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_frames, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    import torch.distributed as dist
    from tests import emu
    emu.activate()
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        traj = synthetic.random_trajectory(n_atoms=60, box=(3.15, 2.35, 2.2), n_frames=n_frames, seed=8, types=(6, 14))
        calc = MultisliceCalculator()
        calc.setup(traj, aperture=0.0, voltage_eV=100e3)                     # sharded: world > 1
        assert calc.shard is not None and calc.shard.counts == [n_frames // 2 + n_frames % 2, n_frames // 2]
        wf = calc.run()
        assert wf.wavefunction_data.shape[1] == calc.shard.counts[rank]
        tac = TACAWData(wf)
        single = MultisliceCalculator()
        single.setup(traj, aperture=0.0, voltage_eV=100e3, shard_frames=False)
        tac1 = TACAWData(single.run())
        r0, r1 = tac.row_range
        assert tac.intensity.shape[2] == r1 - r0 and tac1.intensity.shape[2] == 32
        assert torch.equal(tac.intensity, tac1.intensity[:, :, r0:r1])       # multi-rank == single-rank, bitwise
        np.testing.assert_allclose(tac.spectrum(), tac1.spectrum(), rtol=1e-12)
        np.testing.assert_allclose(tac.diffraction(), tac1.diffraction(), rtol=1e-6)
        np.testing.assert_allclose(tac.spectral_diffraction(10.0), tac1.spectral_diffraction(10.0), rtol=0, atol=0)
        np.testing.assert_allclose(tac.spectrum_image(10.0), tac1.spectrum_image(10.0), rtol=1e-12)
        kx, ky = np.linspace(-3, 3, 5), np.linspace(-1, 2, 5)
        np.testing.assert_allclose(tac.dispersion(kx, ky), tac1.dispersion(kx, ky), rtol=0, atol=0)
        mask = (np.arange(32)[:, None] + np.arange(24)[None, :]) % 3 == 0
        np.testing.assert_allclose(tac.masked_spectrum(mask), tac1.masked_spectrum(mask), rtol=1e-12)
        # several probes + layer taps on a grid whose kx rows do not split evenly (23 rows over 2 ranks): multi-layer slab
        # store, the layer-wise exchange, HAADF over the blocks, and the detector-only run that keeps no exit waves
        from pyslice_b200.postprocessing.haadf_data import HAADFData
        traj2 = synthetic.random_trajectory(n_atoms=40, box=(2.25, 1.55, 1.2), n_frames=3, seed=4, types=(6, 14))
        pp = [(0.5, 0.4), (1.5, 1.0)]
        kw = dict(aperture=30.0, voltage_eV=100e3, probe_positions=pp, layer_every=2)
        calc2 = MultisliceCalculator()
        calc2.setup(traj2, **kw)
        wf2 = calc2.run()
        single2 = MultisliceCalculator()
        single2.setup(traj2, shard_frames=False, **kw)
        wf2s = single2.run()
        assert wf2.wavefunction_data.shape == (2, calc2.shard.counts[rank], 23, 16, 2)
        f0 = calc2.shard.start
        assert torch.equal(wf2.wavefunction_data.dense(), wf2s.wavefunction_data[:, f0:f0 + calc2.shard.counts[rank]])
        for layer in (0, 1):
            ta, tb = TACAWData(wf2, layer_index=layer), TACAWData(wf2s, layer_index=layer)
            r0, r1 = ta.row_range
            assert torch.equal(ta.intensity, tb.intensity[:, :, r0:r1])
        adf, adf1 = HAADFData(wf2).calculateADF(20.0), HAADFData(wf2s).calculateADF(20.0)
        np.testing.assert_allclose(adf.numpy(), adf1.numpy(), rtol=1e-6)
        calc3 = MultisliceCalculator()
        calc3.setup(traj2, adf_collection_angle=20.0, **kw)
        wf3 = calc3.run()
        assert wf3.wavefunction_data is None and wf3.adf_sums.shape == (2, 2, calc3.shard.counts[rank])
        np.testing.assert_allclose(HAADFData(wf3).calculateADF(20.0).numpy(), adf1.numpy(), rtol=1e-6)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        import traceback
        ret[rank] = "FAIL: " + traceback.format_exc()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [5])
def test_two_rank_frame_sharding_matches_single(n_frames):
    if torch.cuda.is_available():
        pytest.skip("CPU/gloo host-logic test; the GPU box runs the NCCL path in bench.py")
    from tests import emu
    emu.build()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_frames, ret), nprocs=2, join=True)
    assert ret.get(0) == "ok", ret.get(0)
    assert ret.get(1) == "ok", ret.get(1)
