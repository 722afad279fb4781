"""Worker of tests/test_gpu_nccl.py (launched under torch.distributed.run, one process per GPU, NCCL):
frame-sharded run + the pipeline's one all-to-all + row-sharded time transform against the same job computed
unsharded on every rank.  The reference has no multi-device path (sequential frame loop, src/multislice/calculators.py:172),
so the single-GPU result is the oracle: intensity rows must agree bit for bit, reducers to float64 round-off."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from pyslice_b200 import synthetic
    from pyslice_b200.multislice.calculators import MultisliceCalculator
    from pyslice_b200.postprocessing.tacaw_data import TACAWData
    import bench
    assert bench.multi_gpu_parity(dev, rank, world) == "bitwise"
    # several probes, a non-power-of-two grid (generic kernels, ragged kx-row split) and every reducer
    for box, n_frames, pp, ap in [((4.75, 3.95, 3.2), 4 * world + 1, [(1.0, 2.0), (3.3, 1.4), (2.2, 2.2)], 30.0),
                                  ((25.55, 25.55, 3.1), 2 * world + 1, None, 0.0)]:
        traj = synthetic.random_trajectory(n_atoms=150, box=box, n_frames=n_frames, seed=9, types=(5, 7))
        calc = MultisliceCalculator(device=dev)
        calc.setup(traj, aperture=ap, voltage_eV=100e3, probe_positions=pp)
        tac = TACAWData(calc.run())
        single = MultisliceCalculator(device=dev)
        single.setup(traj, aperture=ap, voltage_eV=100e3, probe_positions=pp, shard_frames=False)
        tac1 = TACAWData(single.run())
        r0, r1 = tac.row_range
        assert torch.equal(tac.intensity, tac1.intensity[:, :, r0:r1]), "intensity rows differ"
        np.testing.assert_allclose(tac.spectrum(), tac1.spectrum(), rtol=1e-12)
        np.testing.assert_allclose(tac.spectrum(0), tac1.spectrum(0), rtol=1e-12)
        np.testing.assert_allclose(tac.diffraction(), tac1.diffraction(), rtol=1e-6)
        np.testing.assert_allclose(tac.spectral_diffraction(10.0), tac1.spectral_diffraction(10.0), rtol=0, atol=0)
        np.testing.assert_allclose(tac.spectrum_image(10.0), tac1.spectrum_image(10.0), rtol=1e-12)
        kx, ky = np.linspace(-3, 3, 5), np.linspace(-1, 2, 5)
        np.testing.assert_allclose(tac.dispersion(kx, ky), tac1.dispersion(kx, ky), rtol=0, atol=0)
        nx, ny = len(tac1.kxs), len(tac1.kys)
        mask = (np.arange(nx)[:, None] + np.arange(ny)[None, :]) % 3 == 0
        np.testing.assert_allclose(tac.masked_spectrum(mask), tac1.masked_spectrum(mask), rtol=1e-12)
    dist.barrier()
    if rank == 0:
        print("NCCL_PARITY_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
