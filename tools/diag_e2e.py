"""Where the end-to-end arm of bench.py spends its time beyond the device-timed step (C2 by default): host-side
perf_counter around setup(), run(), TACAWData and the device -> host copy, each followed by a synchronize.
usage: python tools/diag_e2e.py [workload alias]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyslice_b200.multislice.calculators import MultisliceCalculator
from pyslice_b200.multislice.trajectory import Trajectory
from pyslice_b200.postprocessing.tacaw_data import TACAWData

name = bench.ALIASES.get(sys.argv[1] if len(sys.argv) > 1 else "c2")
wl = dict(bench.WORKLOADS[name])
if len(sys.argv) > 2:
    wl["frames"] = int(sys.argv[2])
dev = torch.device("cuda", 0)
traj = bench.make_traj(wl, wl["frames"])
pp = bench.probe_positions(wl, traj.box_matrix)
pinned = torch.empty(traj.positions.shape, dtype=torch.float64).pin_memory()
pinned.numpy()[...] = traj.positions
host = Trajectory(traj.atom_types, pinned.numpy(), np.zeros_like(traj.positions), traj.box_matrix, traj.timestep)
nx, ny, nz = wl["grid"]
out_host = torch.empty((1, wl["frames"], nx, ny), dtype=torch.float32).pin_memory()
rows = []
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    c = MultisliceCalculator(device=dev)
    c.setup(host, aperture=wl["aperture"], voltage_eV=bench.VOLTAGE, probe_positions=pp, layer_every=wl["layer_every"], shard_frames=False)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    wf = c.run()
    t1b = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    tac = TACAWData(wf)
    torch.cuda.synchronize(); t3 = time.perf_counter()
    if pp is None:
        out_host.copy_(tac.intensity, non_blocking=True)
    else:
        tac._sum_k(); tac.diffraction()
    torch.cuda.synchronize(); t4 = time.perf_counter()
    rows.append([1e3 * (b - a) for a, b in ((t0, t1), (t1, t1b), (t1, t2), (t2, t3), (t3, t4), (t0, t4))])
for r in rows:
    print("setup %7.2f  run: host returns after %7.2f, done %7.2f  tacaw %6.2f  result to host %6.2f  total %7.2f ms" % tuple(r))
