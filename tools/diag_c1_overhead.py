"""Where does the time of a small job (C1: 20 frames, 3 ms of kernels) go?  Wall-clock per call with syncs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyslice_b200 import engine, synthetic
from pyslice_b200.multislice.calculators import MultisliceCalculator
from pyslice_b200.postprocessing.tacaw_data import TACAWData
traj = synthetic.silicon_trajectory(cells=(5, 5, 10), a=5.11, n_frames=20, seed=0)
calc = MultisliceCalculator(device="cuda:0")
def tick(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print(f"{label:28s} {1e3*(t1-t0):8.2f} ms"); return time.perf_counter()
for rep in range(3):
    print("--- rep", rep)
    t0 = time.perf_counter()
    calc.setup(traj, aperture=0.0, voltage_eV=100e3); t0 = tick("setup", t0)
    wf = calc.run(); t0 = tick("run", t0)
    tac = TACAWData(wf); t0 = tick("TACAWData", t0)
    s = tac.spectrum(); t0 = tick("spectrum", t0)
    plan = calc._plan
    pos = torch.from_numpy(traj.positions).cuda(); t0 = tick("upload", t0)
    t = engine.build_transmission(plan, pos); t0 = tick("build_transmission", t0)
    store = torch.empty((1, 1, 20, plan.nx, plan.ny), dtype=torch.complex64, device="cuda"); t0 = tick("alloc store", t0)
    engine.propagate(plan, calc._probes, t, wf_out=store); t0 = tick("propagate", t0)
    x = engine.tacaw_intensity(store[0]); t0 = tick("tacaw_intensity", t0)
    fb = engine.batch_sizes(plan, 1, 20); t0 = tick("batch_sizes", t0)
    del t, store, x, wf, tac
