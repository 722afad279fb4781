"""Wall-clock breakdown of a STEM job through the public API (C5 recipe: 8 x 8 probes, 512 x 512 x 67, layers every
10th slice, 20 frames): where does the time outside the potential / propagate phases go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyslice_b200 import engine, synthetic
from pyslice_b200.multislice.calculators import MultisliceCalculator
from pyslice_b200.multislice.multislice import probe_grid
from pyslice_b200.postprocessing.tacaw_data import TACAWData
F = int(sys.argv[1]) if len(sys.argv) > 1 else 20
traj = synthetic.hbn_graphene_trajectory(n_frames=F, seed=4)
lx, ly = traj.box_matrix[0, 0], traj.box_matrix[1, 1]
pp = [tuple(p) for p in probe_grid([0.25 * lx, 0.75 * lx], [0.25 * ly, 0.75 * ly], 8, 8)]
calc = MultisliceCalculator(device="cuda:0")
def tick(label, t0):
    torch.cuda.synchronize(); t1 = time.perf_counter(); print(f"{label:28s} {1e3*(t1-t0):9.2f} ms", flush=True); return time.perf_counter()
for rep in range(3):
    print("--- rep", rep)
    t0 = time.perf_counter()
    calc.setup(traj, aperture=30.0, voltage_eV=100e3, probe_positions=pp, layer_every=10); t0 = tick("setup", t0)
    timer = engine.PhaseTimer(calc.device)
    wf = calc.run(timer=timer); t0 = tick("run", t0)
    print("   phases", {k: round(v, 2) for k, v in timer.totals().items()})
    t0 = time.perf_counter()
    tac = TACAWData(wf); t0 = tick("TACAWData", t0)
    x = wf.wavefunction_data[:, :, :, :, -1]; print("   layer view strides", x.stride(), x.shape)
    t0 = time.perf_counter()
    y = engine.tacaw_intensity(x); t0 = tick("tacaw_intensity alone", t0)
    s = tac.spectrum(); t0 = tick("spectrum", t0)
    del wf, tac, x, y; t0 = tick("del", t0)
