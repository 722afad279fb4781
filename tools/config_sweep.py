"""One-GPU timings of the BASELINE.json configurations other than the bench workload (C2), through the public API
(MultisliceCalculator.setup()/run() + TACAWData) with positions resident in HBM, CUDA-event timed.

    python tools/config_sweep.py [c1 c3 c4 c5]

Frame counts are bounded where the full configuration does not fit one GPU's 180 GB (stated per line):
  c1  Si 2 000 atoms, 256 x 256 x 103, 20 frames, plane wave                      (full)
  c3  hBN/graphene 9 600 atoms, 512 x 512 x 67, 100 frames, 16 x 16 probes, 30 mrad (full: 53.7 GB of exit waves)
  c4  Si 38 400 atoms, 1024 x 1024 x 123, 250 of 2 000 frames (one GPU's share of the 8-GPU run), plane wave
  c5  c3's sample, 8 x 8 probes, layer every 10th slice (7 layers), 100 of 500 frames (94 GB of exit waves)
Prints one JSON line per configuration: slice-steps/s for the whole job, phase times, roofline fraction of the
propagate phase against B_ss = nx*ny*(16 + 4/P).
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyslice_b200 import engine, synthetic
from pyslice_b200.multislice.calculators import MultisliceCalculator
from pyslice_b200.multislice.multislice import probe_grid
from pyslice_b200.multislice.trajectory import Trajectory
from pyslice_b200.postprocessing.tacaw_data import TACAWData

CONFIGS = {
    "c1": dict(make=lambda: synthetic.silicon_trajectory(cells=(5, 5, 10), a=5.11, n_frames=20, seed=0), probes=0,
               aperture=0.0, layer_every=0, note="full configuration"),
    "c3": dict(make=lambda: synthetic.hbn_graphene_trajectory(n_frames=100, seed=2), probes=16, aperture=30.0,
               layer_every=0, note="full configuration"),
    "c4": dict(make=lambda: synthetic.silicon_trajectory(cells=(20, 20, 12), a=5.1175, n_frames=250, seed=3,
                                                         displacement="phonon"), probes=0, aperture=0.0, layer_every=0,
               note="250 of 2000 frames = one GPU's share of the 8-GPU run"),
    "c5": dict(make=lambda: synthetic.hbn_graphene_trajectory(n_frames=100, seed=4), probes=8, aperture=30.0,
               layer_every=10, note="100 of 500 frames (exit waves of 7 layers: 94 GB)"),
}


def run(name):
    cfg = CONFIGS[name]
    dev = torch.device("cuda", 0)
    traj = cfg["make"]()
    pos = torch.from_numpy(traj.positions).to(dev)
    dt = Trajectory.__new__(Trajectory)
    dt.atom_types, dt.positions, dt.velocities = traj.atom_types, pos, None
    dt.box_matrix, dt.timestep = traj.box_matrix, traj.timestep
    lx, ly = traj.box_matrix[0, 0], traj.box_matrix[1, 1]
    pp = None
    if cfg["probes"]:
        n = cfg["probes"]
        pp = [tuple(p) for p in probe_grid([0.25 * lx, 0.75 * lx], [0.25 * ly, 0.75 * ly], n, n)]
    calc = MultisliceCalculator(device=dev)
    calc.setup(dt, aperture=cfg["aperture"], voltage_eV=100e3, probe_positions=pp, layer_every=cfg["layer_every"],
               shard_frames=False)
    P, T, nz = calc.n_probes, calc.n_frames, calc.nz
    timer = engine.PhaseTimer(dev)

    def step(tm=None):
        wf = calc.run(timer=tm)
        tac = TACAWData(wf)                  # last layer (the exit wave)
        return wf, tac

    wf, tac = step()                      # warm-up (tables, allocator)
    del wf, tac
    torch.cuda.synchronize()
    l0 = engine.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for _ in range(reps):
        wf, tac = step(timer)
        finite = bool(torch.isfinite(tac.intensity[0, :, 0, 0]).all())
        del wf, tac
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    ph = {k: v / reps for k, v in timer.totals().items()}
    steps = P * T * nz
    b_ss = calc.nx * calc.ny * (16 + 4 / P)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    prop = ph.get("propagate", 0.0)
    print(json.dumps({
        "config": name, "note": cfg["note"], "grid": [calc.nx, calc.ny, nz], "atoms": int(traj.n_atoms), "frames": T,
        "probes": P, "layers": int(calc.n_layers), "slice_steps": steps, "ms_per_job": ms,
        "slice_steps_per_s": steps / (ms * 1e-3), "phases_ms": ph,
        "propagate_slice_steps_per_s": steps / (prop * 1e-3) if prop else None,
        "roofline_frac_propagate": (steps * b_ss / (prop * 1e-3) / 1e9 / peak) if prop else None,
        "hbm_peak_gbs": peak, "frames_per_batch": engine.batch_sizes(calc._plan, P, T)[0],
        "gpu_launches_per_job": (engine.launch_count() - l0) // reps, "finite": finite}), flush=True)


if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CONFIGS)):
        run(name)
        torch.cuda.empty_cache()
