"""Stage-by-stage A/B of the fused kernels against the generic ones (diagnostics)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyslice_b200 import engine, hostmath, synthetic
from pyslice_b200.multislice.multislice import Probe, create_batched_probes

def report(name, a, b):
    d = (a - b).abs()
    rel = float(d.norm() / b.norm())
    idx = torch.nonzero(d > 1e-4 * float(b.abs().max()))
    print(f"{name}: rel-L2 {rel:.3e}  max|d| {float(d.max()):.3e}  n_bad {idx.shape[0]}", flush=True)
    if idx.shape[0]:
        print("   first bad:", idx[:8].tolist(), " last bad:", idx[-3:].tolist())
        for c in range(idx.shape[1]):
            u = torch.unique(idx[:, c])
            print(f"   axis {c}: {u.numel()} distinct, e.g. {u[:12].tolist()}")

for name, traj, pp, ap in [
    ("si_c1", synthetic.silicon_trajectory(cells=(5, 5, 10), a=5.11, n_frames=1, seed=0), None, 0.0),
    ("gas5", synthetic.random_trajectory(n_atoms=600, box=(25.55, 25.55, 6.1), n_frames=5, seed=31, types=(6, 14)), [(3.0, 4.0), (12.2, 20.1), (21.0, 7.7)], 30.0),
]:
    print("=====", name)
    xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
    plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
    pos = torch.from_numpy(traj.positions).cuda()
    res = {}
    for fast in (True, False):
        engine.set_fast_path(fast)
        t, V = engine.build_transmission(plan, pos, want_potential=True)
        t2 = engine.build_transmission(plan, pos)
        res[fast] = (t.clone(), V.clone(), t2.clone())
    engine.set_fast_path(True)
    report("V fast vs generic", res[True][1], res[False][1])
    report("t fast vs generic", res[True][0], res[False][0])
    report("t (no v_out) fast vs generic", res[True][2], res[False][2])
    report("t fast: v_out vs no v_out", res[True][2], res[True][0])
    base = Probe(xs, ys, ap, 100e3)
    probes = create_batched_probes(base, pp if pp else [(xs[-1] / 2, ys[-1] / 2)]).array
    tt = res[False][0]
    out = {}
    for fast in (True, False):
        engine.set_fast_path(fast)
        out[fast] = engine.propagate(plan, probes, tt).clone()
    engine.set_fast_path(True)
    report("propagate (same t) fast vs generic", out[True], out[False])
    # single slice step: nz = 2 view
    for nzs in (2, 3):
        plan2 = engine.make_plan(xs, ys, zs[:nzs], traj.atom_types.tolist(), 100e3)
        t_s = tt[:, :nzs].contiguous()
        o = {}
        for fast in (True, False):
            engine.set_fast_path(fast)
            o[fast] = engine.propagate(plan2, probes, t_s).clone()
        engine.set_fast_path(True)
        report(f"propagate nz={nzs} fast vs generic", o[True], o[False])
