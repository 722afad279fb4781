"""Does the potential build gain from two chunk pipelines in flight?  Builds the phase stack of F frames of the C2 geometry
(a) on one stream, (b) as two halves on two streams with their own scratch, and reports both times.  The structure-factor
kernel is FMA-bound and the two inverse transforms are memory-bound, and every chunk pays ~18 us of launch ramps and
tails (profiles/r2x_potential_chunk_sweep.txt), so overlapping neighbouring chunks is the candidate.
usage: python tools/microbench_potential_lanes.py [frames] [scratch_MB ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyslice_b200 import engine, hostmath, synthetic

F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sizes = [int(x) for x in sys.argv[2:]] or [32, 48, 64]
traj = synthetic.silicon_trajectory(cells=(5, 5, 50), a=5.11, n_frames=F, seed=1, displacement="phonon")
xs, ys, zs, *_ = hostmath.grid_from_box(traj.box_matrix)
plan = engine.make_plan(xs, ys, zs, traj.atom_types.tolist(), 100e3)
pos = torch.from_numpy(traj.positions).cuda()
out = torch.empty((F, plan.nz, plan.nx, plan.ny), dtype=torch.float32, device="cuda")
ref = torch.empty_like(out)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
h = F // 2
halves = [(0, h), (h, F)]
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def one():
    engine.build_transmission(plan, pos, out=ref, phase=True, scratch=scr1, bins=bins1)


def two():
    cur = torch.cuda.current_stream()
    for s in streams:
        s.wait_stream(cur)
    for k, (lo, hi) in enumerate(halves):
        with torch.cuda.stream(streams[k]):
            engine.build_transmission(plan, pos[lo:hi], out=out[lo:hi], phase=True, scratch=scr2[k], bins=bins2[k])
    for s in streams:
        cur.wait_stream(s)


for mb in sizes:
    engine.SCRATCH_BYTES = mb << 20
    n = plan.nx * plan.ny * engine.chunk_images(plan, F)
    scr1 = torch.empty((n,), dtype=torch.complex64, device="cuda")
    scr2 = [torch.empty((n,), dtype=torch.complex64, device="cuda") for _ in range(2)]
    bins1 = engine.bin_buffers(plan, F, pos.shape[1])
    bins2 = [engine.bin_buffers(plan, hi - lo, pos.shape[1]) for lo, hi in halves]
    for name, fn in (("one stream ", one), ("two streams", two)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(4):
            fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 4
        print(f"{name} scratch {mb:3d} MB each  F={F}: {ms:8.3f} ms  {1e3 * ms / (F * plan.nz):6.3f} us per slice", flush=True)
    print("   identical:", torch.equal(out, ref), flush=True)
