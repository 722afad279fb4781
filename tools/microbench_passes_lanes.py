"""Slice step with the batch split over two streams (two independent half batches whose kernels fill each other's ramps and
tails) against one stream.  usage: python tools/microbench_passes_lanes.py [n] [nz] [batches...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyslice_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batches = [int(x) for x in sys.argv[3:]] or [127, 148]
dev = torch.device("cuda")
L = n * 0.1 - 0.05
xs = np.linspace(0, L, n, endpoint=False); zs = np.linspace(0, nz * 0.5, nz, endpoint=False)
plan = engine.make_plan(xs, xs, zs, [14], 100e3)
probe = torch.ones((1, n, n), dtype=torch.complex64, device=dev)
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
for F in batches:
    t = torch.rand((F, nz, n, n), device=dev) * 6.28
    out = torch.empty((1, 1, F, n, n), dtype=torch.complex64, device=dev)
    ref = torch.empty_like(out)
    work = torch.empty((F, n, n), dtype=torch.complex64, device=dev)
    t0 = torch.empty((F, n, n), dtype=torch.complex64, device=dev)
    h = (F + 1) // 2
    halves = [(0, h), (h, F)]

    def one():
        engine.propagate(plan, probe, t, wf_out=ref, work=work, t0=t0)

    def two():
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for k, (lo, hi) in enumerate(halves):
            with torch.cuda.stream(streams[k]):
                engine.propagate(plan, probe, t[lo:hi], wf_out=out, frame0=lo, work=work[lo:hi], t0=t0[lo:hi])
        for s in streams:
            cur.wait_stream(s)

    for name, fn in (("one stream ", one), ("two streams", two)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            fn()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"n={n} nz={nz} F={F:4d} {name}: {ms:8.3f} ms  {F * nz / (ms * 1e-3) / 1e6:6.3f} M slice-steps/s", flush=True)
    print("   identical:", torch.equal(out, ref), flush=True)
    del t, out, ref, work
