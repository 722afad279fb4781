"""Time the slice-step passes in isolation (random data) for several batch sizes / grids.
usage: python tools/microbench_passes.py [n] [nz] [batches...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyslice_b200 import engine, hostmath, _lib
if os.environ.get("PSB_VARIANT_LIB"):          # tuning experiments: time an alternative build of libpsb
    _lib._lib = _lib.load(os.path.abspath(os.environ["PSB_VARIANT_LIB"]))
    print("variant library:", os.environ["PSB_VARIANT_LIB"])

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 64
batches = [int(x) for x in sys.argv[3:]] or [24, 48, 96, 148, 192, 296]
dev = torch.device("cuda")
L = n * 0.1 - 0.05
xs = np.linspace(0, L, n, endpoint=False); zs = np.linspace(0, nz * 0.5, nz, endpoint=False)
plan = engine.make_plan(xs, xs, zs, [14], 100e3)
probe = torch.ones((1, n, n), dtype=torch.complex64, device=dev)
for F in batches:
    ph = torch.rand((F, nz, n, n), device=dev) * 6.28
    # PSB_PHASE=1: the stack as float32 phases (single-probe format), else complex64 t
    t = ph if os.environ.get("PSB_PHASE") == "1" else torch.polar(torch.ones_like(ph), ph)
    del ph
    out = torch.empty((1, 1, F, n, n), dtype=torch.complex64, device=dev)
    work = torch.empty((F, n, n), dtype=torch.complex64, device=dev)
    for fast in ([True, False] if os.environ.get("PSB_AB", "1") == "1" else [True]):
        engine.set_fast_path(fast)
        for _ in range(2):
            engine.propagate(plan, probe, t, wf_out=out, work=work)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        a.record()
        for _ in range(reps):
            engine.propagate(plan, probe, t, wf_out=out, work=work)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        ss = F * nz / (ms * 1e-3)
        bss = n * n * 20
        print(f"n={n} nz={nz} F={F:4d} {'fused  ' if fast else 'generic'}: {ms:8.3f} ms  {ss/1e6:6.3f} M slice-steps/s  "
              f"{ss*bss/1e9:7.1f} GB/s algorithmic  ({1e3*ms/nz:6.1f} us per slice-step batch)", flush=True)
    engine.set_fast_path(True)
    del t, out, work
