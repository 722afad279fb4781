"""Time psb_tacaw_intensity (|fftshift FFT_t(psi - mean)|^2) at configuration scale, tiled mixed-radix kernel (level 1)
against the generic Bluestein line pass (level 0).  Algorithmic bytes: 12 per (p, w, k) element.
usage: python tools/microbench_tacaw.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyslice_b200 import engine, _lib
if os.environ.get("PSB_VARIANT_LIB"):          # tuning experiments: time an alternative build of libpsb
    _lib._lib = _lib.load(os.path.abspath(os.environ["PSB_VARIANT_LIB"]))
    print("variant library:", os.environ["PSB_VARIANT_LIB"])
CASES = [("C2: P=1, T=500, 256x256", 1, 500, 256, 256), ("C3 quarter: P=64, T=100, 512x512", 64, 100, 512, 512),
         ("C4 per GPU of 8: P=1, T=2000, 128x1024", 1, 2000, 128, 1024), ("C1: P=1, T=20, 256x256", 1, 20, 256, 256)]
for name, P, T, nx, ny in CASES:
    x = torch.randn((P, T, nx, ny), device="cuda", dtype=torch.float32).to(torch.complex64) + 2.0
    for level in (1, 0):
        engine.set_fast_path(level)
        for _ in range(2):
            out = engine.tacaw_intensity(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            out = engine.tacaw_intensity(x)
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 3
        gb = 12.0 * P * T * nx * ny / 1e9
        print(f"{name:42s} level {level}: {ms:9.3f} ms  {gb / (ms * 1e-3):8.1f} GB/s algorithmic", flush=True)
    engine.set_fast_path(True)
    del x, out
