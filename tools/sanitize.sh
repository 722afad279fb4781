#!/bin/bash
# compute-sanitizer over the small end-to-end cases of smoke() with both kernel generations:
# memcheck (out-of-bounds / misaligned accesses) and racecheck (shared-memory hazards between the threads of a CTA; it does
# not see the async proxy, so the TMA landing buffers are covered by the bit-exact A/B tests instead).  Run on the GPU box:
#     gpurun -- bash tools/sanitize.sh          -> gpurun_out/sanitize_*.log
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
cat > /tmp/psb_sanitize_case.py <<'PY'
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pyslice_b200 import engine, synthetic
from pyslice_b200.multislice.calculators import MultisliceCalculator
from pyslice_b200.postprocessing.tacaw_data import TACAWData
level = int(sys.argv[1])
engine.set_fast_path(level)
engine.set_graph_mode(False)
cases = [((6.35, 6.35, 2.1), 120, None, 0.0),                       # 64 x 64: generic kernels
         ((25.55, 25.55, 2.1), 200, None, 0.0),                      # 256 x 256 plane wave: phase stack, fused kernels
         ((25.55, 51.15, 1.6), 200, [(3.0, 7.0), (12.0, 30.0)], 25.0)]   # 256 x 512, two probes: complex stack
if len(sys.argv) > 2 and sys.argv[2] == "big":
    cases.append(((102.35, 102.35, 1.1), 300, None, 0.0))           # 1024 x 1024: warp-pair row pass, in-place middle stage
for box, n, pp, ap in cases:
    traj = synthetic.random_trajectory(n_atoms=n, box=box, n_frames=4, seed=3, types=(6, 14))
    calc = MultisliceCalculator()
    calc.setup(traj, aperture=ap, voltage_eV=100e3, probe_positions=pp, layer_every=2)
    tac = TACAWData(calc.run())
    s = tac.spectrum()
    assert np.isfinite(s).all()
# time transform alone: radix-10 / 20 prime-factor stages (T = 20 single stage, 100 = 10 x 10 on 320-thread tiles, 500, 2000 on a
# whole-SM tile), ragged pixel counts
for P, T, npix in [(2, 20, 70), (2, 100, 4 * 64 + 9), (1, 500, 40), (1, 2000, 20), (1, 48, 33)]:
    x = (torch.randn((P, T, 1, npix), device="cuda") + 1j * torch.randn((P, T, 1, npix), device="cuda")).to(torch.complex64) + 2.0
    out = engine.tacaw_intensity(x)
    assert torch.isfinite(out).all()
torch.cuda.synchronize()
print("sanitize case ok, level", level)
PY
for tool in memcheck racecheck; do
  for level in 1 0; do
    echo "== compute-sanitizer --tool $tool, kernel generation $level"
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python /tmp/psb_sanitize_case.py $level $([ $tool = memcheck ] && echo big) \
        > gpurun_out/sanitize_${tool}_level${level}.log 2>&1
    tail -4 gpurun_out/sanitize_${tool}_level${level}.log
  done
done
