#!/bin/bash
# session-3 run B: first run of the fused slice-step kernels: parity (fused vs generic vs oracle), microbench A/B, suite, bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fused parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or si_c1" 2>&1 | tail -15 | tee gpurun_out/pytest_fused.log
echo "== microbench"; timeout 300 python tools/microbench_passes.py 256 64 74 96 2>&1 | tee gpurun_out/micro_256.log
timeout 300 python tools/microbench_passes.py 512 32 23 24 2>&1 | tee gpurun_out/micro_512.log
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
