#!/bin/bash
# session 4, call h: folded radix-2 stage (512), staged offsets in sf_tiles, fewest-rounds frame batches
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/s4h_pytest_gpu.log
echo "== potential microbench"; PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 100 64 2>&1 | grep level | tee gpurun_out/s4h_micro_pot.log
echo "== slice-step microbench"; timeout 300 python tools/microbench_passes.py 256 64 100 127 148 2>&1 | tee gpurun_out/s4h_micro_256.log
timeout 300 python tools/microbench_passes.py 512 32 37 74 2>&1 | tee gpurun_out/s4h_micro_512.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/s4h_bench.log
