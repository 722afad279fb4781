#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== microbench"; PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 74 125 148 2>&1 | tee gpurun_out/micro_256.log
PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | tee gpurun_out/micro_512.log
echo "== potential microbench"; timeout 600 python tools/microbench_potential.py 32 64 2>&1 | tee gpurun_out/micro_potential.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
