#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== potential microbench"; timeout 600 python tools/microbench_potential.py 32 32 64 96 2>&1 | tee gpurun_out/micro_potential.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
