#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== pytest" ; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== microbench"; timeout 600 python tools/microbench_passes.py 256 64 96 148 2>&1 | tee gpurun_out/micro_256.log
timeout 600 python tools/microbench_passes.py 512 32 24 2>&1 | tee gpurun_out/micro_512.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
