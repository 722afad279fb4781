#!/bin/bash
# session-3 run E: pipelined structure factor + fused inverse transforms: parity, full suite, bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "potential or fused" 2>&1 | tail -15 | tee gpurun_out/pytest_potential.log
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench.log
