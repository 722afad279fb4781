#!/bin/bash
# session 4, call n: first-slice totals stashed in shared memory (structure-factor tiles)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "potential" 2>&1 | tail -3 | tee gpurun_out/s4n_pytest_potential.log
PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 100 64 2>&1 | grep level | tee -a gpurun_out/s4n_micro.log
PSB_GEOM=c4 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 64 2>&1 | grep level | tee -a gpurun_out/s4n_micro.log
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/s4n_bench.log
