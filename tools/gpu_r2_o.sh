#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${1:-r2o}
echo "== nufft parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "nufft" 2>&1 | grep -E "rel-L2|passed|failed|Error|error|assert" | tail -20 | tee gpurun_out/${T}_pytest_nufft.log
for mode in 2 1; do PSB_SF_MODE=$mode PSB_GEOM=c4 PSB_LEVELS=1 PSB_PHASE=1 timeout 600 python tools/microbench_potential.py 8 64 2>&1 | tee -a gpurun_out/${T}_micro_pot.log; done
bash tools/gpu_prof_nufft.sh ${T}
echo "== bench c4 250 (auto mode)"; timeout 900 python bench.py --workload c4 --frames 250 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
