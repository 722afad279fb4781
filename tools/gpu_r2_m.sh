#!/bin/bash
# round 2, call g: structure factor through the 1-D NUFFT (parity, speed at C4 geometry), workspace cache / e2e
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2m
echo "== nufft parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "nufft" 2>&1 | grep -E "rel-L2|passed|failed|Error|error|assert" | tail -30 | tee gpurun_out/${T}_pytest_nufft.log
echo "== potential microbench, C4 geometry"
for mode in 1 2; do PSB_SF_MODE=$mode PSB_GEOM=c4 PSB_LEVELS=1 PSB_PHASE=1 timeout 600 python tools/microbench_potential.py 8 64 2>&1 | tee -a gpurun_out/${T}_micro_pot.log; done
echo "== ncu launch list, nufft at C4"
PSB_SF_MODE=2 PSB_GRAPHS=0 PSB_LEVELS=1 PSB_GEOM=c4 PSB_PHASE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 80 --csv \
   --log-file gpurun_out/${T}_launches_pot_c4_nufft.csv python tools/microbench_potential.py 2 64 > gpurun_out/${T}_ncu_run2.log 2>&1
echo "== bench c4 250 (auto mode)"; timeout 900 python bench.py --workload c4 --frames 250 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
echo "== c4 grid parity test"; timeout 600 python -m pytest tests/test_gpu_config_scale.py -q -m gpu -x -k "c4_grid" 2>&1 | tail -3 | tee -a gpurun_out/${T}_pytest_nufft.log
echo "== bench default"; timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
ls -la gpurun_out | grep ${T}
