#!/bin/bash
# round 2, call r: full suite + evidence with the NUFFT structure factor in the default path
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2r
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/${T}_smoke.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
for mode in 2 1; do PSB_SF_MODE=$mode PSB_GEOM=c4 PSB_LEVELS=1 PSB_PHASE=1 timeout 600 python tools/microbench_potential.py 8 64 2>&1 | tee -a gpurun_out/${T}_micro_pot.log; done
PSB_SF_MODE=2 PSB_GRAPHS=0 PSB_LEVELS=1 PSB_GEOM=c4 PSB_PHASE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 80 --csv \
   --log-file gpurun_out/${T}_launches_pot_c4_nufft.csv python tools/microbench_potential.py 2 64 > gpurun_out/${T}_ncu_run2.log 2>&1
echo "== bench c4 250"; timeout 900 python bench.py --workload c4 --frames 250 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
echo "== bench c4 full (2000 frames, one GPU)"; timeout 1200 python bench.py --workload c4 --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_full.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
