#!/bin/bash
# round 2, call ag: live DRAM traffic of the potential kernels against the chunk size (does a smaller chunk stay in L2?)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2ag
for mb in 16 32 48 64; do
PSB_GRAPHS=0 PSB_PHASE=1 PSB_LEVELS=1 timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none \
    -k regex:'sf_tiles|fast_|phase_tables' -s 400 -c 256 --csv --log-file gpurun_out/${T}_traffic_${mb}mb.csv python tools/microbench_potential.py 8 $mb > gpurun_out/${T}_run_${mb}.log 2>&1
done
ls -la gpurun_out | grep ${T}
