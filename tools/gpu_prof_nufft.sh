#!/bin/bash
# full ncu capture of the NUFFT column kernel at C4 geometry (one frame): gpurun_out/<tag>_prof_nufft.ncu-rep
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=${1:-prof}
PSB_SF_MODE=2 PSB_GRAPHS=0 PSB_LEVELS=1 PSB_GEOM=c4 PSB_PHASE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nufft_cols -s 2 -c 1 \
   -o gpurun_out/${T}_prof_nufft python tools/microbench_potential.py 1 64 > gpurun_out/${T}_ncu_run.log 2>&1
tail -3 gpurun_out/${T}_ncu_run.log
