#!/bin/bash
# round 2, call z2: ncu full-set captures of the TACAW kernel at C3-quarter scale, both variants
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2z
for v in 0 1; do
PSB_TACAW_TMA=$v PSB_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tacaw_ -s 6 -c 1 -o gpurun_out/${T}_prof_tacaw_tma$v \
    python tools/microbench_tacaw.py > gpurun_out/${T}_ncu_tacaw_run$v.log 2>&1
done
ls -la gpurun_out | grep ${T}_prof
