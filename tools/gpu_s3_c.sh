#!/bin/bash
# session-3 run C: ncu of the fused slice-step kernels (full set, 2 launches) + launch list of the microbench + batch sweep
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== batch sweep"; PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 37 74 111 148 2>&1 | tee gpurun_out/micro_256_sweep.log
echo "== ncu launch list (microbench)"
PSB_AB=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_micro_r1d.csv \
    python tools/microbench_passes.py 256 16 74 > gpurun_out/ncu_launches_micro.log 2>&1
echo "== ncu full fused kernels"
PSB_AB=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_ -s 20 -c 2 -o gpurun_out/prof_fused_r1d \
    python tools/microbench_passes.py 256 16 74 > gpurun_out/ncu_full_fused.log 2>&1
tail -3 gpurun_out/ncu_full_fused.log
ls -la gpurun_out
