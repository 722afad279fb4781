#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== microbench"; timeout 600 python tools/microbench_passes.py 256 64 24 48 96 148 192 296 2>&1 | tee gpurun_out/micro_256.log
timeout 600 python tools/microbench_passes.py 512 32 12 24 37 74 2>&1 | tee gpurun_out/micro_512.log
timeout 600 python tools/microbench_passes.py 1024 16 6 12 18 2>&1 | tee gpurun_out/micro_1024.log
echo "== ncu full (R and C pass at 256, F=96)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psb_kernel -s 20 -c 2 -o gpurun_out/prof_slice_step_v1 \
    python tools/microbench_passes.py 256 16 96 > gpurun_out/ncu_full_run.log 2>&1
tail -3 gpurun_out/ncu_full_run.log
ls -la gpurun_out
