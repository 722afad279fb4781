#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== ncu launches (96 frames)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_r1b.csv \
    python bench.py --frames 96 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
tail -c 600 gpurun_out/ncu_launches_run.log
echo "== ncu full potential kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psb_kernel -s 3 -c 3 -o gpurun_out/prof_potential_r1b \
    python bench.py --frames 96 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
echo "== ncu full slice-step kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psb_kernel -s 20 -c 2 -o gpurun_out/prof_slice_step_r1b \
    python tools/microbench_passes.py 256 16 96 > gpurun_out/ncu_full_run2.log 2>&1
ls -la gpurun_out
