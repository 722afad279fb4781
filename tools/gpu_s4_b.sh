#!/bin/bash
# session 4, call b: warm-cache (no flush) per-kernel times of the potential chain + full-set captures
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential microbench"; PSB_FAST_ONLY=1 timeout 300 python tools/microbench_potential.py 32 64 2>&1 | head -3 | tee gpurun_out/s4b_micro_pot.log
echo "== ncu warm launch list (potential, 16 frames)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 600 --csv --log-file gpurun_out/s4b_launches_warm.csv \
    python tools/microbench_potential.py 16 64 > gpurun_out/s4b_run1.log 2>&1
echo "== ncu full, warm, potential kernels"
timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'sf_tiles|fast_|phase_tables' -s 40 -c 8 -o gpurun_out/s4b_prof_potential \
    python tools/microbench_potential.py 16 64 > gpurun_out/s4b_run2.log 2>&1
echo "== ncu full, warm, slice step"
PSB_AB=0 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:fast_ -s 40 -c 2 -o gpurun_out/s4b_prof_slice_step_warm \
    python tools/microbench_passes.py 256 32 148 > gpurun_out/s4b_run3.log 2>&1
ls -la gpurun_out | tail -8
