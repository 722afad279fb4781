#!/bin/bash
# session 4, call c: fused structure factor + column transform, SFU cis: parity, then timings
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -s -k "potential" 2>&1 | tail -25 | tee gpurun_out/s4c_pytest_potential.log
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/s4c_pytest_gpu.log
echo "== potential microbench"; PSB_LEVELS=2,1 timeout 300 python tools/microbench_potential.py 32 64 2>&1 | tee gpurun_out/s4c_micro_pot.log
echo "== bench"; timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/s4c_bench.log
echo "== ncu warm launch list (potential, 16 frames)"
PSB_LEVELS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file gpurun_out/s4c_launches_warm.csv \
    python tools/microbench_potential.py 16 64 > gpurun_out/s4c_run1.log 2>&1
echo "== ncu full, potential kernels"
PSB_LEVELS=2 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'sf_cols|fast_rows|sfc_tables' -s 30 -c 3 -o gpurun_out/s4c_prof_potential \
    python tools/microbench_potential.py 16 64 > gpurun_out/s4c_run2.log 2>&1
ls -la gpurun_out | tail -5
