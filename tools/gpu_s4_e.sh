#!/bin/bash
# session 4, call e: streaming t stores / evict-first chunk reads in the transmission pass, chunk sizes in whole rounds
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "potential" 2>&1 | tail -3 | tee gpurun_out/s4e_pytest_potential.log
echo "== potential microbench"; PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 32 32 48 64 80 96 2>&1 | grep level | tee gpurun_out/s4e_micro_pot.log
echo "== slice-step microbench"; PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 100 148 2>&1 | tee gpurun_out/s4e_micro_256.log
echo "== ncu warm launch list (potential, 16 frames)"
PSB_LEVELS=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 300 --csv --log-file gpurun_out/s4e_launches_warm.csv \
    python tools/microbench_potential.py 16 64 > gpurun_out/s4e_run1.log 2>&1
echo "== bench"; timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/s4e_bench.log
