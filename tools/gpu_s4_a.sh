#!/bin/bash
# session 4, call a: full validation of the committed state (tests, bench, ncu launch list, ncu full of the slice step)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/s4a_smi.log
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/s4a_pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s4a_bench.log
echo "== microbench"; timeout 300 python tools/microbench_passes.py 256 64 148 2>&1 | tee gpurun_out/s4a_micro_256.log
timeout 300 python tools/microbench_passes.py 512 32 37 74 2>&1 | tee gpurun_out/s4a_micro_512.log
echo "== ncu launches (125 frames)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/s4a_launches.csv \
    python bench.py --frames 125 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s4a_ncu_launches_run.log 2>&1
tail -c 400 gpurun_out/s4a_ncu_launches_run.log
echo "== ncu full slice-step kernels"
PSB_AB=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_ -s 40 -c 4 -o gpurun_out/s4a_prof_slice_step \
    python tools/microbench_passes.py 256 32 148 > gpurun_out/s4a_ncu_full_run.log 2>&1
tail -3 gpurun_out/s4a_ncu_full_run.log
ls -la gpurun_out | tail -12
