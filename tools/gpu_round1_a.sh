#!/bin/bash
# first GPU call: parity tests, smoke, bench, ncu launch list + full capture of the slice-step kernels
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest" ; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 2 2>&1 | tail -3 | tee gpurun_out/bench.log
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --frames 8 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:LinePass -s 40 -c 4 -o gpurun_out/prof_slice_step \
    python bench.py --frames 96 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
