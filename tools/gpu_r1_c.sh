#!/bin/bash
# parity tests, smoke, both bench arms, ncu launch list + full captures (slice-step passes, potential kernels)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench" ; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench.log
echo "== bench ref" ; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
echo "== microbench"; timeout 600 python tools/microbench_passes.py 256 64 48 96 192 2>&1 | tee gpurun_out/micro_256.log
timeout 600 python tools/microbench_passes.py 512 32 12 24 48 2>&1 | tee gpurun_out/micro_512.log
timeout 600 python tools/microbench_passes.py 1024 16 6 12 2>&1 | tee gpurun_out/micro_1024.log
echo "== ncu launches (96 frames)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_r1c.csv \
    python bench.py --frames 96 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
tail -c 300 gpurun_out/ncu_launches_run.log
echo "== ncu full potential kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psb_kernel -s 3 -c 3 -o gpurun_out/prof_potential_r1c \
    python bench.py --frames 96 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full_run.log 2>&1
echo "== ncu full slice-step kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:psb_kernel -s 20 -c 2 -o gpurun_out/prof_slice_step_r1c \
    python tools/microbench_passes.py 256 16 96 > gpurun_out/ncu_full_run2.log 2>&1
ls -la gpurun_out
