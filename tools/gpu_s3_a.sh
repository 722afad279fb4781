#!/bin/bash
# session-3 run A: hardware microbenchmarks (L2 bandwidth, fp32x2 issue), occupancy variant of the line passes,
# GPU parity suite, bench
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== l2bw"; timeout 120 tools/ubench/build/l2bw 2>&1 | tee gpurun_out/l2bw.log
echo "== fp32x2"; timeout 60 tools/ubench/build/fp32x2 2>&1 | tee gpurun_out/fp32x2.log
echo "== microbench base"; timeout 300 python tools/microbench_passes.py 256 64 96 2>&1 | tee gpurun_out/micro_256_base.log
timeout 300 python tools/microbench_passes.py 512 32 24 2>&1 | tee gpurun_out/micro_512_base.log
echo "== microbench minblocks=3"; PSB_VARIANT_LIB=pyslice_b200/libpsb_mb3.so timeout 300 python tools/microbench_passes.py 256 64 96 2>&1 | tee gpurun_out/micro_256_mb3.log
PSB_VARIANT_LIB=pyslice_b200/libpsb_mb3.so timeout 300 python tools/microbench_passes.py 512 32 24 2>&1 | tee gpurun_out/micro_512_mb3.log
echo "== pytest" ; timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps 2 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench.log
