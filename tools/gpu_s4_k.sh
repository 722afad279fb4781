#!/bin/bash
# session 4, call k: 2-GPU bench (frame sharding + NCCL all-to-all + row-sharded time FFT)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/s4k_bench_2gpu.log
