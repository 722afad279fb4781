#!/bin/bash
# 2-GPU check: frame sharding + NCCL all-to-all + k-row sharded TACAW through bench.py, and the reference arm
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 2 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_2gpu.log
