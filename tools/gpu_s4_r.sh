#!/bin/bash
# session 4, call r: 8-GPU bench (frame sharding + NCCL all-to-all to kx rows + row-sharded tiled time transform)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/s4r_bench_${N}gpu.log
