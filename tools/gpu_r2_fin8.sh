#!/bin/bash
# round 2, final call on 8 GPUs: C4 strong scaling (2000 frames in all) and the C2 weak-scaling line, final kernels
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2fin8
echo "== bench c4, 8 GPUs, 2000 frames in all"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_c4_8gpu.log
echo "== bench c2, 8 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_c2_8gpu.log
