#!/bin/bash
# round 2, final call on 2 GPUs: NCCL parity tests, C2 weak-scaling line, C4 strong-scaling line with 500 frames in all
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2fin2
echo "== nccl tests"; timeout 900 python -m pytest tests/test_gpu_nccl.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_nccl.log
echo "== bench c2, 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_c2_2gpu.log
echo "== bench c4, 2 GPUs, 500 frames in all"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --workload c4 --frames 500 --steps 2 --warmup 3 --no-cpu-baseline 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_c4_2gpu.log
