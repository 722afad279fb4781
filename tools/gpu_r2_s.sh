#!/bin/bash
# round 2, call s (8 GPUs): NCCL parity at world 2 and 4, C2 weak scaling and C4 strong scaling (2000 frames) at 8 and 4 ranks
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2s
nvidia-smi -L | tee gpurun_out/${T}_host.log
echo "== nccl parity tests"; timeout 900 python -m pytest tests/test_gpu_nccl.py -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_nccl.log
run() { n=$1; shift; timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n "$@" 2>&1 | grep '^{' | tail -1; }
echo "== c4 strong, 8 GPUs"; run 8 --workload c4 --steps 3 --warmup 2 | tee gpurun_out/${T}_bench_c4_8gpu.log
echo "== c4 strong, 4 GPUs"; run 4 --workload c4 --steps 2 --warmup 1 | tee gpurun_out/${T}_bench_c4_4gpu.log
echo "== c2 weak, 8 GPUs"; run 8 --steps 5 --warmup 3 | tee gpurun_out/${T}_bench_c2_8gpu.log
echo "== c3 weak, 8 GPUs (25 frames per GPU)"; run 8 --workload c3 --frames 25 --steps 1 --warmup 1 | tee gpurun_out/${T}_bench_c3_8gpu.log
ls -la gpurun_out | grep ${T}
