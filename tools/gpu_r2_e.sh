#!/bin/bash
# round 2, call e (2 GPUs): N-GPU == 1-GPU over NCCL, bench on 2 ranks (C2 weak, C4 strong), TACAW tile-size knob
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2e
nvidia-smi -L | tee gpurun_out/${T}_host.log
echo "== nccl parity test"; timeout 900 python -m pytest tests/test_gpu_nccl.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_nccl.log
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
echo "== bench 2 gpus c2"; timeout 900 $RUN bench.py --gpus 2 --steps 4 --warmup 3 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_2gpu.log
echo "== bench 2 gpus c4 (400 frames, strong)"; timeout 900 $RUN bench.py --gpus 2 --workload c4 --frames 400 --steps 2 --warmup 2 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_2gpu_c4.log
echo "== bench 2 gpus c3 (20 frames per GPU)"; timeout 900 $RUN bench.py --gpus 2 --workload c3 --frames 20 --steps 1 --warmup 1 2>&1 | grep '^{' | tail -1 | tee gpurun_out/${T}_bench_2gpu_c3.log
echo "== tacaw tile sizes"
for kb in 96 48 32 16; do echo "tile_kb=$kb"; PSB_TACAW_TILE_KB=$kb timeout 300 python tools/microbench_tacaw.py 2>&1 | grep "level 1"; done | tee gpurun_out/${T}_tacaw_tiles.log
ls -la gpurun_out | grep ${T}
