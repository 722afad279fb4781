#!/bin/bash
# round 2, call ad: slice-step throughput against the number of images in the batch (L2 residency of psi against whole rounds)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2ad
for F in 185 222 250 296 500; do PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 256 32 $F 2>&1 | grep "n=" | tee -a gpurun_out/${T}_batch_sweep_large.log; done
for F in 55 74 111 148; do PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 16 $F 2>&1 | grep "n=" | tee -a gpurun_out/${T}_batch_sweep_large.log; done
for F in 16 24 37; do PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 1024 8 $F 2>&1 | grep "n=" | tee -a gpurun_out/${T}_batch_sweep_large.log; done
