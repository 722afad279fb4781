#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== fused parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or si_c1" 2>&1 | tail -5 | tee gpurun_out/pytest_fused.log
echo "== batch sweep"; PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 74 111 148 185 2>&1 | tee gpurun_out/micro_256_sweep.log
PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 23 37 2>&1 | tee gpurun_out/micro_512_sweep.log
