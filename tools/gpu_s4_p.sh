#!/bin/bash
# session 4, call p: tiled mixed-radix TACAW kernel: parity, then timings at configuration scale
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== tacaw parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "tacaw or small64 or parseval" 2>&1 | tail -8 | tee gpurun_out/s4p_pytest_tacaw.log
echo "== microbench"; timeout 600 python tools/microbench_tacaw.py 2>&1 | tee gpurun_out/s4p_micro_tacaw.log
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/s4p_pytest_gpu.log
