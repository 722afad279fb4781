#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "potential or small64 or hbn or si_c1" 2>&1 | tail -5 | tee gpurun_out/pytest_potential.log
echo "== potential microbench"; timeout 600 python tools/microbench_potential.py 32 64 2>&1 | tee gpurun_out/micro_potential.log
