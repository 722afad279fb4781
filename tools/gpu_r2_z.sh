#!/bin/bash
# round 2, call z: TACAW time transform -- radix-10 / 20 prime-factor stages; the persistent tensor-copy variant
# (PSB_TACAW_TMA=1) against the first-stage-from-global kernel (default)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2z
for v in 0 1; do
  echo "### radix 10 (20) stages, PSB_TACAW_TMA=$v" | tee -a gpurun_out/${T}_tacaw_radix10.log
  PSB_TACAW_TMA=$v timeout 200 python tools/microbench_tacaw.py 2>&1 | grep "level 1" | tee -a gpurun_out/${T}_tacaw_radix10.log
done
echo "== parity"; timeout 600 python -m pytest tests -q -m gpu -x -k "tacaw" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_tacaw4.log
PSB_TACAW_TMA=1 timeout 600 python -m pytest tests -q -m gpu -x -k "tacaw" 2>&1 | tail -3 | tee -a gpurun_out/${T}_pytest_tacaw4.log
