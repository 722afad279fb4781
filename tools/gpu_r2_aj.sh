#!/bin/bash
# round 2, call aj: NUFFT gather on the packed pipe; the added TACAW lengths
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2aj
echo "== parity"; timeout 900 python -m pytest tests -q -m gpu -x -k "nufft or tacaw_time_fft_lengths or c4 or potential" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do
  PSB_GEOM=c4 PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 72 72 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_potential_c4.log
done
