#!/bin/bash
# round 2, call x: potential build at C2 -- chunk (scratch) size sweep now that ncu shows the 64 MB chunk missing L2,
# inverse column pass with 3 CTAs per SM (default) against 2 (inv0)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2x
for lib in pyslice_b200/libpsb_inv0.so ""; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_potential.log
  PSB_VARIANT_LIB=$lib PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 16 16 24 32 40 48 64 2>&1 | grep "level" | tee -a gpurun_out/${T}_potential.log
done
echo "### 512 grid (C3 geometry uses complex stack)" | tee -a gpurun_out/${T}_potential.log
echo "== parity"; timeout 900 python -m pytest tests -q -m gpu -x -k "potential or binning or golden" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
