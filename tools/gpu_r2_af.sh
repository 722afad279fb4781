#!/bin/bash
# round 2, call af: potential build with the inverse passes' copies two tiles / units ahead (default) against one (inv1buf)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2af
for rep in 1 2; do
for lib in pyslice_b200/libpsb_inv1buf.so ""; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_potential.log
  PSB_VARIANT_LIB=$lib PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 16 64 64 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_potential.log
  PSB_VARIANT_LIB=$lib PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 16 64 64 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_potential.log
  PSB_VARIANT_LIB=$lib PSB_GEOM=c4 PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 72 72 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_potential.log
done; done
echo "== parity"; timeout 900 python -m pytest tests -q -m gpu -x -k "potential or binning or golden or recipe or nufft or c3 or c4 or c5" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
