#!/bin/bash
# round 2, call aq: NUFFT gather with two rows per lane (default) against one (rpl1)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2ar
echo "== parity"; timeout 900 python -m pytest tests -q -m gpu -x -k "nufft or c4 or potential" 2>&1 | tail -2 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do
for lib in pyslice_b200/libpsb_rpl1.so ""; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_potential_c4.log
  PSB_VARIANT_LIB=$lib PSB_GEOM=c4 PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 72 72 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_potential_c4.log
done; done
