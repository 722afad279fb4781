#!/bin/bash
# round 2, call ah: persistent structure-factor tiles (default) against one CTA per tile and pair (sfnp)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2ah
echo "== parity first"; timeout 900 python -m pytest tests -q -m gpu -x -k "potential or binning or golden or recipe or nufft or c3 or c5 or layers" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
for rep in 1 2; do
for lib in pyslice_b200/libpsb_sfnp.so ""; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_potential.log
  PSB_VARIANT_LIB=$lib PSB_PHASE=1 PSB_LEVELS=1 PSB_CHUNK_PAIRS=1 timeout 300 python tools/microbench_potential.py 16 128 128 74 111 148 2>&1 | grep "level" | tail -4 | tee -a gpurun_out/${T}_potential.log
  PSB_VARIANT_LIB=$lib PSB_GEOM=c4 PSB_SF_MODE=1 PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 72 72 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_potential.log
done; done
