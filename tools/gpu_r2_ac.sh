#!/bin/bash
# round 2, call ac: 1024-point column pass as two 256-thread CTAs per SM working through 8-column tiles in two halves
# (halves) against one 512-thread CTA (default)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2ac
for rep in 1 2; do
for lib in "" pyslice_b200/libpsb_halves.so; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 1024 16 9 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_GEOM=c4 PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 72 72 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_micro.log
done; done
echo "== parity of the variant"; PSB_VARIANT_LIB=pyslice_b200/libpsb_halves.so timeout 900 python tools/run_variant.py -m pytest tests -q -m gpu -x -k "1024 or c4" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_variant.log
