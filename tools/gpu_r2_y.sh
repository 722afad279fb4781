#!/bin/bash
# round 2, call y: 256-point column pass as 128-thread CTAs with 8-column tiles, four per SM (c128) against the default
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2y
for rep in 1 2; do
for lib in "" pyslice_b200/libpsb_c128.so; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 256 64 127 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_PHASE=1 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 16 64 64 2>&1 | grep "level" | tail -1 | tee -a gpurun_out/${T}_micro.log
done; done
echo "== parity of the variant"; PSB_VARIANT_LIB=pyslice_b200/libpsb_c128.so timeout 900 python tools/run_variant.py -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_slice_step or phase_stack or full_size_properties_c2 or potential" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_variant.log
