#!/bin/bash
# round 2, call t: slice-step variants -- Px in registers (column pass), phases by streaming loads (row pass), both
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2t
for lib in "" pyslice_b200/libpsb_pxr.so pyslice_b200/libpsb_ldg.so pyslice_b200/libpsb_both.so; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 256 64 127 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 1024 16 9 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
done
echo "== parity of the combined variant"; PSB_VARIANT_LIB=pyslice_b200/libpsb_both.so timeout 900 python tools/run_variant.py -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_slice_step or phase_stack or full_size_properties_c2" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_variant.log
