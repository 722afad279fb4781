#!/bin/bash
# round 2, call ai: exp(i*phase) of the phase-format row pass two pixels at a time (default) against one (pc0)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2ai
for rep in 1 2; do
for lib in pyslice_b200/libpsb_pc0.so ""; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 256 64 127 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 1024 16 9 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
done; done
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_slice_step or phase_stack or full_size_properties_c2 or graph" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest.log
