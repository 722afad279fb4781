#!/bin/bash
# session 4, call j: row pass with two transmission landing buffers (12 warps) against one (16 warps)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "" vt "" vt; do
  if [ -n "$v" ]; then export PSB_VARIANT_LIB=pyslice_b200/libpsb_$v.so; else unset PSB_VARIANT_LIB; fi
  echo "== variant '$v'" | tee -a gpurun_out/s4j_micro.log
  PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 100 127 148 2>&1 | grep fused | tee -a gpurun_out/s4j_micro.log
  PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | grep fused | tee -a gpurun_out/s4j_micro.log
done
export PSB_VARIANT_LIB=pyslice_b200/libpsb_vt.so
echo "== slice-step parity with the variant"; timeout 600 python tools/run_variant.py -m pytest tests/test_gpu_parity.py -q -x -k "fused_slice_step" 2>&1 | tail -3
