#!/bin/bash
# session 4, call t: column pass with one exchange buffer and 3 CTAs per SM (vx) against two buffers and 2 CTAs ('')
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "" vx "" vx; do
  if [ -n "$v" ]; then export PSB_VARIANT_LIB=pyslice_b200/libpsb_$v.so; else unset PSB_VARIANT_LIB; fi
  echo "== variant '$v'" | tee -a gpurun_out/s4t_micro.log
  PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 111 127 148 2>&1 | grep fused | tee -a gpurun_out/s4t_micro.log
  PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | grep fused | tee -a gpurun_out/s4t_micro.log
  PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 100 64 2>&1 | grep level | tee -a gpurun_out/s4t_micro.log
done
export PSB_VARIANT_LIB=pyslice_b200/libpsb_vx.so
echo "== parity with the variant"; timeout 600 python tools/run_variant.py -m pytest tests/test_gpu_parity.py -q -x -k "fused_slice_step or potential_pipelined or full_size" 2>&1 | tail -3 | tee -a gpurun_out/s4t_micro.log
