#!/bin/bash
# session 4, call d: full/empty ring variants of the fused structure-factor kernel, L2 prefetch in the row pass, HBM by direction
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== hbm by direction"; timeout 120 python tools/ubench/hbm_rw.py 2>&1 | tee gpurun_out/s4d_hbm_rw.log
echo "== potential parity"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "potential" 2>&1 | tail -3 | tee gpurun_out/s4d_pytest_potential.log
for v in "" vb vc vd ve; do
  echo "== potential microbench variant '$v'"
  if [ -n "$v" ]; then export PSB_VARIANT_LIB=pyslice_b200/libpsb_$v.so; else unset PSB_VARIANT_LIB; fi
  PSB_LEVELS=2 timeout 300 python tools/microbench_potential.py 32 64 2>&1 | grep -E "level|variant" | tee -a gpurun_out/s4d_micro_pot.log
done
unset PSB_VARIANT_LIB
PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 32 64 2>&1 | grep -E "level" | tee -a gpurun_out/s4d_micro_pot.log
echo "== slice-step microbench (L2 prefetch of t)"; PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 100 148 2>&1 | tee gpurun_out/s4d_micro_256.log
PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | tee gpurun_out/s4d_micro_512.log
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/s4d_pytest_gpu.log
