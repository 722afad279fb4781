#!/bin/bash
# round 2, call d: column-pass tile width at 512 points (8 columns x 2 CTAs vs 16 columns x 1 CTA), TACAW first-stage unroll
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2d
for lib in "" pyslice_b200/libpsb_c512.so; do
  PSB_VARIANT_LIB=$lib PSB_AB=0 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | tee -a gpurun_out/${T}_micro.log
done
echo "== tacaw"; timeout 600 python tools/microbench_tacaw.py 2>&1 | tee gpurun_out/${T}_tacaw.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "tacaw" 2>&1 | tail -3 | tee -a gpurun_out/${T}_tacaw.log
echo "== ncu per-kernel at 512 (variant)"
PSB_VARIANT_LIB=pyslice_b200/libpsb_c512.so PSB_GRAPHS=0 PSB_AB=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'fast_' -s 60 -c 40 --csv \
   --log-file gpurun_out/${T}_launches_512_c512.csv python tools/microbench_passes.py 512 32 37 > gpurun_out/${T}_ncu_run.log 2>&1
ls -la gpurun_out | grep ${T}
