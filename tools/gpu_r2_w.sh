#!/bin/bash
# round 2, call w: Px/Py tables of the 1024-point column pass in shared memory (A/B against the build that reads them
# through L1), TMA 2-D box microbenchmark, full-set captures of the C2 potential kernels with a warm L2
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2w
echo "== tma2d ubench"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma2d_bw tools/ubench/tma2d_bw.cu -lcuda && timeout 120 /tmp/tma2d_bw 2>&1 | tee gpurun_out/${T}_ubench_tma2d.txt
echo "== 1024 column tables: L1 (tab0) vs shared memory (default)"
for rep in 1 2; do
for lib in pyslice_b200/libpsb_tab0.so ""; do
  echo "### lib=${lib:-default}" | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 1024 16 9 2>&1 | grep "n=" | tee -a gpurun_out/${T}_micro.log
done; done
echo "== parity 1024"; timeout 900 python -m pytest tests -q -m gpu -x -k "1024 or fused_slice_step or phase_stack or c4" 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_1024.log
echo "== bench c4 250 frames"; timeout 900 python bench.py --workload c4 --frames 250 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
echo "== ncu full: C2 potential kernels, warm L2"
PSB_GRAPHS=0 PSB_PHASE=1 PSB_LEVELS=1 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'sf_tiles|fast_|phase_tables' -s 80 -c 4 -o gpurun_out/${T}_prof_potential_c2 \
    python tools/microbench_potential.py 8 64 > gpurun_out/${T}_ncu_full_run1.log 2>&1
tail -3 gpurun_out/${T}_ncu_full_run1.log
ls -la gpurun_out | grep ${T}
