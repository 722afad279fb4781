#!/bin/bash
# round 2, call b: fused 1024-point kernels (parity + speed, 512- vs 256-thread column pass), operand-bandwidth and DSMEM
# microbenchmarks, diagnosis of the between-phase stall of the bench's device arm (clock sampler on/off, per-step times)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2b
echo "== ubench"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32x2_operands tools/ubench/fp32x2_operands.cu && /tmp/fp32x2_operands 2>&1 | tee gpurun_out/${T}_fp32x2_operands.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dsmem_bw tools/ubench/dsmem_bw.cu && timeout 120 /tmp/dsmem_bw 2>&1 | tee gpurun_out/${T}_dsmem.log
echo "== 1024 parity"; timeout 1200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fused_slice_step_vs_generic or potential_pipelined or phase_stack_equals" 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_1024.log
timeout 600 python -m pytest tests/test_gpu_config_scale.py -q -m gpu -x -k "c4_grid" 2>&1 | tail -5 | tee -a gpurun_out/${T}_pytest_1024.log
echo "== microbench 1024"
for lib in "" pyslice_b200/libpsb_c256.so; do
  PSB_VARIANT_LIB=$lib PSB_AB=1 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 1024 16 9 10 2>&1 | tee -a gpurun_out/${T}_micro.log
done
PSB_GEOM=c4 PSB_LEVELS=1,0 PSB_PHASE=1 timeout 600 python tools/microbench_potential.py 4 64 2>&1 | tee -a gpurun_out/${T}_micro.log
echo "== bench c4 (250-frame share)"; timeout 900 python bench.py --workload c4 --frames 250 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
echo "== bench default: sampler on / off"
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_clocks.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-clocks 2>&1 | tail -1 | tee gpurun_out/${T}_bench_noclocks.log
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-clocks --no-e2e 2>&1 | tail -1 | tee gpurun_out/${T}_bench_noclocks_noe2e.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
ls -la gpurun_out | grep ${T}
