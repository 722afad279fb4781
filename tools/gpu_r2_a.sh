#!/bin/bash
# round 2, call a: validation of the new parity tests / bench arms + first experiments (20-warp row pass, DSMEM bandwidth,
# live DRAM traffic of the slice step with a warm L2)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2a
nproc > gpurun_out/${T}_host.log; nvidia-smi -L >> gpurun_out/${T}_host.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
echo "== new gpu tests"; timeout 1200 python -m pytest tests/test_gpu_config_scale.py tests/test_gpu_nccl.py -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/${T}_pytest_new.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_config_scale.py 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
echo "== microbench: default vs 20-warp row pass"
for lib in "" pyslice_b200/libpsb_w20.so; do
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 256 64 127 148 2>&1 | tee -a gpurun_out/${T}_micro.log
  PSB_VARIANT_LIB=$lib PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | tee -a gpurun_out/${T}_micro.log
done
echo "== dsmem"; nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dsmem_bw tools/ubench/dsmem_bw.cu && timeout 120 /tmp/dsmem_bw 2>&1 | tee gpurun_out/${T}_dsmem.log
echo "== bench default"; timeout 900 python bench.py --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | tee gpurun_out/${T}_bench_reference.log
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c3.log
echo "== live traffic (warm L2, no cache control): one batch of 127 images through 32 slices"
PSB_AB=0 PSB_PHASE=1 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --cache-control none --clock-control none \
    -k regex:fast_ -s 40 -c 40 --csv --log-file gpurun_out/${T}_live_traffic_256.csv python tools/microbench_passes.py 256 32 127 > gpurun_out/${T}_ncu_live_run.log 2>&1
PSB_AB=0 timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --cache-control none --clock-control none \
    -k regex:fast_ -s 40 -c 40 --csv --log-file gpurun_out/${T}_live_traffic_512.csv python tools/microbench_passes.py 512 32 37 > gpurun_out/${T}_ncu_live_run512.log 2>&1
ls -la gpurun_out | grep ${T}
