#!/bin/bash
# round 2, call c: CUDA-graph replay (parity, effect on the sampler stalls), slab layout / detector-only mode on the GPU,
# per-kernel times of the 1024-point slice step
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2c
echo "== graph + adf tests"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_recipes.py -q -m gpu -x -k "graph_replay or small64_runs or test_04_haadf" 2>&1 | tail -8 | tee gpurun_out/${T}_pytest_new.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
echo "== bench default (graphs on / off)"
timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_graphs.log
PSB_GRAPHS=0 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_nographs.log
echo "== bench c4 250"; timeout 900 python bench.py --workload c4 --frames 250 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
echo "== bench c1"; timeout 900 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c1.log
echo "== ncu launch list 1024"
PSB_GRAPHS=0 PSB_AB=0 PSB_PHASE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'fast_|LinePass' -s 60 -c 60 --csv \
   --log-file gpurun_out/${T}_launches_1024.csv python tools/microbench_passes.py 1024 16 9 > gpurun_out/${T}_ncu_run.log 2>&1
PSB_GRAPHS=0 PSB_LEVELS=1 PSB_GEOM=c4 PSB_PHASE=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 30 -c 80 --csv \
   --log-file gpurun_out/${T}_launches_pot_c4.csv python tools/microbench_potential.py 2 64 > gpurun_out/${T}_ncu_run2.log 2>&1
ls -la gpurun_out | grep ${T}
