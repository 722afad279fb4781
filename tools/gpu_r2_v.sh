#!/bin/bash
# round 2, call v: evidence of the final state -- launch list of the bench command, full-set captures of the slice-step,
# potential and NUFFT kernels, live DRAM traffic at 1024 points, bench lines
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2v
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.log
echo "== ncu launch list (127 frames of C2, one step)"
PSB_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --frames 127 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launches_run.log 2>&1
echo "== ncu full: slice step 256 (phase), warm L2"
PSB_GRAPHS=0 PSB_AB=0 PSB_PHASE=1 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:fast_ -s 40 -c 2 -o gpurun_out/${T}_prof_slice_step_256 \
    python tools/microbench_passes.py 256 32 127 > gpurun_out/${T}_ncu_full_run1.log 2>&1
echo "== ncu full: slice step 512 (complex), warm L2"
PSB_GRAPHS=0 PSB_AB=0 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:fast_ -s 40 -c 2 -o gpurun_out/${T}_prof_slice_step_512 \
    python tools/microbench_passes.py 512 32 37 > gpurun_out/${T}_ncu_full_run2.log 2>&1
echo "== ncu full: slice step 1024 (phase), warm L2"
PSB_GRAPHS=0 PSB_AB=0 PSB_PHASE=1 timeout 900 ncu --set full --clock-control none --cache-control none --import-source on -k regex:fast_ -s 40 -c 2 -o gpurun_out/${T}_prof_slice_step_1024 \
    python tools/microbench_passes.py 1024 16 9 > gpurun_out/${T}_ncu_full_run3.log 2>&1
echo "== ncu full: tacaw"
PSB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tacaw_fast -s 6 -c 1 -o gpurun_out/${T}_prof_tacaw \
    python tools/microbench_tacaw.py > gpurun_out/${T}_ncu_full_run4.log 2>&1
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${T}_bench_reference.log
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c3.log
echo "== bench c1"; timeout 900 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c1.log
ls -la gpurun_out | grep ${T}
