#!/bin/bash
# session 4, call z: validation + evidence of the round's final state (phase format on the bench path)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/s4z_smoke.log
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/s4z_pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/s4z_bench.log
echo "== bench again (run-to-run)"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/s4z_bench_b.log
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | tee gpurun_out/s4z_bench_reference.log
echo "== microbenchmarks"
PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 256 64 127 2>&1 | tee gpurun_out/s4z_micro.log
PSB_AB=0 timeout 300 python tools/microbench_passes.py 256 64 127 2>&1 | tee -a gpurun_out/s4z_micro.log
PSB_AB=0 PSB_PHASE=1 timeout 300 python tools/microbench_passes.py 512 32 37 2>&1 | tee -a gpurun_out/s4z_micro.log
PSB_LEVELS=1 PSB_PHASE=1 timeout 300 python tools/microbench_potential.py 100 64 2>&1 | grep -E "level" | tee -a gpurun_out/s4z_micro.log
PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 100 64 2>&1 | grep -E "level" | tee -a gpurun_out/s4z_micro.log
echo "== ncu launch list (127 frames, one step)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/s4z_launches.csv \
    python bench.py --frames 127 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/s4z_ncu_launches_run.log 2>&1
echo "== ncu full: slice step (phase format)"
PSB_AB=0 PSB_PHASE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fast_ -s 40 -c 2 -o gpurun_out/s4z_prof_slice_step \
    python tools/microbench_passes.py 256 32 127 > gpurun_out/s4z_ncu_full_run1.log 2>&1
echo "== ncu full: potential chain (phase format)"
PSB_LEVELS=1 PSB_PHASE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sf_tiles|fast_|phase_tables' -s 40 -c 4 -o gpurun_out/s4z_prof_potential \
    python tools/microbench_potential.py 16 64 > gpurun_out/s4z_ncu_full_run2.log 2>&1
ls -la gpurun_out | grep s4z
