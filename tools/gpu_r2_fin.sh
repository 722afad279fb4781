#!/bin/bash
# round 2, final call: whole GPU suite, bench lines of every workload and of the reference arm, launch list of the bench
# command, full-set capture of the TACAW kernel in its final form
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2fin
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${T}_bench_reference.log
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c3.log
echo "== bench c4 250"; timeout 900 python bench.py --workload c4 --frames 250 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c4_250.log
echo "== bench c5"; timeout 900 python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c5.log
echo "== bench c1"; timeout 900 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c1.log
echo "== ncu launch list (127 frames of C2, one step)"
PSB_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --frames 127 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu_launches_run.log 2>&1
echo "== ncu full: tacaw (final)"
PSB_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tacaw_fast -s 6 -c 1 -o gpurun_out/${T}_prof_tacaw \
    python tools/microbench_tacaw.py > gpurun_out/${T}_ncu_full_tacaw.log 2>&1
grep "level 1" gpurun_out/${T}_ncu_full_tacaw.log | head -4
ls -la gpurun_out | grep ${T}
