#!/bin/bash
# launch list of one 125-frame batch of the C2 workload (ncu, per-launch durations)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r1e.csv \
    python bench.py --frames 125 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launches_run.log 2>&1
tail -c 400 gpurun_out/ncu_launches_run.log
