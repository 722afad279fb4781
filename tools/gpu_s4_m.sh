#!/bin/bash
# session 4, call m: ncu of the structure-factor tiles in the long-segment regime (C4 geometry: ~300 atoms per slice)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
PSB_GEOM=c4 PSB_LEVELS=1 timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:'sf_tiles' -s 6 -c 2 -o gpurun_out/s4m_prof_sf_c4 \
    python tools/microbench_potential.py 2 64 > gpurun_out/s4m_run.log 2>&1
tail -3 gpurun_out/s4m_run.log
PSB_GEOM=c4 PSB_LEVELS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 200 --csv --log-file gpurun_out/s4m_launches_c4.csv \
    python tools/microbench_potential.py 2 64 > gpurun_out/s4m_run2.log 2>&1
