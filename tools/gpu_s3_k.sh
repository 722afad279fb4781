#!/bin/bash
# per-kernel view of the potential chain: launch list (small) + full-set capture of sf_tiles and the transmit pass
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 120 --csv --log-file gpurun_out/launches_pot_r1f.csv \
    python tools/microbench_potential.py 4 64 > gpurun_out/ncu_pot_run.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sf_tiles|fast_rows_kernel" -s 12 -c 2 -o gpurun_out/prof_pot_r1f \
    python tools/microbench_potential.py 4 64 > gpurun_out/ncu_pot_full.log 2>&1
tail -2 gpurun_out/ncu_pot_full.log
