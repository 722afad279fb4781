#!/bin/bash
# session 4, call i: one-GPU timings of configs C1, C3, C5, C4 (bounded) + 2-GPU bench (all-to-all path)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for c in c1 c3 c5 c4; do
  echo "== $c"; timeout 600 python tools/config_sweep.py $c 2>&1 | tail -1 | tee -a gpurun_out/s4i_sweep.log
done
