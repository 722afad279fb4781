#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2u
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${T}_smoke.log
echo "== e2e diag c2"; timeout 600 python tools/diag_e2e.py c2 2>&1 | tail -7 | tee gpurun_out/${T}_diag_e2e.log
echo "== e2e diag c4 250"; timeout 600 python tools/diag_e2e.py c4 250 2>&1 | tail -7 | tee -a gpurun_out/${T}_diag_e2e.log
echo "== bench default"; timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
