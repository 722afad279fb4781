#!/bin/bash
# round 2, call f: sanitizer pass, full GPU suite, evidence bench lines (C2 default incl. CPU arm, C3, C5)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=r2f
echo "== all gpu tests"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/${T}_pytest_gpu.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/${T}_bench.log
echo "== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${T}_bench_reference.log
echo "== bench c3"; timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c3.log
echo "== bench c5"; timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/${T}_bench_c5.log
echo "== sanitizer"; bash tools/sanitize.sh 2>&1 | tail -30 | tee gpurun_out/${T}_sanitize_summary.log
ls -la gpurun_out | grep -E "${T}|sanitize"
