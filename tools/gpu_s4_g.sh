#!/bin/bash
# session 4, call g: the same A/B matrix inside bench.py (the potential phase regressed there but not in the microbenchmark)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { echo "== $1 rule=$2" | tee -a gpurun_out/s4g_bench.log
  if [ -n "$1" ]; then export PSB_VARIANT_LIB=pyslice_b200/libpsb_$1.so; else unset PSB_VARIANT_LIB; fi
  PSB_CHUNK_RULE=$2 timeout 300 python tools/run_variant.py bench.py --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['phases_ms_per_step'])" | tee -a gpurun_out/s4g_bench.log; }
run "" 1
run "" 0
run vc 0
run vc 1
run va 0
run vb 0
