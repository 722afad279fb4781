#!/bin/bash
# session 4, call l: ring depth of the structure-factor tiles (2 / 3 / 4 stages) at C2 and C4 geometry
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== potential parity (4 stages)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "potential" 2>&1 | tail -3 | tee gpurun_out/s4l_pytest_potential.log
for v in v2 v3 "" v2 ""; do
  if [ -n "$v" ]; then export PSB_VARIANT_LIB=pyslice_b200/libpsb_$v.so; else unset PSB_VARIANT_LIB; fi
  echo "== stages variant '$v' ('' = 4)" | tee -a gpurun_out/s4l_micro.log
  PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 100 64 2>&1 | grep level | tee -a gpurun_out/s4l_micro.log
  PSB_GEOM=c4 PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 8 64 2>&1 | grep level | tee -a gpurun_out/s4l_micro.log
done
