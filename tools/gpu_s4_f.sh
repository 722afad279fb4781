#!/bin/bash
# session 4, call f: which L2-policy change slows the 100-frame potential build? (A/B builds x chunk rule)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "" va vb vc; do
 for rule in 1 0; do
  if [ -n "$v" ]; then export PSB_VARIANT_LIB=pyslice_b200/libpsb_$v.so; else unset PSB_VARIANT_LIB; fi
  echo "== variant '$v' (va = plain stores, vb = no evict-first loads, vc = neither) chunk rule $rule" | tee -a gpurun_out/s4f_micro_pot.log
  PSB_CHUNK_RULE=$rule PSB_LEVELS=1 timeout 300 python tools/microbench_potential.py 100 64 80 2>&1 | grep level | tee -a gpurun_out/s4f_micro_pot.log
 done
done
