#!/bin/bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/diag_fast.py 2>&1 | tee gpurun_out/diag_fast.log
