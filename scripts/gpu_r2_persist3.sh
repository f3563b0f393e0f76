#!/bin/bash
set -x
mkdir -p gpurun_out
BTFEM_PROFILE_PERSIST=1 BTFEM_PROFILE_PERSIST_FILE=gpurun_out/r2g_warp_times.txt timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2 | tee gpurun_out/r2g_prof.txt
