#!/bin/bash
set -x
mkdir -p gpurun_out
{
echo "== persistent (8,4,2)"; BTFEM_PS_DEEP=0 BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
echo "== persistent (8,5,3)"; BTFEM_PROFILE_PERSIST=1 BTFEM_PROFILE_PERSIST_FILE=gpurun_out/r2i_warp_times.txt timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
echo "== persistent (8,5,3) unprofiled"; timeout 200 python scripts/spmv_quick.py 78 1 2>&1 | tail -1
} | tee gpurun_out/r2i_deep.txt
