#!/bin/bash
set -x
mkdir -p gpurun_out
{
for k in 0 4 8 12 16 26; do
echo "== L2 prefetch of $k pieces per warp at pass end"; BTFEM_PS_L2AHEAD=$k BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
done
} | tee gpurun_out/r2o_l2ahead.txt
