#!/bin/bash
mkdir -p gpurun_out
{
for pad in 0 61 509 4093; do
echo "== stream start padding up to $pad x 128 B"; BTFEM_PS_PAD=$pad BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
done
} | tee gpurun_out/r2u_pad.txt
