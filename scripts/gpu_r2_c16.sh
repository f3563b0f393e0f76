#!/bin/bash
mkdir -p gpurun_out
{
echo "== 16-bit column offsets (18 B per nonzero)"; BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
echo "== int32 columns (20 B per nonzero)"; BTFEM_PS_COL32=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -1
} | tee gpurun_out/r2aa_c16.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5
