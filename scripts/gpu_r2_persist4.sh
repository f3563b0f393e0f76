#!/bin/bash
set -x
mkdir -p gpurun_out
{
for w in 8 12 16; do
echo "== persistent, $w warps/block"; BTFEM_PS_WARPS=$w BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
done
} | tee gpurun_out/r2h_warps.txt
BTFEM_PS_WARPS=12 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py tests/test_gpu_2d.py -m gpu -q -x 2>&1 | tail -4
