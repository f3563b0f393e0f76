#!/bin/bash
set -x
mkdir -p gpurun_out
{
echo "== persistent"; timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
echo "== persistent, profiled"; BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
} | tee gpurun_out/r2f_ab.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -m gpu -q -x 2>&1 | tail -5
