#!/bin/bash
mkdir -p gpurun_out
{
BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2
timeout 200 python scripts/spmv_quick.py 78 1 2>&1 | tail -1
} | tee gpurun_out/r2w_ll_reduce.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py tests/test_gpu_2d.py -m gpu -q -x 2>&1 | tail -4
