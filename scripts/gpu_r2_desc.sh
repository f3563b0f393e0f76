#!/bin/bash
mkdir -p gpurun_out
BTFEM_PROFILE_PERSIST=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | tail -2 | tee gpurun_out/r2ac_desc_prefetch.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
