#!/bin/bash
set -x
mkdir -p gpurun_out
{
echo "== sell (register-staged)"; BTFEM_NO_STREAM_KERNEL=1 timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
echo "== stream, 1 block/SM, ring depth 4"; timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
echo "== stream, 2 blocks/SM, ring depth 2"; BTFEM_PS_BPS=2 timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
} | tee gpurun_out/r2d_ab.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
