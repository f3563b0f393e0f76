#!/bin/bash
mkdir -p gpurun_out
{
echo "== no proxy fence before refills"; timeout 200 python scripts/spmv_quick.py 78 1 2>&1 | tail -1
echo "== fence.proxy.async before every refill"; BTFEM_PS_FENCE=1 timeout 200 python scripts/spmv_quick.py 78 1 2>&1 | tail -1
} | tee gpurun_out/r2af_fence.txt
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
