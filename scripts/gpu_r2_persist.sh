#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2e_pytest.txt
{
echo "== chain, sell"; BTFEM_PERSIST=0 BTFEM_NO_STREAM_KERNEL=1 timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
echo "== chain, stream"; BTFEM_PERSIST=0 timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
echo "== persistent"; timeout 200 python scripts/spmv_quick.py 2>&1 | tail -1
} | tee gpurun_out/r2e_ab.txt
