#!/bin/bash
# HARDI sweep (configs[4]) on 1 GPU: batch layouts and batch sizes
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "batch" 2>&1 | tail -5
{
echo "== member layout (round 1 default), batch 16"; BTFEM_BATCH_LAYOUT=member timeout 300 python scripts/hardi_bench.py 64 16 2>&1 | grep HARDI
for b in 16 32 64; do echo "== interleaved, batch $b"; timeout 300 python scripts/hardi_bench.py 64 $b 2>&1 | grep -E "HARDI|rror"; done
} | tee gpurun_out/r2n_hardi_layouts.txt
