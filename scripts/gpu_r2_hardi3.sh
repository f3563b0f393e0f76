#!/bin/bash
mkdir -p gpurun_out
{
for b in 16 32 64; do echo "== member layout, batch $b"; timeout 300 python scripts/hardi_bench.py 64 $b 2>&1 | grep -E "HARDI|rror"; done
echo "== member layout, batch 32, 12 iterations per WHILE pass"; BTFEM_UNROLL=12 timeout 300 python scripts/hardi_bench.py 64 32 2>&1 | grep -E "HARDI|rror"
} | tee gpurun_out/r2v_hardi_batch.txt
