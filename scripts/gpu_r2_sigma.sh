#!/bin/bash
# SELL sorting window against gather locality on the small HARDI mesh (chain batch of 16; 16 directions = 64 signals)
mkdir -p gpurun_out
{
for s in 1024 256 128 64 32; do
echo "== BTFEM_SELL_SIGMA=$s"
BTFEM_TIMING=1 BTFEM_SELL_SIGMA=$s timeout 150 python scripts/hardi_bench.py 16 16 2>&1 | grep -E "HARDI|rror|SELL-32" | sort -u | tail -3
done
} > gpurun_out/r2ak_sell_sigma_hardi.txt 2>&1
cat gpurun_out/r2ak_sell_sigma_hardi.txt
