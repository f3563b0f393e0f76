#!/bin/bash
mkdir -p gpurun_out
{
for b in 12 20 16; do
echo "== coop batch kernel, batch $b"
timeout 150 python scripts/hardi_bench.py 64 $b 2>&1 | grep -E "HARDI|rror" | tail -3
done
} > gpurun_out/r2at_cbatch_sizes2.txt 2>&1
cat gpurun_out/r2at_cbatch_sizes2.txt
