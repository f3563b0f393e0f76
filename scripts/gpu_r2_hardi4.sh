#!/bin/bash
mkdir -p gpurun_out
{
echo "== batch 16 (member layout)"; timeout 300 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror"
for k in 8 16 24 37; do echo "== $k concurrent persistent solves on SM shares"; timeout 300 python scripts/hardi_bench.py 64 16 $k 2>&1 | grep -E "HARDI|rror|Trace" | tail -3; done
} | tee gpurun_out/r2ab_hardi_concurrent.txt
