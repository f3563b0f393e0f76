#!/bin/bash
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/gather_bench.cu -o /tmp/gather_bench 2>&1 | grep -v warning | head -5
timeout 120 /tmp/gather_bench > gpurun_out/r2bb_gather_bench.txt 2>&1
cat gpurun_out/r2bb_gather_bench.txt
