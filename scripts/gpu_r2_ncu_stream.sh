#!/bin/bash
# ncu --set full (source-level stall reasons) of the TMA warp-ring SpMV inside a short bench run
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 30 -c 2 -f -o gpurun_out/r2c_prof_stream python bench.py --no-hardi --no-cpu --steps 1 --warmup 1 > gpurun_out/r2c_ncu.log 2>&1
tail -5 gpurun_out/r2c_ncu.log
ls -la gpurun_out/*.ncu-rep
