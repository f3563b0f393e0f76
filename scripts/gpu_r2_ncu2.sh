#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bicgstab_persistent -c 1 -f -o gpurun_out/r2ad_prof_persistent python scripts/spmv_quick.py 78 0 > gpurun_out/r2ad_ncu.log 2>&1
tail -2 gpurun_out/r2ad_ncu.log
