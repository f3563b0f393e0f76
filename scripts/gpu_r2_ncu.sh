#!/bin/bash
# Round 2 ncu evidence (1 GPU): launch list of one solve, full captures of the persistent kernel and of the TMA SpMV
set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2m_launches.csv python scripts/spmv_quick.py 78 0 > gpurun_out/r2m_launches.log 2>&1
tail -2 gpurun_out/r2m_launches.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bicgstab_persistent -c 1 -f -o gpurun_out/r2m_prof_persistent python scripts/spmv_quick.py 78 0 > gpurun_out/r2m_ncu_persistent.log 2>&1
tail -2 gpurun_out/r2m_ncu_persistent.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_stream -s 5 -c 2 -f -o gpurun_out/r2m_prof_stream python scripts/spmv_quick.py 78 0 > gpurun_out/r2m_ncu_stream.log 2>&1
tail -2 gpurun_out/r2m_ncu_stream.log
ls -la gpurun_out/r2m*
