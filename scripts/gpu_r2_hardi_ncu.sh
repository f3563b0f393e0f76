#!/bin/bash
set -x
mkdir -p gpurun_out
BTFEM_LOOP=host timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --cache-control none -k regex:k_hb_ -s 600 -c 60 --csv --log-file gpurun_out/r2q_hb_kernels.csv python scripts/hardi_bench.py 4 16 > gpurun_out/r2q_hb.log 2>&1
tail -2 gpurun_out/r2q_hb.log
BTFEM_LOOP=host BTFEM_BATCH_LAYOUT=member timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --cache-control none -k regex:"k_spmv_sell|k_update" -s 600 -c 60 --csv --log-file gpurun_out/r2q_member_kernels.csv python scripts/hardi_bench.py 4 16 > gpurun_out/r2q_member.log 2>&1
tail -2 gpurun_out/r2q_member.log
