#!/bin/bash
set -x
mkdir -p gpurun_out
BTFEM_LOOP=host timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none --cache-control none -k regex:k_hb_ -s 600 -c 40 --csv --log-file gpurun_out/r2q_hb_kernels.csv python scripts/hardi_bench.py 4 16 > gpurun_out/r2q_hb.log 2>&1
{
for b in 8 16 64; do echo "== interleaved, batch $b"; timeout 300 python scripts/hardi_bench.py 64 $b 2>&1 | grep -E "HARDI|rror"; done
} | tee gpurun_out/r2n_hardi_layouts2.txt
