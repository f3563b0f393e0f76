#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -s 400 -c 60 --csv --log-file gpurun_out/insitu2.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_insitu2.log 2>&1
BTFEM_L2_PERSIST=0 timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -s 400 -c 60 --csv --log-file gpurun_out/insitu2_nopersist.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_insitu2b.log 2>&1
python -c "
import ctypes
cudart=ctypes.CDLL('libcudart.so')
" 2>/dev/null
nvidia-smi --query-gpu=name,l2_cache_size --format=csv 2>/dev/null || true
