#!/bin/bash
# in-situ kernel times + DRAM traffic inside the solve (caches NOT flushed between launches), then the usual quick loop
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 600 ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct -s 400 -c 60 --csv --log-file gpurun_out/insitu.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_insitu.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 3500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
