#!/bin/bash
# Run on the GPU box (via gpurun): bench line, lanes sweep, ncu launch list, one full ncu capture.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 3 --cpu-sample-steps 8 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python scripts/lanes_sweep.py 78 > gpurun_out/lanes_sweep.txt 2>&1; cat gpurun_out/lanes_sweep.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 40 -c 3 -o gpurun_out/prof_spmv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
