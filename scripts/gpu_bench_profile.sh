#!/bin/bash
# Run on the GPU box (via gpurun): parity tests, bench line (both arms), ncu launch list, one full ncu capture.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
timeout 900 python bench.py --cpu-sample-steps 16 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 4500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-sample-steps 8 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 1500 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_sell -s 40 -c 2 -o gpurun_out/prof_spmv_sell python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
