#!/bin/bash
# quick GPU loop: parity tests, SpMV variant sweep, bench without the CPU leg
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python scripts/lanes_sweep.py 78 2>&1 | tee gpurun_out/lanes_sweep.txt
timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 3500 gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
