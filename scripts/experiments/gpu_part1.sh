#!/bin/bash
# one GPU: all parity tests (incl. the row-partition protocol with thread ranks) + bench with the device-driven
# and the host-driven iteration loop
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -25 | tee gpurun_out/part1_pytest.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/part1_bench.json 2> gpurun_out/part1_bench.err; tail -c 3000 gpurun_out/part1_bench.json; tail -5 gpurun_out/part1_bench.err
BTFEM_LOOP=host timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/part1_bench_hostloop.json 2> gpurun_out/part1_bench.err; tail -c 3000 gpurun_out/part1_bench_hostloop.json | cut -c1-400; tail -5 gpurun_out/part1_bench.err
