#!/bin/bash
set -x
mkdir -p gpurun_out
BTFEM_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_driver.py -x -q --timeout 400 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -8 | tee gpurun_out/part1_pytest.txt
BTFEM_DEBUG=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/part1_bench.json 2> gpurun_out/part1_bench.err; tail -c 3000 gpurun_out/part1_bench.json; tail -5 gpurun_out/part1_bench.err
