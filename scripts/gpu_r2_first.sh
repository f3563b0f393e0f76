#!/bin/bash
# Round 2, first GPU call (1 GPU): strong periodic + pending tests on hardware, whole suite, streaming ceilings, baseline bench.
set -x
mkdir -p gpurun_out
BTFEM_TEST_STRONG=1 BTFEM_STRONG=1 timeout 300 python -m pytest tests/test_gpu_strong.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2a_strong.txt
BTFEM_TEST_PENDING=1 timeout 300 python -m pytest tests/test_gpu_pending.py -m gpu -q 2>&1 | tail -20 | tee gpurun_out/r2a_pending.txt
timeout 120 ./scripts/membench2 2>&1 | tee gpurun_out/r2a_membench2.txt
timeout 120 ./scripts/membench 2>&1 | tee gpurun_out/r2a_membench.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2a_pytest.txt
timeout 600 python bench.py --cpu-sample-steps 8 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -c 1500 gpurun_out/r2a_bench.json; tail -3 gpurun_out/r2a_bench.err
