#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/spmv_big.py 158 2>&1 | tee gpurun_out/spmv_big.txt
timeout 900 python scripts/ecs_bench.py 400 2>&1 | grep -E "^mesh|^b=" | tee gpurun_out/ecs_bench.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -x -k "layered_2c or box_1c_px or batched or gmres_restart" 2>&1 | tail -12 | tee gpurun_out/sanitizer_memcheck.txt
