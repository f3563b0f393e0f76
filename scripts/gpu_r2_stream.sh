#!/bin/bash
# Round 2: the TMA warp-ring SpMV (k_spmv_stream) against the register-staged SELL kernel: parity suite + quick bench A/B.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.txt
timeout 300 python bench.py --no-hardi --no-cpu --steps 2 --warmup 1 > gpurun_out/r2b_bench_stream.json 2> gpurun_out/r2b_bench_stream.err
tail -c 1800 gpurun_out/r2b_bench_stream.json; tail -3 gpurun_out/r2b_bench_stream.err
BTFEM_NO_STREAM_KERNEL=1 timeout 300 python bench.py --no-hardi --no-cpu --steps 2 --warmup 1 > gpurun_out/r2b_bench_sell.json 2> gpurun_out/r2b_bench_sell.err
tail -c 1800 gpurun_out/r2b_bench_sell.json; tail -3 gpurun_out/r2b_bench_sell.err
