#!/bin/bash
# 1 GPU: the round-end checks (parity tests, both bench arms) + HARDI N=1 + where the e2e time goes.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/final_pytest.txt
timeout 400 python bench.py --cpu-sample-steps 8 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
tail -c 600 gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-sample-steps 8 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
tail -c 400 gpurun_out/final_bench_reference.json
timeout 200 python scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/final_hardi_n1.txt
timeout 200 python scripts/e2e_breakdown.py 2>&1 | tail -3 | tee gpurun_out/final_e2e_breakdown.txt
