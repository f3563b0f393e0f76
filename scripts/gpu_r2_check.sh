#!/bin/bash
# full GPU suite + the default bench line (1 GPU)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/r2l_pytest.txt
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
tail -c 3000 gpurun_out/r2l_bench.json; tail -3 gpurun_out/r2l_bench.err
timeout 200 python scripts/e2e_breakdown.py 2>&1 | tail -4 | tee gpurun_out/r2l_e2e_breakdown.txt
