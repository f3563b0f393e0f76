#!/bin/bash
# Round-2 state on one GPU: parity suite, both bench arms, smoke()
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2t_pytest.txt
timeout 300 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2t_smoke.txt
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 1200 gpurun_out/r2t_bench.json; tail -3 gpurun_out/r2t_bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2t_bench_reference.json 2> gpurun_out/r2t_bench_reference.err
tail -c 1500 gpurun_out/r2t_bench_reference.json; tail -3 gpurun_out/r2t_bench_reference.err
