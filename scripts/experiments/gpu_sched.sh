#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 2>&1 | tail -6 | tee gpurun_out/sched_pytest.txt
python scripts/spmv_sizes.py | tee gpurun_out/spmv_sizes_sched.txt
BTFEM_NO_SCHED=1 python scripts/spmv_sizes.py | tee gpurun_out/spmv_sizes_roundrobin.txt
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/sched_bench.json 2> gpurun_out/sched_bench.err; tail -c 3000 gpurun_out/sched_bench.json; tail -3 gpurun_out/sched_bench.err
BTFEM_NO_SCHED=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu 2>/dev/null | cut -c1-330
