#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/hardi_n1.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/hardi_n2.txt
