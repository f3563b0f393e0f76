#!/bin/bash
# 8 GPUs of one box: configs[4] HARDI sweep (64 directions x 4 b), balanced direction-major sharding.
set -x
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29811 scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/hardi8b_w8_b16.txt
