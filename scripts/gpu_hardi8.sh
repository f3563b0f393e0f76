#!/bin/bash
# 8 GPUs of one box: configs[4] HARDI sweep (64 directions x 4 b) with the balanced sharding, and the N=1 run on the same box.
set -x
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29811 scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/hardi8_w8_b16.txt
run 29812 scripts/hardi_bench.py 64 32 2>&1 | grep HARDI | tee gpurun_out/hardi8_w8_b32.txt
timeout 300 python scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/hardi8_w1_b16.txt
