#!/bin/bash
# 8 GPUs of one box: the bench line under torchrun exactly as the driver launches it (+ the reference arm's N>1 behaviour)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2z_bench_n8.json 2> gpurun_out/r2z_bench_n8.err
tail -c 2500 gpurun_out/r2z_bench_n8.json; tail -5 gpurun_out/r2z_bench_n8.err
