#!/bin/bash
# 2 GPUs of one box: the bench line under torchrun (sweep sharding + HARDI + the row-partitioned figure)
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err
tail -c 3500 gpurun_out/r2r_bench_n2.json; tail -5 gpurun_out/r2r_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_partition.py -m gpu -q 2>&1 | tail -4
