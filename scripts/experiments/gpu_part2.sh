#!/bin/bash
# two GPUs: partition tests (threads + two processes over CUDA IPC), then the partitioned 1 M-DOF solve vs one GPU
set -x
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_gpu_partition.py -x -q --timeout 400 2>&1 | tail -25 | tee gpurun_out/part2_pytest.txt
for n in 78; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  scripts/dist_solve.py --nbox $n --check --reps 2 > gpurun_out/part2_dist_n${n}.json 2> gpurun_out/part2_dist_n${n}.err
tail -c 2500 gpurun_out/part2_dist_n${n}.json; tail -5 gpurun_out/part2_dist_n${n}.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 \
  scripts/dist_solve.py --nbox 124 --check --reps 1 > gpurun_out/part2_dist_n124.json 2> gpurun_out/part2_dist_n124.err
tail -c 2500 gpurun_out/part2_dist_n124.json; tail -5 gpurun_out/part2_dist_n124.err
