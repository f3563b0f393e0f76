#!/bin/bash
# two GPUs: timeline of the partitioned 1 M-DOF solve + single-GPU reference points at half the size
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 \
  scripts/dist_solve.py --nbox 78 --reps 2 --trace 12000 > gpurun_out/part3_dist_n78.json 2> gpurun_out/part3.err
tail -c 1500 gpurun_out/part3_dist_n78.json; tail -3 gpurun_out/part3.err
python scripts/trace_summary.py | tee gpurun_out/part3_trace_summary.txt
# one rank = plain partition code path without peers, half-size mesh: what the local work costs
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29713 \
  scripts/dist_solve.py --nbox 62 --check --reps 2 > gpurun_out/part3_dist1_n62.json 2>> gpurun_out/part3.err
tail -c 1500 gpurun_out/part3_dist1_n62.json
