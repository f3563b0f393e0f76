#!/bin/bash
# N GPUs (set W): partitioned solves of the 1 M-DOF and the 4 M-DOF cell-in-box workloads, with timelines
set -x
W=${W:-4}
mkdir -p gpurun_out
for n in 78 124; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2971$W \
  scripts/dist_solve.py --nbox $n --reps 2 --trace 12000 > gpurun_out/part4_w${W}_n${n}.json 2> gpurun_out/part4_w${W}_n${n}.err
tail -c 600 gpurun_out/part4_w${W}_n${n}.json | cut -c1-600; tail -3 gpurun_out/part4_w${W}_n${n}.err
python scripts/trace_summary.py "gpurun_out/trace_w${W}_r[01].npy" | tee gpurun_out/part4_w${W}_n${n}_trace.txt
done
