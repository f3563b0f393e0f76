#!/bin/bash
# W GPUs: partition tests, then partitioned solves with timelines and the check against one GPU
set -x
W=${W:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_partition.py -x -q --timeout 400 2>&1 | tail -15 | tee gpurun_out/part5_pytest.txt
for n in 78 124; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2972$W \
  scripts/dist_solve.py --nbox $n --reps 2 --check --trace 12000 > gpurun_out/part5_w${W}_n${n}.json 2> gpurun_out/part5_w${W}_n${n}.err
python - <<PY
import json
r = json.loads([l for l in open("gpurun_out/part5_w${W}_n${n}.json") if l.startswith("{")][-1])
print({k: r[k] for k in r if k != "ranks"})
PY
tail -3 gpurun_out/part5_w${W}_n${n}.err
python scripts/trace_summary.py "gpurun_out/trace_w${W}_r[01].npy" | tee gpurun_out/part5_w${W}_n${n}_trace.txt
done
