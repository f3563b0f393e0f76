#!/bin/bash
set -x
W=${W:-2}
mkdir -p gpurun_out
# (size curve: scripts/spmv_sizes.py, profiles/r1c_spmv_sizes.txt)
timeout 600 python -m pytest tests/test_gpu_partition.py -x -q --timeout 400 2>&1 | tail -5
for n in 78; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 2973$W \
  scripts/dist_solve.py --nbox $n --reps 2 --trace 12000 > gpurun_out/part6_w${W}_n${n}.json 2> gpurun_out/part6_w${W}_n${n}.err
python - <<PY
import json
r = json.loads([l for l in open("gpurun_out/part6_w${W}_n${n}.json") if l.startswith("{")][-1])
print({k: r[k] for k in r if k != "ranks"})
PY
python scripts/trace_summary.py "gpurun_out/trace_w${W}_r[01].npy" | tee gpurun_out/part6_w${W}_n${n}_trace.txt
done
