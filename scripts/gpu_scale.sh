#!/bin/bash
# W GPUs of one box: row-partitioned solves (1 M DOF, 4 M DOF, periodic ECS) against one GPU, the HARDI sweep,
# and the weak-scaling bench line.  Results -> gpurun_out/scale_w$W_*.
set -x
W=${W:-2}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29801 scripts/dist_solve.py --nbox 78 --reps 2 --check --trace 12000 > gpurun_out/scale_w${W}_n78.json 2> gpurun_out/scale_w${W}.err
python scripts/trace_summary.py "gpurun_out/trace_w${W}_r[01].npy" > gpurun_out/scale_w${W}_n78_trace.txt
run 29802 scripts/dist_solve.py --nbox 124 --reps 2 --check > gpurun_out/scale_w${W}_n124.json 2>> gpurun_out/scale_w${W}.err
run 29803 scripts/dist_solve.py --ecs 400 --reps 2 --check > gpurun_out/scale_w${W}_ecs400.json 2>> gpurun_out/scale_w${W}.err
run 29804 scripts/dist_solve.py --ecs 800 --reps 2 --check > gpurun_out/scale_w${W}_ecs800.json 2>> gpurun_out/scale_w${W}.err
run 29805 scripts/hardi_bench.py 64 16 2>&1 | grep HARDI | tee gpurun_out/scale_w${W}_hardi.txt
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/scale_w${W}_*.json")):
    try:
        r = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, {k: (round(v, 5) if isinstance(v, float) else v) for k, v in r.items() if k not in ("ranks",)})
    except Exception as e:
        print(f, "no result", e)
PY
tail -5 gpurun_out/scale_w${W}.err
cat gpurun_out/scale_w${W}_n78_trace.txt
