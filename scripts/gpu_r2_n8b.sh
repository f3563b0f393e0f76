#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2ay_bench_n8.json 2> gpurun_out/r2ay_bench_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ay_bench_n8.json").read().strip().splitlines()[-1])
print("N=8 value %.4g e2e %.4g hardi %s" % (d["value"], d["e2e"]["value"], d.get("hardi", {}).get("value")))
p=d.get("partitioned") or {}
for k,v in p.items(): print(k, {x: v.get(x) for x in ("us_per_iteration","speedup_vs_single_gpu","rel_signal_err_vs_single_gpu")})
PY
tail -3 gpurun_out/r2ay_bench_n8.err
