#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/solve %.1f e2e %.4g (%.4f s/solve) hardi %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_solve"], d.get("hardi", {}).get("value")))
print(d["e2e"]["rank0_steps"])
PY
tail -3 gpurun_out/r2t_bench.err
