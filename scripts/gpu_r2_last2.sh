#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-full --no-hardi --steps 2 > gpurun_out/r2bd_bench.json 2> gpurun_out/r2bd_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2bd_bench.json").read().strip().splitlines()[-1])
print("value %.4g e2e %.4f s/solve; e2e clocks %s; steps %s" % (d["value"], d["e2e"]["seconds_per_solve"], d["e2e"]["clocks"], d["e2e"]["rank0_steps"]))
PY
tail -2 gpurun_out/r2bd_bench.err
