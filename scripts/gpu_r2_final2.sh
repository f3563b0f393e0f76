#!/bin/bash
# default bench line + reference arm on one GPU (after the e2e loop change)
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
tail -c 600 gpurun_out/r2t_bench.json; tail -3 gpurun_out/r2t_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench.json").read().strip().splitlines()[-1])
print("value %.4g ms/solve %.1f e2e %.4g (%.4f s/solve) hardi %s steps %d warmup %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_solve"], d.get("hardi", {}).get("value"), d["steps"], d["warmup"]))
PY
timeout 900 python bench.py --impl reference > gpurun_out/r2t_bench_reference.json 2> gpurun_out/r2t_bench_reference.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2t_bench_reference.json").read().strip().splitlines()[-1])
print("reference value %.4g cores %s" % (d["value"], d["cpu_baseline"]["cores"]))
PY
