#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
BTFEM_TIMING=1 timeout 600 python bench.py --steps 2 --warmup 3 --no-hardi --no-full > gpurun_out/r2av_bench_$i.json 2> gpurun_out/r2av_bench_$i.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2av_bench_$i.json").read().strip().splitlines()[-1])
print("run $i: value %.4g  ms/solve %.1f  e2e s/solve %.4f" % (d["value"], d["ms_per_step"], d["e2e"]["seconds_per_solve"]))
PY
grep -E "assemble:|set_mesh|pattern:" gpurun_out/r2av_bench_$i.err | tail -24
done
