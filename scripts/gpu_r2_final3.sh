#!/bin/bash
# last check of the round after moving the batch kernels into their own headers: GPU suite, smoke, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/r2t_pytest.txt
timeout 200 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
timeout 400 python bench.py --no-full --steps 2 > gpurun_out/r2ba_bench_short.json 2> gpurun_out/r2ba_bench_short.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2ba_bench_short.json").read().strip().splitlines()[-1])
print("value %.4g ms/solve %.1f e2e %.4g (%.4f s/solve) hardi %s launches %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["seconds_per_solve"], d.get("hardi", {}).get("value"), d["gpu_launches"]))
PY
