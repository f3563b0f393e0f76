#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-hardi --no-cpu > gpurun_out/r2ae_bench_quick.json 2> gpurun_out/r2ae_bench_quick.err
tail -3 gpurun_out/r2ae_bench_quick.err
python - <<'PY'
import json
b=json.loads([l for l in open('gpurun_out/r2ae_bench_quick.json') if l.startswith('{')][-1])
print('value',b['value'],'ms',b['ms_per_step'],'e2e',b['e2e']['value'],b['e2e']['seconds_per_solve'])
PY
timeout 300 python -m pytest tests/test_gpu_driver.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
