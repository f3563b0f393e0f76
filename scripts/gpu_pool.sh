#!/bin/bash
mkdir -p gpurun_out
timeout 45 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pool_pytest.txt
timeout 25 python bench.py --steps 1 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g s/solve %.3f assemble_s %.3f' % (d['value'], d['e2e']['value'], d['e2e']['seconds_per_solve'], d['assemble_s']))" | tee gpurun_out/pool_bench.txt
