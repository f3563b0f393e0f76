#!/bin/bash
# A/B: L2 persistence window and 16-bit column offsets
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for cfg in "1 1" "0 1" "1 0" "0 0"; do set -- $cfg
  BTFEM_L2_PERSIST=$1 BTFEM_IDX16=$2 timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu 2>&1 | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('persist=$1 idx16=$2 value %.4g ms/solve %.1f e2e %.4g spmv cold %.1f us warm %.1f us frac %.3f'%(d['value'], d['ms_per_step'], d['e2e']['value'], 1e3*d['roofline']['ms_per_launch_l2_flushed'], 1e3*d['roofline']['ms_per_launch_back_to_back'], d['roofline']['frac']))"
done
