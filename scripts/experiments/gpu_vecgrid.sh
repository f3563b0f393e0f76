#!/bin/bash
for vb in 8 4 2 1; do
  BTFEM_VEC_BLOCKS=$vb timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu 2>&1 | python -c "
import sys, json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('vec_blocks=$vb ms/solve %.1f value %.4g iters %d'%(d['ms_per_step'], d['value'], d['config']['iters_per_solve']))"
done
