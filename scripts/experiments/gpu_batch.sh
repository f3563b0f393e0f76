#!/bin/bash
# 1 GPU: parity tests, then the HARDI sweep (configs[4]) on the shared-operator batch kernel vs the per-member layout.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/batch_pytest.txt
for cfg in "1 8 16" "1 8 32" "1 4 16" "1 8 64" "0 8 16"; do
  set -- $cfg
  BTFEM_BATCH_SHARED=$1 BTFEM_BATCH_GROUP=$2 timeout 300 python scripts/hardi_bench.py 64 $3 2>&1 | grep "HARDI\|Error\|error" | sed "s/^/shared=$1 group=$2 /" | tee -a gpurun_out/batch_hardi.txt
done
