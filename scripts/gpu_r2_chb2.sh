#!/bin/bash
# coop hb kernel with batched operator loads: tuning variants
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "persistent_batch" > gpurun_out/r2as_chb_pytest.txt 2>&1
tail -3 gpurun_out/r2as_chb_pytest.txt
grep -q "passed" gpurun_out/r2as_chb_pytest.txt || { tail -60 gpurun_out/r2as_chb_pytest.txt; exit 1; }
{
for cfg in 0 1 2 3 4 5; do
echo "== coop hb kernel cfg $cfg: phases of block 0, one batch of 16"
BTFEM_BATCH_PERSIST=hb BTFEM_CHB_CFG=$cfg BTFEM_PROFILE_PERSIST=1 timeout 150 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel|coop batch" | tail -3
done
} > gpurun_out/r2as_chb_variants.txt 2>&1
cat gpurun_out/r2as_chb_variants.txt
