#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "persistent_batch or batched_solves or interleaved_batch" > gpurun_out/r2aq_cbatch_pytest.txt 2>&1
tail -3 gpurun_out/r2aq_cbatch_pytest.txt
grep -q "passed" gpurun_out/r2aq_cbatch_pytest.txt || { tail -60 gpurun_out/r2aq_cbatch_pytest.txt; exit 1; }
{
for b in 16 24 32; do
echo "== coop batch kernel, batch $b"
timeout 150 python scripts/hardi_bench.py 64 $b 2>&1 | grep -E "HARDI|rror" | tail -3
done
} > gpurun_out/r2aq_cbatch_sizes.txt 2>&1
cat gpurun_out/r2aq_cbatch_sizes.txt
