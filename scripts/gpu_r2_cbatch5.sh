#!/bin/bash
# coop batch kernel, slice-major items
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "persistent_batch or batched_solves or interleaved_batch" > gpurun_out/r2ar_cbatch_pytest.txt 2>&1
tail -3 gpurun_out/r2ar_cbatch_pytest.txt
grep -q "passed" gpurun_out/r2ar_cbatch_pytest.txt || { tail -60 gpurun_out/r2ar_cbatch_pytest.txt; exit 1; }
{
echo "== coop batch kernel: phases of block 0, one batch of 16"
BTFEM_PROFILE_PERSIST=1 timeout 150 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel|coop batch" | tail -3
for b in 16 8 24 32; do
echo "== coop batch kernel, batch $b"
timeout 150 python scripts/hardi_bench.py 64 $b 2>&1 | grep -E "HARDI|rror" | tail -3
done
} > gpurun_out/r2ar_cbatch_slice_major.txt 2>&1
cat gpurun_out/r2ar_cbatch_slice_major.txt
