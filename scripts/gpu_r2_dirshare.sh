#!/bin/bash
# chain batch with one operator copy per direction (default) against one per member
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "persistent_batch or batched_solves or interleaved_batch" > gpurun_out/r2aj_dirshare_pytest.txt 2>&1
tail -5 gpurun_out/r2aj_dirshare_pytest.txt
grep -q "passed" gpurun_out/r2aj_dirshare_pytest.txt || { echo "tests did not pass: skipping the sweep"; exit 1; }
{
echo "== kernel chain, one operator copy per direction (default), batch 16"
timeout 150 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -3
echo "== kernel chain, one operator copy per member (BTFEM_BATCH_DIRSHARE=0), batch 16"
BTFEM_BATCH_DIRSHARE=0 timeout 150 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -3
echo "== per direction, batch 32"
timeout 150 python scripts/hardi_bench.py 64 32 2>&1 | grep -E "HARDI|rror" | tail -3
echo "== per direction, batch 64"
timeout 150 python scripts/hardi_bench.py 64 64 2>&1 | grep -E "HARDI|rror" | tail -3
} > gpurun_out/r2aj_dirshare_hardi.txt 2>&1
cat gpurun_out/r2aj_dirshare_hardi.txt
