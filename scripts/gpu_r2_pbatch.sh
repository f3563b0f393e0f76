#!/bin/bash
# persistent batch kernel: parity tests, then the HARDI sweep on one GPU with and without it
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "persistent_batch or batched_solves or interleaved_batch" > gpurun_out/r2ag_pbatch_pytest.txt 2>&1
tail -15 gpurun_out/r2ag_pbatch_pytest.txt
grep -q "passed" gpurun_out/r2ag_pbatch_pytest.txt || { echo "tests did not pass: skipping the sweep"; exit 1; }
{
echo "== persistent batch kernel, batch 16"
timeout 150 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -3
echo "== kernel chain (BTFEM_BATCH_PERSIST=0), batch 16"
BTFEM_BATCH_PERSIST=0 timeout 150 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -3
echo "== persistent batch kernel, batch 8"
timeout 150 python scripts/hardi_bench.py 64 8 2>&1 | grep -E "HARDI|rror" | tail -3
} > gpurun_out/r2ag_pbatch_hardi.txt 2>&1
cat gpurun_out/r2ag_pbatch_hardi.txt
