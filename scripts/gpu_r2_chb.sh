#!/bin/bash
# member-interleaved batch as one cooperative kernel (BTFEM_BATCH_PERSIST=hb): tests, phases, HARDI sweep
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "persistent_batch or batched_solves or interleaved_batch" > gpurun_out/r2an_chb_pytest.txt 2>&1
tail -5 gpurun_out/r2an_chb_pytest.txt
grep -q "passed" gpurun_out/r2an_chb_pytest.txt || { echo "tests did not pass: skipping the sweep"; tail -60 gpurun_out/r2an_chb_pytest.txt; exit 1; }
{
for cfg in 0 1 2 3; do
echo "== coop hb kernel cfg $cfg: phases of block 0, one batch of 16"
BTFEM_BATCH_PERSIST=hb BTFEM_CHB_CFG=$cfg BTFEM_PROFILE_PERSIST=1 timeout 150 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel|coop batch" | tail -3
done
echo "== coop hb kernel cfg 0, 64 directions, batch 16"
BTFEM_BATCH_PERSIST=hb timeout 150 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -3
echo "== coop member-layout kernel (default), batch 16"
timeout 150 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -3
} > gpurun_out/r2an_chb_hardi.txt 2>&1
cat gpurun_out/r2an_chb_hardi.txt
