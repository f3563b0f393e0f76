#!/bin/bash
# coop batch kernel: chunks cut by the L1-wavefront cost of the slices (host-counted) against the closed-form cost
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "persistent_batch or batch_on_an_sm" 2>&1 | tail -2
{
echo "== cost = L1 wavefronts (default): phases, one batch of 16"
BTFEM_PROFILE_PERSIST=1 timeout 100 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel|coop batch" | tail -3
echo "== cost = columns + 6 (BTFEM_CB_PLAIN_COST=1): phases, one batch of 16"
BTFEM_CB_PLAIN_COST=1 BTFEM_PROFILE_PERSIST=1 timeout 100 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel|coop batch" | tail -3
echo "== cost = L1 wavefronts, 64 directions"
timeout 100 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -2
echo "== cost = columns + 6, 64 directions"
BTFEM_CB_PLAIN_COST=1 timeout 100 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -2
} > gpurun_out/r2bc_cbatch_wavefront_cost.txt 2>&1
cat gpurun_out/r2bc_cbatch_wavefront_cost.txt
