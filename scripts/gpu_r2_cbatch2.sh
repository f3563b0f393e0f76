#!/bin/bash
mkdir -p gpurun_out
{
BTFEM_PROFILE_PERSIST=1 BTFEM_PROFILE_PERSIST_FILE=gpurun_out/r2am_cbatch_blocks.txt timeout 150 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel|coop batch" | tail -3
} > gpurun_out/r2am_cbatch_phases.txt 2>&1
cat gpurun_out/r2am_cbatch_phases.txt
