#!/bin/bash
# persistent batch kernel: where block 0 spends a 16-member batch on the HARDI mesh
mkdir -p gpurun_out
{
BTFEM_PROFILE_PERSIST=1 timeout 150 python scripts/hardi_bench.py 4 16 2>&1 | grep -E "HARDI|rror|persistent kernel" | tail -4
} > gpurun_out/r2ah_pbatch_phases.txt 2>&1
cat gpurun_out/r2ah_pbatch_phases.txt
