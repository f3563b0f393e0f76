#!/bin/bash
mkdir -p gpurun_out
BTFEM_TIMING=1 timeout 300 python scripts/e2e_profile.py 2>&1 | tail -60 | tee gpurun_out/r2s_e2e_profile.txt
