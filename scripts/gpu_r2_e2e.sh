#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/e2e_profile.py > gpurun_out/r2au_e2e_profile.txt 2>&1
tail -30 gpurun_out/r2au_e2e_profile.txt
