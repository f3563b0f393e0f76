#!/bin/bash
mkdir -p gpurun_out
BTFEM_TIMING=1 timeout 300 python scripts/e2e_profile.py 2>&1 | grep -E "set_mesh|e2e " | tail -16 | tee gpurun_out/r2y_set_mesh.txt
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "theta_loop" > gpurun_out/r2y_racecheck.txt 2>&1
echo "exit code $?" >> gpurun_out/r2y_racecheck.txt
grep -h "=========" gpurun_out/r2y_racecheck.txt | grep -v "Host Frame\|^========= *$" | cut -c1-160 | sort | uniq -c | sort -rn | head -12
tail -4 gpurun_out/r2y_racecheck.txt
