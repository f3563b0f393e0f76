#!/bin/bash
# compute-sanitizer memcheck over the round-2 kernels (persistent BiCGStab, TMA SpMV, ILU(0), device GMRES, interleaved batch)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ilu.py -m gpu -q -x -k "theta_loop or fused_spmv or interleaved or ilu0 or gmres_device or pattern_and_values" > gpurun_out/r2x_memcheck.txt 2>&1
echo "exit code $?" >> gpurun_out/r2x_memcheck.txt
grep -h "=========" gpurun_out/r2x_memcheck.txt | grep -v "Host Frame\|^========= *$" | sort | uniq -c | sort -rn | head -20
tail -5 gpurun_out/r2x_memcheck.txt
