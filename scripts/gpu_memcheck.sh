#!/bin/bash
# compute-sanitizer memcheck over the triangle-mesh tests and the shared-operator batch kernel
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_2d.py -q -m gpu -k "not config3" > gpurun_out/memcheck_2d.txt 2>&1
BTFEM_BATCH_SHARED=1 timeout 120 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "batched" > gpurun_out/memcheck_batch_shared.txt 2>&1
grep -h "=========" gpurun_out/memcheck_2d.txt gpurun_out/memcheck_batch_shared.txt | grep -v "Host Frame\|^========= *$" | sort | uniq -c | sort -rn | head -20
