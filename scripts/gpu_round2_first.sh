#!/bin/bash
# First GPU call of the next round (1 GPU): what was written after this round's GPU budget ran out.
#   1. strong pseudo-periodic BC on hardware (csrc/strong.cu): the opt-in parity tests, then memcheck on them
#   2. the whole suite and the bench line with the new `hardi` key
# Results -> gpurun_out/r2_first_*.
set -x
mkdir -p gpurun_out
BTFEM_TEST_STRONG=1 BTFEM_STRONG=1 timeout 300 python -m pytest tests/test_gpu_strong.py -m gpu -q 2>&1 | tail -40 | tee gpurun_out/r2_first_strong.txt
BTFEM_TEST_STRONG=1 BTFEM_STRONG=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_strong.py -m gpu -q -k box_1c > gpurun_out/r2_first_strong_memcheck.txt 2>&1
grep -h "=========" gpurun_out/r2_first_strong_memcheck.txt | grep -v "Host Frame\|^========= *$" | sort | uniq -c | sort -rn | head
BTFEM_TEST_PENDING=1 timeout 300 python -m pytest tests/test_gpu_pending.py -m gpu -q 2>&1 | tail -20 | tee gpurun_out/r2_first_pending.txt
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_first_pytest.txt
timeout 600 python bench.py --cpu-sample-steps 8 > gpurun_out/r2_first_bench.json 2> gpurun_out/r2_first_bench.err
tail -c 1200 gpurun_out/r2_first_bench.json; tail -3 gpurun_out/r2_first_bench.err
