#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_2d.py -m gpu -q 2>&1 | tail -60 | tee gpurun_out/gpu2d_pytest.txt
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_2d.py 2>&1 | tail -8 | tee gpurun_out/gpu2d_pytest_rest.txt
