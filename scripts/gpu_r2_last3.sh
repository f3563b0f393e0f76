#!/bin/bash
mkdir -p gpurun_out
timeout 110 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/r2t_pytest.txt
