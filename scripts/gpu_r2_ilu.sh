#!/bin/bash
set -x
timeout 600 python -m pytest tests/test_gpu_ilu.py -m gpu -q -x 2>&1 | tail -30
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8
