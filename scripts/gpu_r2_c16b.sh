#!/bin/bash
BTFEM_TIMING=1 timeout 200 python scripts/spmv_quick.py 78 0 2>&1 | grep -E "warp streams|kernel 2"
