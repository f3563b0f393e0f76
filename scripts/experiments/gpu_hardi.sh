#!/bin/bash
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for nb in 1 4 8 16 32; do timeout 300 python scripts/hardi_bench.py 8 $nb 2>&1 | grep -E "signals|Error|error" ; done
