#!/bin/bash
for k in 1 4 16; do echo "== $k concurrent"; timeout 300 python scripts/hardi_bench.py 16 16 $k 2>&1 | grep -E "HARDI|per solve|rror" | tail -3; done
