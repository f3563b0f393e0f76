#!/bin/bash
# last sanity of the round on the committed library: batch tests + one HARDI sweep + smoke
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "persistent_batch or batch_on_an_sm or batched_solves or theta_loop" 2>&1 | tail -2
timeout 100 python scripts/hardi_bench.py 64 16 2>&1 | grep -E "HARDI|rror" | tail -2
timeout 100 python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
