#!/bin/bash
mkdir -p gpurun_out
for fl in "--no-full" "--no-full --no-hardi"; do
BENCH_DEBUG=1 timeout 600 python bench.py $fl > gpurun_out/r2aw_b.json 2> gpurun_out/r2aw_b.err
echo "flags: $fl"; grep "\[bench\]" gpurun_out/r2aw_b.err
done
BENCH_DEBUG=1 BTFEM_TIMING=1 timeout 600 python bench.py --no-full --no-hardi > gpurun_out/r2aw_b.json 2> gpurun_out/r2aw_b.err
echo "flags: TIMING --no-full --no-hardi"; grep "\[bench\]" gpurun_out/r2aw_b.err
