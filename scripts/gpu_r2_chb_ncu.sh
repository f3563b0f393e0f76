#!/bin/bash
# ncu --set full of the two cooperative batch kernels on one 16-member HARDI batch (second launch: the first is the warm-up)
mkdir -p gpurun_out
BTFEM_BATCH_PERSIST=hb timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bicgstab_coop_hb -s 1 -c 1 -f -o gpurun_out/r2ao_prof_coop_hb python scripts/hardi_bench.py 4 16 > gpurun_out/r2ao_ncu_hb.log 2>&1
tail -2 gpurun_out/r2ao_ncu_hb.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bicgstab_coop_batch -s 1 -c 1 -f -o gpurun_out/r2ao_prof_coop_batch python scripts/hardi_bench.py 4 16 > gpurun_out/r2ao_ncu_cb.log 2>&1
tail -2 gpurun_out/r2ao_ncu_cb.log
ls -la gpurun_out/r2ao*
