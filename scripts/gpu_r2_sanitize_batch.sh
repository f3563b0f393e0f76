#!/bin/bash
# compute-sanitizer over the cooperative batch kernels (memcheck: out-of-bounds; racecheck: shared-memory hazards)
mkdir -p gpurun_out
out=gpurun_out/r2az_sanitizer_batch_kernels.txt
: > $out
for w in "" ring hb; do
  for tool in memcheck racecheck; do
    echo "=== $tool, BTFEM_BATCH_PERSIST='$w'" >> $out
    timeout 170 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/sanitize_batch.py $w > gpurun_out/san.tmp 2>&1
    echo "exit code $?" >> $out
    grep -E "kernel |ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error" gpurun_out/san.tmp | sort | uniq -c | sort -rn | head -8 >> $out
  done
done
cat $out
