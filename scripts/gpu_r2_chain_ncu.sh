#!/bin/bash
# per-kernel durations of the kernel-chain batch (16 members, HARDI mesh), host-driven loop so that ncu sees the kernels
mkdir -p gpurun_out
BTFEM_LOOP=host BTFEM_BATCH_PERSIST=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/r2ai_chain_batch_launches.csv python scripts/hardi_one_batch.py 1000 > gpurun_out/r2ai_chain_batch.log 2>&1
tail -2 gpurun_out/r2ai_chain_batch.log
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r2ai_chain_batch_launches.csv")) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
d = collections.defaultdict(list)
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    d[r[ik][:60]].append(v * (1e-3 if u in ("ns", "nsecond") else 1.0 if u in ("us", "usecond") else 1e3))
for k, v in d.items():
    v.sort()
    print("%-62s n=%4d median %.1f us  min %.1f  max %.1f" % (k, len(v), v[len(v)//2], v[0], v[-1]))
PY
