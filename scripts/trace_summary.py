"""Summarise the device timelines written by scripts/dist_solve.py --trace (gpurun_out/trace_w{W}_r{R}.npy)."""
import glob, sys
import numpy as np
names = {0: "u push / signal", 1: "SpMV rhs", 2: "SpMV resid", 3: "SpMV v=Ap", 4: "SpMV t=As", 5: "update p (+push)",
         6: "update s (+push)", 7: "update x,r"}
for f in sorted(glob.glob(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_w*_r*.npy")):
    t = np.load(f).astype(np.int64)
    t = t[len(t) // 4:]                      # steady state
    if len(t) < 20:
        continue
    start, local, done, wait, kind = t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4]
    gap = start[1:] - done[:-1]              # end of previous kernel's collective -> first block of the next
    print(f, "entries", len(t), "span %.3f ms" % ((done[-1] - start[0]) * 1e-6))
    print("  %-18s %6s %9s %9s %9s %9s" % ("kernel", "count", "work us", "coll us", "wait us", "gap->next"))
    for k in sorted(set(kind.tolist())):
        m = kind == k
        mg = m[:-1]
        print("  %-18s %6d %9.2f %9.2f %9.2f %9.2f" % (names.get(k, str(k)), m.sum(), np.median(local[m] - start[m]) * 1e-3,
              np.median(done[m] - local[m]) * 1e-3, np.median(wait[m]) * 1e-3, np.median(gap[mg]) * 1e-3 if mg.any() else 0))
    it = kind == 7
    if it.sum() > 2:
        d = np.diff(done[it])
        print("  iteration period (update x,r to update x,r): median %.2f us, mean %.2f us" % (np.median(d) * 1e-3, d.mean() * 1e-3))
