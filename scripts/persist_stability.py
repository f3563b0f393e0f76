"""Are the per-block pass times of the persistent kernel a stable property of the device (block -> SM placement and
the SM's share of the memory system)?  Two solves in one process, per-warp times of both written to gpurun_out/."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.entry.load_package()
from dmri_fem_cloud_b200 import btfem  # noqa: E402

os.environ["BTFEM_PROFILE_PERSIST"] = "1"
xyz, tets, phase = bench.workload(78)
mp, ts, f, fp = bench.sequence(k=200.0)
q = mp.qvalue
out = []
for rep in range(2):
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        fem.set_diffusion(3e-3)
        fem.set_relaxation(1e-16)
        fem.set_permeability(1e-5)
        fem.assemble()
        for k in range(2):
            path = os.path.join(ROOT, "gpurun_out", "stab_%d_%d.txt" % (rep, k))
            os.environ["BTFEM_PROFILE_PERSIST_FILE"] = path
            fem.solve(200.0, 0.5, q * f, q * fp, [0, 1, 0] if k == 0 else [1, 0, 0], rtol=1e-9, atol=1e-10, maxit=100000)
            d = np.loadtxt(path)
            nw = len(d) // 148
            out.append((d[:, 1].reshape(148, nw).mean(axis=1), d[::nw, 5]))
for i in range(len(out)):
    for j in range(i + 1, len(out)):
        print("runs %d,%d: corr of per-block time %.3f, same SM placement: %s" % (
            i, j, np.corrcoef(out[i][0], out[j][0])[0, 1], np.array_equal(out[i][1], out[j][1])))
t = out[0][0] / out[0][0].mean()
sm = out[0][1].astype(int)
print("rel time by SM id (sorted by SM):", np.round(t[np.argsort(sm)], 2).tolist())
