"""Is a block's pass time a property of the SM it runs on or of the slices it was dealt?  Same solve with the warp
lists rotated by 37 blocks (BTFEM_PS_ROTATE), per-block times compared both ways."""
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
for rot in (0, 37):
    os.environ["BTFEM_PS_ROTATE"] = str(rot)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        fem.set_diffusion(3e-3)
        fem.set_relaxation(1e-16)
        fem.set_permeability(1e-5)
        fem.assemble()
        path = os.path.join(ROOT, "gpurun_out", "rot_%d.txt" % rot)
        os.environ["BTFEM_PROFILE_PERSIST_FILE"] = path
        fem.solve(200.0, 0.5, q * f[:60], q * fp[:60], [0, 1, 0], rtol=1e-9, atol=1e-10, maxit=100000)
        d = np.loadtxt(path)
        nw = len(d) // 148
        out.append(d[:, 1].reshape(148, nw).mean(axis=1))
a, b = out
print("corr(block time, block time after rotation)            [SM-bound if ~1]: %.3f" % np.corrcoef(a, b)[0, 1])
print("corr(block time, time of the block that got its lists) [data-bound if ~1]: %.3f" % np.corrcoef(a, np.roll(b, 37))[0, 1],
      "/ other roll direction %.3f" % np.corrcoef(a, np.roll(b, -37))[0, 1])
