"""Small batch through the three cooperative batch kernels (for compute-sanitizer: short sequence, small mesh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, meshes

xyz, tets, ph = meshes.box_with_sphere(10.0, 6, 5.0)
k = 200.0
ts = np.arange(1, 7) * k
f = np.where(ts <= 600.0, 1.0, -1.0)
fp = np.concatenate([[f[0]], f[:-1]])
dirs = meshes.fibonacci_hemisphere(3)
members = [(q * f, q * fp, d) for d in dirs for q in (1e-4, 3e-4)] + [(2e-4 * f, 2e-4 * fp, dirs[0])]
which = sys.argv[1] if len(sys.argv) > 1 else ""
if which:
    os.environ["BTFEM_BATCH_PERSIST"] = which
with btfem.BTFem(0) as fem:
    fem.set_mesh(xyz, tets, ph)
    fem.set_diffusion(3e-3)
    fem.set_relaxation(1.0 / 4e4)
    fem.set_permeability(5e-5)
    fem.assemble()
    single = [fem.solve(k, 0.5, cA, cb, g, rtol=1e-10, atol=1e-14) for cA, cb, g in members[:2]]
    batch = fem.solve_batch(k, 0.5, members, rtol=1e-10, atol=1e-14)
err = max(abs(s["signal"] - b["signal"]) / abs(s["signal"]) for s, b in zip(single, batch))
print("kernel %-6s kernels %d  iterations %s  max rel diff to single solves %.2e" % (
    which or "coop", batch[0]["n_kernels"], [int(b["total_iters"]) for b in batch], err))
