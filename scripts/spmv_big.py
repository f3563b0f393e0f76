"""SpMV roofline on a working set far larger than L2 (4 M vertices, ~1.2 GB of matrix)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, meshes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 158
xyz, tets = meshes.box_mesh((-10,) * 3, (10,) * 3, n, n, n)
fem = btfem.BTFem(0)
fem.set_mesh(xyz, tets)
fem.set_diffusion(3e-3)
t0 = time.perf_counter(); fem.assemble(); print("assemble %.2f s" % (time.perf_counter() - t0))
b = 20.0 * fem.nnz + 36.0 * fem.ndof
print("ndof", fem.ndof, "nnz", fem.nnz, "alg MB %.1f" % (b / 1e6))
for lanes in (0, 204, 8):
    for fl in (True, False):
        ms = fem.spmv_bench(200.0, 0.5, 1.5e-5, [0, 1, 0], lanes=lanes, nrep=10, flush_l2=fl)
        print("variant %3d flush %-5s ms %.4f GB/s %.0f frac_of_6534 %.3f" % (lanes, fl, ms, b / ms / 1e6, b / ms / 1e6 / 6534.1))
