"""SpMV kernel timing for every lanes-per-row variant on the cell-in-box mesh (GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as e  # noqa: E402

e.load_package()
from dmri_fem_cloud_b200 import btfem, meshes  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 78
xyz, tets, ph = meshes.box_with_sphere(10.0, n, 5.0)
fem = btfem.BTFem(0)
fem.set_mesh(xyz, tets, ph)
fem.set_diffusion(3e-3)
fem.set_permeability(1e-5)
fem.assemble()
b = 20.0 * fem.nnz + 36.0 * fem.ndof
print("ndof", fem.ndof, "nnz", fem.nnz, "nnz/row %.2f" % (fem.nnz / fem.ndof), "alg MB %.1f" % (b / 1e6))
for lanes in (0, 102, 103, 104, 106, 202, 203, 204, 4, 8):
    for fl in (True, False):
        ms = fem.spmv_bench(200.0, 0.5, 1.5e-5, [0, 1, 0], lanes=lanes, nrep=30, flush_l2=fl)
        print("lanes %3d flush %-5s ms %.4f GB/s %.0f" % (lanes, fl, ms, b / ms / 1e6))
