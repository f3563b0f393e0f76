"""Fused SpMV time against problem size on one GPU (warm, back to back): where the kernel leaves the bandwidth
regime -- the per-GPU sizes of a row-partitioned 1 M-DOF solve are 250 k (2 GPUs) ... 62 k rows (8 GPUs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem
for nbox in (24, 31, 39, 49, 62, 78, 98):
    xyz, tets, phase = bench.workload(nbox)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase); fem.set_diffusion(3e-3); fem.set_relaxation(1e-16); fem.set_permeability(1e-5)
        fem.assemble()
        ms = fem.spmv_bench(200.0, 0.5, 1e-5, [0, 1, 0], lanes=0, nrep=200, flush_l2=False)
        b = 20.0 * fem.nnz + 36.0 * fem.ndof
        print("n_box %3d rows %8d nnz %9d  %7.2f us  %7.1f GB/s algorithmic  %6.2f ns/kilorow" % (
            nbox, fem.ndof, fem.nnz, 1e3 * ms, b / ms / 1e6, 1e6 * ms / (fem.ndof / 1e3)))
