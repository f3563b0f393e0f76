"""Quick A/B of the fused-SpMV variants on the bench workload (1 GPU): SpMV time cold / back to back and the loop
time of one solve.  Environment switches (BTFEM_NO_STREAM_KERNEL, BTFEM_PS_BPS, ...) select the variant."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.entry.load_package()
from dmri_fem_cloud_b200 import btfem  # noqa: E402

n_box = int(sys.argv[1]) if len(sys.argv) > 1 else 78
nsolve = int(sys.argv[2]) if len(sys.argv) > 2 else 1
xyz, tets, phase = bench.workload(n_box)
mp, ts, f, fp = bench.sequence(k=200.0)
q = mp.qvalue
g = np.array([0.0, 1.0, 0.0])
with btfem.BTFem(0) as fem:
    fem.set_mesh(xyz, tets, phase)
    fem.set_diffusion(3e-3)
    fem.set_relaxation(1e-16)
    fem.set_permeability(1e-5)
    fem.assemble()
    alg = 20.0 * fem.nnz + 36.0 * fem.ndof
    cold = fem.spmv_bench(200.0, 0.5, q, g, lanes=0, nrep=20, flush_l2=True)
    warm = fem.spmv_bench(200.0, 0.5, q, g, lanes=0, nrep=50, flush_l2=False)
    line = "kernel %d  spmv cold %.2f us (%.0f GB/s, %.3f of 6548.5)  warm %.2f us (%.0f GB/s)" % (
        fem.spmv_kernel, 1e3 * cold, alg / cold / 1e6, alg / cold / 1e6 / 6548.5, 1e3 * warm, alg / warm / 1e6)
    for _ in range(nsolve):
        res = fem.solve(200.0, 0.5, q * f, q * fp, g, rtol=1e-9, atol=1e-10, maxit=100000)
    t0 = time.perf_counter()
    res = fem.solve(200.0, 0.5, q * f, q * fp, g, rtol=1e-9, atol=1e-10, maxit=100000)
    wall = time.perf_counter() - t0
    print(line + "  | solve loop %.1f ms setup %.1f ms wall %.1f ms, %d iters, %.2f us/iter, signal %.13e" % (
        res["loop_ms"], res["setup_ms"], 1e3 * wall, res["total_iters"], 1e3 * res["loop_ms"] / res["total_iters"],
        res["signal"] / res["voi"]))
