"""Config 5 (HARDI sweep on a neuron-like mesh): signals/s on this rank's share of directions x b-values."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sympy as sp
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes, sweep

ndir = int(sys.argv[1]) if len(sys.argv) > 1 else 8
xyz, tets = meshes.neuron_like(h=0.7)
xyz, tets = meshes.rcm_order(*meshes.shuffle_vertices(xyz, tets, 0))
mp = dl.MRI_parameters()
mp.delta, mp.Delta = 10600.0, 43100.0
mp.T = mp.delta + mp.Delta
mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
mp.bvalue = 1000.0
mp.Apply()
sim = dl.MRI_simulation()
sim.k = 200.0
dirs = meshes.fibonacci_hemisphere(ndir)
bvals = [1000.0, 2000.0, 3000.0, 4000.0]
fem = btfem.BTFem(0)
fem.set_mesh(xyz, tets)
fem.set_diffusion(3e-3)
fem.set_relaxation(1e-16)
fem.assemble()
print("mesh", len(xyz), len(tets), "nnz", fem.nnz)
par = dict(rtol=1e-9, atol=1e-10, maxit=100000)
sweep.run_sweep(fem, mp, sim, dirs[:1], bvals[:1], par)          # warm-up
t0 = time.perf_counter()
mine, sig = sweep.run_sweep(fem, mp, sim, dirs, bvals, par, batch=int(sys.argv[2]) if len(sys.argv) > 2 else 1)
dt = time.perf_counter() - t0
print("signals", len(sig), "seconds %.3f" % dt, "signals/s %.2f" % (len(sig) / dt), "ms/solve %.1f" % (1e3 * dt / len(sig)))
r = fem.solve(200.0, 0.5, np.zeros(270), np.zeros(270), [1, 0, 0], **par)
print("q=0 solve: iters", r["total_iters"], "loop_ms %.1f" % r["loop_ms"], "kernels", r["n_kernels"])
mp.bvalue = 4000.0; mp.gvalue = None; mp.Apply()
ts = sim.time_grid(mp); f, _ = mp.profiles_on_grid(ts); fp = np.concatenate([[f[0]], f[:-1]])
r = fem.solve(200.0, 0.5, mp.qvalue * f, mp.qvalue * fp, dirs[0], **par)
print("b=4000 solve: iters", r["total_iters"], "loop_ms %.1f" % r["loop_ms"], "us/iter %.1f" % (1e3 * r["loop_ms"] / r["total_iters"]), "signal", r["signal"] / r["voi"])
print(sig[:8])
