"""Config 4 on one GPU: extracellular space of 226 cylinders, two compartments, weak pseudo-periodic BC in x and y
(ECS_226Cylinders.ipynb bbox): D=2e-3, kappa=1e-5, delta=10000, Delta=13000, k=200, g=(1,1,0)/sqrt2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sympy as sp
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import dmrifemlib as dl, meshes

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 400
xyz, tets, phase = meshes.ecs_slab(nx, nx, 2)
print("mesh", len(xyz), len(tets), "phase1 cells", int(phase.sum()))
for b in (1000.0, 10000.0):
    mesh = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.bvalue = b
    mp.delta, mp.Delta = 10000.0, 13000.0
    mp.T = mp.delta + mp.Delta
    mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
    mp.set_gradient_dir(mesh, 1, 1, 0)
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 200.0
    sim.verbose = False
    t0 = time.perf_counter()
    md = dl.MyDomain(mesh, mp)
    md.phase, md.IsDomainMultiple, md.kappa = phase, True, 1e-5
    md.PeriodicDir = [1, 1, 0]
    md.Apply()
    md.D0 = 2e-3
    md.D = md.D0
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    ls.parameters.update({"relative_tolerance": 1e-9, "absolute_tolerance": 1e-10, "maximum_iterations": 100000})
    sim.solve(md, mp, ls)
    dt = time.perf_counter() - t0
    st = sim.stats
    nd = 2 * sim.fem.ndof
    print("b=%g ndof_real=%d steps=%d iters=%d loop_ms=%.1f e2e_s=%.2f DOF-steps/s(loop)=%.3g e2e=%.3g signal=%.6e hmin=%.3f kappa_e=%.3e"
          % (b, nd, st["n_steps"], st["total_iters"], st["loop_ms"], dt, nd * st["n_steps"] / (st["loop_ms"] * 1e-3),
             nd * st["n_steps"] / dt, st["signal"] / st["voi"], md.hmin, md.kappa_e_scalar))
    sim.fem.close()
