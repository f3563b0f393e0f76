"""One lock-step batch (4 directions x 4 b-values) on the HARDI mesh: for per-kernel profiles of the batch paths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sympy as sp
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes, sweep

nsteps_T = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0     # > 0: shorten the sequence (profiling)
xyz, tets = meshes.neuron_like(h=0.7)
xyz, tets = meshes.rcm_order(*meshes.shuffle_vertices(xyz, tets, 0))
mp = dl.MRI_parameters()
mp.delta, mp.Delta = 10600.0, 43100.0
if nsteps_T > 0:
    mp.delta, mp.Delta = nsteps_T, 2 * nsteps_T
mp.T = mp.delta + mp.Delta
mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
mp.bvalue = 1000.0
mp.Apply()
sim = dl.MRI_simulation()
sim.k = 200.0
dirs = meshes.fibonacci_hemisphere(4)
bvals = [1000.0, 2000.0, 3000.0, 4000.0]
par = dict(rtol=1e-9, atol=1e-10, maxit=100000)
fem = btfem.BTFem(0)
fem.set_mesh(xyz, tets)
fem.set_diffusion(3e-3)
fem.set_relaxation(1e-16)
fem.assemble()
t0 = time.perf_counter()
mine, sig = sweep.run_sweep(fem, mp, sim, dirs, bvals, par, batch=16)
print("one batch of 16: %.3f s, signals %s" % (time.perf_counter() - t0, np.round(sig[:4], 6)))
