"""Where does the end-to-end time go? (GPU box)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl

xyz, tets, phase = bench.workload(78)
mp, ts, f, fp = bench.sequence()
q = mp.qvalue
for rep in range(3):
    T = {}
    t0 = time.perf_counter(); fem = btfem.BTFem(0); T["create"] = time.perf_counter() - t0
    t0 = time.perf_counter(); fem.set_mesh(xyz, tets, None); T["set_mesh"] = time.perf_counter() - t0
    t0 = time.perf_counter(); fem.mesh_stats(); T["mesh_stats"] = time.perf_counter() - t0
    t0 = time.perf_counter(); fem.set_phase(phase); T["set_phase"] = time.perf_counter() - t0
    t0 = time.perf_counter(); fem.set_diffusion(3e-3); fem.set_relaxation(1e-16); fem.set_permeability(1e-5); fem.set_initial(np.ones(len(xyz))); T["setters"] = time.perf_counter() - t0
    t0 = time.perf_counter(); fem.assemble(); T["assemble"] = time.perf_counter() - t0
    t0 = time.perf_counter(); r = fem.solve(200.0, 0.5, q * f, q * fp, [0, 1, 0], rtol=1e-9, atol=1e-10); T["solve"] = time.perf_counter() - t0
    T["solve.loop_ms"] = r["loop_ms"] / 1e3; T["solve.setup_ms"] = r["setup_ms"] / 1e3
    t0 = time.perf_counter(); fem.close(); T["close"] = time.perf_counter() - t0
    t0 = time.perf_counter(); sim = dl.MRI_simulation(); sim.k = 200.0; tg = sim.time_grid(mp); mp.profiles_on_grid(tg); mp.profiles_on_grid(tg); T["profiles"] = time.perf_counter() - t0
    t0 = time.perf_counter(); xyz.min(axis=0); xyz.max(axis=0); ((xyz * xyz).sum(axis=1) < 1e6); T["numpy_bbox_ic"] = time.perf_counter() - t0
    print(" ".join("%s=%.1fms" % (k, 1e3 * v) for k, v in T.items()))
