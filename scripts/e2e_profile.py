"""Where does the end-to-end solve (bench.py's e2e: MyDomain / MRI_simulation.solve from host numpy) spend its time?
Wall clock per stage + cProfile of one call."""
import cProfile
import contextlib
import io
import os
import pstats
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

bench.entry.load_package()
from dmri_fem_cloud_b200 import dmrifemlib as dl  # noqa: E402

xyz, tets, phase = bench.workload(78)
mp, ts, f, fp = bench.sequence(k=200.0)
mp.set_gradient_dir(None, 0, 1, 0)
keep = []


def e2e_once():
    T = {}
    t0 = time.perf_counter(); mesh = dl.Mesh(xyz, tets); mesh.device = 0; T["Mesh"] = time.perf_counter() - t0
    t0 = time.perf_counter(); md = dl.MyDomain(mesh, mp); T["MyDomain"] = time.perf_counter() - t0
    md.phase = phase
    md.IsDomainMultiple = True
    md.kappa = 1e-5
    t0 = time.perf_counter(); md.Apply(); T["Apply"] = time.perf_counter() - t0
    md.D0 = 3e-3
    md.D = md.D0
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    ls.parameters.update({"relative_tolerance": 1e-9, "absolute_tolerance": 1e-10, "maximum_iterations": 100000})
    sim = dl.MRI_simulation()
    sim.k = 200.0
    sim.verbose = False
    t0 = time.perf_counter(); sim.solve(md, mp, ls); T["solve"] = time.perf_counter() - t0
    T["  loop_ms"] = sim.stats["loop_ms"] * 1e-3
    T["  setup_ms"] = sim.stats["setup_ms"] * 1e-3
    keep.append(sim.fem)
    return T


with contextlib.redirect_stdout(io.StringIO()):
    e2e_once()
    keep.pop().close()
    t0 = time.perf_counter()
    T = e2e_once()
    total = time.perf_counter() - t0
    pr = cProfile.Profile()
    pr.enable()
    e2e_once()
    pr.disable()
print("e2e %.1f ms: " % (1e3 * total) + " ".join("%s=%.1fms" % (k, 1e3 * v) for k, v in T.items()))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22)
print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:5000])
