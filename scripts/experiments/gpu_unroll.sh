#!/bin/bash
# one GPU: device-driven loop, iterations per WHILE pass; 1 M-DOF solve and the small-mesh HARDI sweep
set -x
mkdir -p gpurun_out
for u in host 1 4 6 8 12; do
  if [ $u = host ]; then export BTFEM_LOOP=host; else export BTFEM_LOOP=device BTFEM_UNROLL=$u; fi
  python - <<'PY' 2>&1 | tee -a gpurun_out/unroll_sweep.txt
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import bench, __graft_entry__ as e
e.load_package()
from dmri_fem_cloud_b200 import btfem, meshes
tag = os.environ.get("BTFEM_LOOP") + " unroll=" + os.environ.get("BTFEM_UNROLL", "-")
mp, ts, f, fp = bench.sequence()
q = mp.qvalue
for name, (xyz, tets, phase) in (("cell-in-box 1M DOF", bench.workload(78)), ("cell-in-box 130k DOF", bench.workload(39)),
                                 ("neuron-like 46k vertices", (*meshes.neuron_like(h=0.7), None))):
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        fem.set_diffusion(3e-3); fem.set_relaxation(1e-16)
        if phase is not None: fem.set_permeability(1e-5)
        fem.assemble()
        fem.solve(200.0, 0.5, q * f, q * fp, [0, 1, 0], rtol=1e-9, atol=1e-10)
        t0 = time.perf_counter()
        r = fem.solve(200.0, 0.5, q * f, q * fp, [0, 1, 0], rtol=1e-9, atol=1e-10)
        wall = time.perf_counter() - t0
        print("%-14s %-26s ndof %7d iters %6d loop %8.2f ms wall %8.2f ms  %6.2f us/iter  signal %.12f" % (
            tag, name, fem.ndof, r["total_iters"], r["loop_ms"], 1e3 * wall, 1e3 * r["loop_ms"] / r["total_iters"], r["signal"] / r["voi"]))
PY
done
