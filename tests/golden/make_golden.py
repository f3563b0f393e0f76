#!/usr/bin/env python3
"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference
for the UFC element tensors; the whole-path fixture needs only the oracle).

  convergence_box_n16.npz  oracle (exact LU stepping) on the BoxMesh n=16 problem whose recorded
                           reference output is 8.440078e-01 (ConvergenceTest.ipynb cell 10)
  ufc_element_tensors.npz  outputs of the reference's own FFC-generated tabulate_tensor kernels
                           (oracle/_ref) on seeded random tets / facet pairs
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
import bt_oracle as orc  # noqa: E402
import ufc_ref  # noqa: E402
from dmri_fem_cloud_b200 import meshes  # noqa: E402


def convergence_box():
    xyz, tets = meshes.box_mesh((-2.5,) * 3, (2.5,) * 3, 16, 16, 16)
    ops = orc.assemble(xyz, tets, D=2e-3)
    seq = orc.pgse(1000.0, 10000.0)
    q = seq.q_from_b(1000.0)
    r = orc.theta_solve(ops, seq, q, [0, 0, 1], 10.0, solver="lu", closed=False)
    np.savez(os.path.join(HERE, "convergence_box_n16.npz"), normalized_signal_lu=r["signal"] / r["voi"],
             signal=r["signal"], voi=r["voi"], n_steps=r["n_steps"], q=q, recorded_reference="8.440078e-01",
             analytic=0.84389487095614)
    print("convergence box:", r["signal"] / r["voi"])


def ufc_tensors():
    out = ufc_ref.golden_element_tensors(seed=2024, ncell=24)
    np.savez(os.path.join(HERE, "ufc_element_tensors.npz"), **out)
    print("ufc tensors:", {k: np.shape(v) for k, v in out.items()})


if __name__ == "__main__":
    convergence_box()
    ufc_tensors()
