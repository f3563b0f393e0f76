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




def fixture_meshes():
    """Oracle signals on the reference's own small mesh fixtures (comri/meshes/*.zip), BASELINE configs[0]/[1]
    parameters.  The meshes themselves (tiny) are stored with the signals so that the GPU box, which has no
    /root/reference, can run the same cases."""
    ref = "/root/reference/comri/meshes"
    out = {}
    seq = orc.pgse(10600.0, 43100.0)
    q = seq.q_from_b(1000.0)
    for name in ("cyl6_r_3E_6_vol", "cyl12_r_3E_6_vol"):
        xyz, tets, marker = meshes.read_gmsh2(os.path.join(ref, name + ".msh.zip"))
        ops = orc.assemble(xyz, tets, D=3e-3, invT2=1e-16)
        r = orc.theta_solve(ops, seq, q, [1, 0, 0], 200.0, solver="lu")        # -M 0 ... -gdir 1 0 0
        out[name + "_xyz"], out[name + "_tets"] = xyz, tets
        out[name + "_signal"], out[name + "_voi"] = r["signal"], r["voi"]
        print(name, len(xyz), len(tets), r["signal"] / r["voi"])
    # two-compartment torus (phase from the compartment sub-mesh), -M 1 -b 1000 -p 1e-5 -k 200 -gdir 0 1 0
    xyz, tets = meshes.read_dolfin_xml(os.path.join(ref, "multi_layer_torus.xml.zip"))
    sx, st = meshes.read_dolfin_xml(os.path.join(ref, "multi_layer_torus_compt1.xml.zip"))
    phase = meshes.phase_from_submesh(xyz, tets, sx, st)
    ops = orc.assemble(xyz, tets, phase, D=3e-3, invT2=1e-16, kappa=1e-5)
    r = orc.theta_solve(ops, seq, q, [0, 1, 0], 200.0, solver="lu")
    out["torus_xyz"], out["torus_tets"], out["torus_phase"] = xyz.astype(np.float32).astype(np.float64), tets, phase
    # (coordinates are stored exactly; float32 round trip only to check they are representable -- they are not
    # in general, so store float64)
    out["torus_xyz"] = xyz
    out["torus_signal"], out["torus_voi"], out["torus_ndof"] = r["signal"], r["voi"], ops.ndof
    print("torus", len(xyz), len(tets), int(phase.sum()), ops.ndof, r["signal"] / r["voi"])
    np.savez_compressed(os.path.join(HERE, "fixture_meshes.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["box", "ufc", "fixtures"]
    if "box" in which:
        convergence_box()
    if "ufc" in which:
        ufc_tensors()
    if "fixtures" in which:
        fixture_meshes()
