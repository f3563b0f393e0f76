"""ILU(0) preconditioner (csrc/ilu.cu) and device-resident GMRES against the oracle's restatements of PETSc's PCILU
defaults / KSPGMRES / KSPBCGS (oracle/bt_oracle.py: ilu0_factor, gmres_petsc, bicgstab_petsc):
KrylovSolver("gmres", "ilu") is the comri C++ demo's solver (comri/one-comp/fenics-cpp/main.cpp:180-183), and
KrylovSolver("bicgstab") -- PETSc's default preconditioner, ILU(0) on one process -- the notebooks'.
Bars: factor values 1e-12, iteration counts as the restatement's (+-2: different rounding of the dot products),
signals 1e-8."""
import numpy as np
import pytest

import bt_oracle as orc
from dmri_fem_cloud_b200 import btfem, meshes

pytestmark = pytest.mark.gpu


def _problem(two_comp):
    if two_comp:
        xyz, tets, ph = meshes.box_with_sphere(4.0, 6, 2.5)
        kw = dict(D=np.where(ph == 1, 1e-3, 2e-3), kappa=5e-5)
    else:
        xyz, tets = meshes.box_mesh((-2.5,) * 3, (2.5,) * 3, 8, 8, 8)
        ph, kw = None, dict(D=2e-3)
    seq = orc.pgse(1000.0, 3000.0)
    q, k = seq.q_from_b(1500.0), 100.0
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    g = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    return xyz, tets, ph, kw, seq, q, k, f, fp, g


@pytest.mark.parametrize("two_comp", [False, True], ids=["box_1c", "cell_in_box_2c"])
@pytest.mark.parametrize("ksp", ["gmres", "bicgstab"])
def test_ilu0_preconditioned_solves_match_the_petsc_restatement(ksp, two_comp):
    xyz, tets, ph, kw, seq, q, k, f, fp, g = _problem(two_comp)
    ops = orc.assemble(xyz, tets, ph, **kw)
    ref = orc.theta_solve(ops, seq, q, g, k, solver=ksp + "_ilu", rtol=1e-10, atol=1e-14, restart=30)
    jac = orc.theta_solve(ops, seq, q, g, k, solver=ksp, rtol=1e-10, atol=1e-14, restart=30)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, ph)
        fem.set_diffusion(kw["D"])
        if ph is not None:
            fem.set_permeability(kw["kappa"])
        fem.assemble()
        res = fem.solve(k, 0.5, q * f, q * fp, g, ksp=ksp, pc="ilu", rtol=1e-10, atol=1e-14, restart=30, want_iters=True)
        lu = fem.ilu_factors()
        again = fem.solve(k, 0.5, q * f, q * fp, g, ksp=ksp, pc="ilu", rtol=1e-10, atol=1e-14, restart=30)
        back = fem.solve(k, 0.5, q * f, q * fp, g, ksp=ksp, pc="jacobi", rtol=1e-10, atol=1e-14, restart=30, want_iters=True)
    # the factors of the LAST time step's operator (f(T) = 0 for PGSE: A = P), against the restatement
    P = (ops.M / k + 0.5 * (ops.S + ops.R + ops.I + ops.B)).tocsr()
    A = (P + 0.5j * q * f[-1] * (g[0] * ops.Jx + g[1] * ops.Jy + g[2] * ops.Jz)).tocsr()
    A.sort_indices()
    want = orc.ilu0_factor(A).lu
    assert np.max(np.abs(lu - want)) <= 1e-12 * np.max(np.abs(want))
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert np.max(np.abs(res["iters"].astype(int) - ref["iters"].astype(int))) <= 2
    assert again["signal"] == res["signal"]                                   # reproducible
    # ILU(0) needs fewer iterations than Jacobi, and the handle goes back to the Jacobi path afterwards
    assert res["iters"].sum() < 0.7 * back["iters"].sum()
    assert np.max(np.abs(back["iters"].astype(int) - jac["iters"].astype(int))) <= 2
    assert abs(back["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])


def test_gmres_device_resident_restart_cycles():
    """GMRES(5) with Jacobi: several restart cycles per time step, convergence inside a cycle (the device-side Givens /
    stop flag), against the restatement -- and the notebooks' GMRES rtol 1e-4 (RealNeurons.ipynb cell 10)."""
    xyz, tets, ph, kw, seq, q, k, f, fp, g = _problem(False)
    ops = orc.assemble(xyz, tets, ph, **kw)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets)
        fem.set_diffusion(kw["D"])
        fem.assemble()
        for restart, rtol in ((5, 1e-10), (30, 1e-4), (7, 1e-8)):
            ref = orc.theta_solve(ops, seq, q, g, k, solver="gmres", rtol=rtol, atol=1e-14, restart=restart)
            res = fem.solve(k, 0.5, q * f, q * fp, g, ksp="gmres", pc="jacobi", rtol=rtol, atol=1e-14, restart=restart,
                            want_iters=True)
            assert abs(res["signal"] - ref["signal"]) <= max(1e-8, 10 * rtol) * abs(ref["signal"])
            assert np.max(np.abs(res["iters"].astype(int) - ref["iters"].astype(int))) <= 2, (restart, rtol)
