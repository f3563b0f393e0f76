"""GPU parity of the STRONGLY imposed pseudo-periodic BC (csrc/strong.cu, btfem_set_periodic_map) against the oracle's
restatement (pinned on the CPU by tests/test_oracle_strong.py).

First run on hardware: round 2 (profiles/r2a_strong.txt, 3 passed).  Bars as everywhere: merged pattern bit-exact,
operator values 1e-12, signals 1e-8; plus the limit the notebooks quote for this mode, exp(-b D0), through the driver."""
import numpy as np
import pytest

import bt_oracle as orc
from dmri_fem_cloud_b200 import btfem, meshes, periodic

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _cases():
    rng = np.random.default_rng(1)
    xyz, tets, ph = meshes.box_with_sphere(4.0, 6, 2.5)
    inner = (np.abs(xyz) < 3.9).all(axis=1)
    xyz = xyz.copy()
    xyz[inner] += 0.05 * rng.standard_normal((inner.sum(), 3))
    yield "cell_in_box_2c_pxy", xyz, tets, ph, [1, 1, 0], np.where(ph == 1, 1e-3, 2e-3), 5e-5
    xyz2, tets2 = meshes.box_mesh((-2, -1.5, -1), (2, 1.5, 1), 5, 4, 3)
    yield "box_1c_pxyz", xyz2, tets2, None, [1, 1, 1], 2e-3, 0.0
    n = 8
    xs = np.linspace(-2, 2, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    xy = np.column_stack([X.ravel(), Y.ravel()])
    idx = np.arange((n + 1) ** 2).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)
    ph2 = (np.linalg.norm(xy[tris].mean(axis=1), axis=1) < 1.2).astype(np.int32)
    yield "square_2c_px_triangles", xy, tris, ph2, [1, 0, 0], 2e-3, 1e-4


CASES = list(_cases())


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_strong_periodic_operators_and_signal(case):
    _, xyz, cells, ph, pdir, D, kappa = case
    x3 = orc.as_xyz3(xyz)
    lo, hi, hmin, _ = orc.domain_sizes(x3, cells)
    vm = periodic.vertex_map(xyz, pdir, lo, hi, 1e-2 * hmin)
    assert np.array_equal(vm, orc.periodic_vertex_map(xyz, pdir, lo, hi, 1e-2 * hmin))
    ops = orc.assemble(xyz, cells, ph, D=D, kappa=kappa, vmaster=vm)
    g = np.array([1.0, 0.5, 0.0 if x3[:, 2].max() == 0 else 0.2])
    g /= np.linalg.norm(g)
    Wo, Go = orc.strong_operators(ops, g)
    seq = orc.pgse(1000.0, 3000.0)
    q, k = seq.q_from_b(1000.0), 100.0
    ref = orc.theta_solve_strong(ops, seq, q, g, k)
    ts = orc.time_grid(seq.T, k)
    Fn = np.array([seq.F(t) for t in ts])
    Fp = np.concatenate([[seq.F(0.0)], Fn[:-1]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, cells, ph)
        fem.set_diffusion(D)
        if ph is not None:
            fem.set_permeability(kappa)
        fem.set_periodic_map(vm)
        fem.assemble()
        assert fem.ndof == ops.ndof and fem.nnz == ops.nnz
        rp, ci = fem.pattern()
        assert np.array_equal(rp, ops.rowptr) and np.array_equal(ci, ops.colidx)            # merged pattern, bit-exact
        dv, dc = fem.dofmap()
        assert np.array_equal(dv, ops.dof_vertex) and np.array_equal(dc, ops.dof_comp)
        for name in ("M", "S", "Jx", "I"):
            want = getattr(ops, name).data
            if np.abs(want).max() > 0:
                assert _relmax(fem.values(name), want) <= 1e-12, name
        W, G = fem.strong_operators(g)
        Wd = np.zeros(ops.nnz)
        Gd = np.zeros(ops.nnz)
        rows = np.repeat(np.arange(ops.ndof), np.diff(ops.rowptr))
        Wd[:] = np.asarray(Wo[rows, ops.colidx]).ravel()
        Gd[:] = np.asarray(Go[rows, ops.colidx]).ravel()
        assert _relmax(W, Wd) <= 1e-12 and _relmax(G, Gd) <= 1e-12
        res = fem.solve(k, 0.5, q * Fn, q * Fp, g, rtol=1e-13, atol=1e-16)
        u = fem.solution()
        again = fem.solve(k, 0.5, q * Fn, q * Fp, g, rtol=1e-13, atol=1e-16)
        # switching the map off gives the ordinary path back on the same handle
        fem.set_periodic_map(None)
        fem.assemble()
        f = np.array([seq.f(t) for t in ts])
        plain = fem.solve(k, 0.5, q * f, q * np.concatenate([[f[0]], f[:-1]]), g, rtol=1e-13, atol=1e-16)
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert _relmax(u, ref["u"]) <= 1e-8
    assert again["signal"] == res["signal"]                                                  # reproducible
    neu = orc.theta_solve(orc.assemble(xyz, cells, ph, D=D, kappa=kappa), seq, q, g, k, solver="lu")
    assert abs(plain["signal"] - neu["signal"]) <= 1e-8 * abs(neu["signal"])


def test_driver_periodic_along_the_axis_gives_free_diffusion():
    """`mydomain.PeriodicDir = [1, 0, 0]: s=exp(-bvalue*D0)` (T2_Relaxation.ipynb / MultilayeredStructures.ipynb /
    DiscontinuousInitialCondition.ipynb cell 12) through MyDomain / MRI_simulation.solve with IsDomainPeriodic = True:
    layered cylinder, gradient and periodicity along its axis, membranes parallel to the gradient.  Same meshes and
    tolerances as the oracle's pin (tests/test_oracle_strong.py); Neumann ends give the restricted value instead."""
    import sympy as sp
    from dmri_fem_cloud_b200 import dmrifemlib as dl
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 5.0, (2, 1, 1), 12, 4)      # axis z
    ph = (marker % 2).astype(np.int32)
    mesh = dl.Mesh(xyz, tets)

    def run(b, k, strong):
        mp = dl.MRI_parameters()
        mp.bvalue = b
        mp.delta, mp.Delta = 10000.0, 10000.0
        mp.T = mp.delta + mp.Delta
        mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
        mp.set_gradient_dir(mesh, 0, 0, 1)
        mp.Apply()
        sim = dl.MRI_simulation()
        sim.k = k
        sim.verbose = False
        md = dl.MyDomain(mesh, mp)
        md.phase, md.IsDomainMultiple, md.kappa = ph, True, 1e-5
        md.PeriodicDir, md.IsDomainPeriodic = ([0, 0, 1], True) if strong else ([0, 0, 0], False)
        md.Apply()
        md.D0 = 3e-3
        md.D = md.D0
        ls = dl.KrylovSolver("bicgstab", "jacobi")
        ls.parameters["relative_tolerance"] = 1e-10
        ls.parameters["absolute_tolerance"] = 1e-12
        sim.solve(md, mp, ls)
        s = sim.stats["signal"] / sim.stats["voi"]
        sim.fem.close()
        return s

    for b, k, tol in ((1000.0, 200.0, 3e-3), (1000.0, 50.0, 3e-4), (3000.0, 50.0, 2e-3)):
        assert abs(run(b, k, True) - np.exp(-b * 3e-3)) <= tol * np.exp(-b * 3e-3)
    assert run(1000.0, 200.0, False) > 0.9
