"""CPU: the oracle's restatement of the STRONGLY imposed pseudo-periodic BC (FuncF_sBC / outer_interface /
inner_interface / PeriodicBD, DmriFemLib.py:147-238, 327-375): the equation for u~ = u exp(+i q F(t) g.x) on a
periodic function space.  The reference tree holds no recorded output of this mode (no notebook runs
`IsDomainPeriodic = True` with a periodic direction), so the restatement is pinned by what must hold exactly or in
the limit: the uniform solution in a periodic box, equivalence with the untransformed equation as dt -> 0, and
invariance under tiling the unit cell."""
import numpy as np

import bt_oracle as orc
from dmri_fem_cloud_b200 import meshes


def _tile_x(xyz, tets, ph, L):
    x2 = xyz.copy()
    x2[:, 0] += L
    allx = np.vstack([xyz, x2])
    t2 = np.vstack([tets, tets + len(xyz)])
    key = np.round(allx * 1e6).astype(np.int64)
    _, first, inv = np.unique(key, axis=0, return_index=True, return_inverse=True)
    return allx[first], inv.ravel()[t2].astype(np.int32), np.concatenate([ph, ph])


def test_periodic_vertex_map_wraps_faces_edges_and_corners():
    xyz, tets = meshes.box_mesh((-2, -1.5, -1), (2, 1.5, 1), 4, 3, 2)
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tets)
    vm = orc.periodic_vertex_map(xyz, [1, 1, 1], lo, hi, 1e-2 * hmin)
    assert len(np.unique(vm)) == 4 * 3 * 2 and (vm[vm] == vm).all()
    corners = np.nonzero((np.abs(np.abs(xyz) - hi) < 1e-12).all(axis=1))[0]
    assert len(corners) == 8 and len(set(vm[corners])) == 1 and np.allclose(xyz[vm[corners[0]]], lo)
    vm_x = orc.periodic_vertex_map(xyz, [1, 0, 0], lo, hi, 1e-2 * hmin)
    assert len(np.unique(vm_x)) == 4 * 4 * 3 and np.allclose(xyz[vm_x][:, 1:], xyz[:, 1:])
    bad = xyz.copy()
    bad[np.argmax(bad[:, 0]), 1] += 0.2                     # a max-face vertex without a partner
    try:
        orc.periodic_vertex_map(bad, [1, 0, 0], lo, hi, 1e-2 * hmin)
        assert False
    except ValueError:
        pass


def test_uniform_solution_in_a_periodic_box_follows_the_scalar_recurrence():
    """No barriers, all directions periodic: u~ stays uniform (C 1 = 0, the facet terms of merged faces cancel), so
    every dof follows u <- u (1/k - theta D q^2 F_p^2) / (1/k + theta D q^2 F_n^2); the signal tends to exp(-bD)."""
    xyz, tets = meshes.box_mesh((-2, -1.5, -1), (2, 1.5, 1), 5, 4, 3)
    rng = np.random.default_rng(2)
    inner = (np.abs(xyz) < np.array([1.9, 1.4, 0.9])).all(axis=1)
    xyz[inner] += 0.05 * rng.standard_normal((inner.sum(), 3))
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tets)
    ops = orc.assemble(xyz, tets, D=2e-3, vmaster=orc.periodic_vertex_map(xyz, [1, 1, 1], lo, hi, 1e-2 * hmin))
    seq = orc.pgse(1000.0, 3000.0)
    q, k = seq.q_from_b(1500.0), 100.0
    g = np.array([0.3, -0.5, 0.8])
    r = orc.theta_solve_strong(ops, seq, q, g, k)
    u, tp = 1.0, 0.0
    for t in orc.time_grid(seq.T, k):
        u *= (1 / k - 0.5 * 2e-3 * (q * seq.F(tp)) ** 2) / (1 / k + 0.5 * 2e-3 * (q * seq.F(t)) ** 2)
        tp = t
    assert np.abs(r["u"] - u).max() <= 1e-13
    assert abs(r["signal"] / r["voi"] - np.exp(-1500.0 * 2e-3)) <= 0.01 * np.exp(-3.0)


def test_transformed_equation_equals_the_original_one_as_dt_goes_to_zero():
    """Without identification the transformed problem IS the Neumann problem (u = u~ at t = T where F = 0): the two
    discretisations differ by the time error of the lagged, discontinuous f (first order) and a small spatial
    term.  Two compartments with different D: exercises the signs of C and of the facet terms on both sides."""
    xy, tris, lay = meshes.disk_triangulation((2.0, 3.0), (6, 4), 32)
    ph = (lay % 2).astype(np.int32)
    ops = orc.assemble(xy, tris, ph, D=np.array([2e-3, 1e-3])[lay], kappa=5e-5)
    seq = orc.pgse(1000.0, 3000.0)
    q = seq.q_from_b(1000.0)
    g = np.array([0.6, 0.8, 0.0])
    diff = []
    for k in (50.0, 12.5):
        a = orc.theta_solve(ops, seq, q, g, k, solver="lu")
        b = orc.theta_solve_strong(ops, seq, q, g, k)
        diff.append(abs(b["signal"] - a["signal"]) / a["signal"])
    assert diff[0] < 1e-2 and diff[1] < 2e-3 and diff[1] < 0.3 * diff[0]
    # tetrahedra: the spatial part of the difference falls with h^2 (1.7e-2 at n = 6, 1.6e-3 at n = 12)
    err = []
    for n in (6, 12):
        xyz, tets = meshes.box_mesh((-2, -1.5, -1), (2, 1.5, 1), n, n, n // 2)
        ops = orc.assemble(xyz, tets, D=2e-3)
        a = orc.theta_solve(ops, seq, q, g, 50.0, solver="lu")
        b = orc.theta_solve_strong(ops, seq, q, g, 50.0)
        err.append(abs(b["signal"] - a["signal"]) / a["signal"])
    assert err[1] < 3e-3 and err[1] < 0.25 * err[0]


def test_tiling_the_unit_cell_does_not_change_the_signal():
    """Periodic lattice of permeable cells, D differs between the compartments, interior vertices jittered: the
    solution on two unit cells glued together equals the solution on one -- to rounding, because the discrete
    problem is translation invariant.  Pins the vertex identification and the facet terms on merged faces."""
    xyz, tets, ph = meshes.box_with_sphere(4.0, 6, 2.5)
    rng = np.random.default_rng(1)
    inner = (np.abs(xyz) < 3.9).all(axis=1)
    xyz = xyz.copy()
    xyz[inner] += 0.05 * rng.standard_normal((inner.sum(), 3))
    seq = orc.pgse(1000.0, 3000.0)
    q, k = seq.q_from_b(1000.0), 100.0
    g = np.array([1.0, 0.5, 0.2])
    pdir = [1, 1, 0]

    def run(x, t, p):
        lo, hi, hmin, _ = orc.domain_sizes(x, t)
        vm = orc.periodic_vertex_map(x, pdir, lo, hi, 1e-2 * hmin)
        ops = orc.assemble(x, t, p, D=np.where(p == 1, 1e-3, 2e-3), kappa=5e-5, vmaster=vm)
        return orc.theta_solve_strong(ops, seq, q, g, k), ops

    r1, o1 = run(xyz, tets, ph)
    r2, o2 = run(*_tile_x(xyz, tets, ph, 8.0))
    assert o2.ndof == 2 * o1.ndof
    assert abs(r1["signal"] / r1["voi"] - r2["signal"] / r2["voi"]) <= 1e-13
    neu = orc.theta_solve(orc.assemble(xyz, tets, ph, D=np.where(ph == 1, 1e-3, 2e-3), kappa=5e-5), seq, q, g, k,
                          solver="lu")
    assert abs(neu["signal"] / neu["voi"] - r1["signal"] / r1["voi"]) > 0.05      # periodicity matters here


def _host_harness(tmp_path):
    """g++ build of tests/strong_host_check.cpp: the SAME header the CUDA kernels include (csrc/strong_math.cuh)."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    so = str(tmp_path / "libstronghost.so")
    subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", os.path.join(here, "strong_host_check.cpp"),
                           "-o", so])
    return ctypes.CDLL(so)


def test_kernel_arithmetic_of_strong_mode_matches_oracle(tmp_path):
    """csrc/strong_math.cuh (what k_strong_cells / k_strong_facet_pairs / k_strong_recombine evaluate), compiled for
    the host, against oracle.strong_operators -- tetrahedra with a full tensor D per cell and a two-compartment
    triangle mesh.  Covers the formulas and signs of W, C and the facet terms; the sort / scatter plumbing around
    them is exercised by the (opt-in) GPU tests."""
    import ctypes as C
    import scipy.sparse as sps
    lib = _host_harness(tmp_path)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rng = np.random.default_rng(3)
    xyz, tets, ph = meshes.box_with_sphere(2.0, 3, 1.2)
    xyz = xyz + 0.05 * rng.standard_normal(xyz.shape)
    A = rng.normal(size=(len(tets), 3, 3))
    Dt = 1e-3 * (np.einsum("cab,cdb->cad", A, A) + 0.3 * rng.normal(size=(len(tets), 3, 3)))     # not symmetric
    xy, tris, lay = meshes.disk_triangulation((2.0, 3.0), (2, 2), 10)
    for x, cells, phase, D, dkind in ((xyz, tets, ph, Dt, 2), (orc.as_xyz3(xy), tris, (lay % 2).astype(np.int32),
                                                               np.array([2e-3, 1e-3])[lay], 1)):
        nc, nvc = cells.shape
        g = np.array([0.3, -0.5, 0.8 if nvc == 4 else 0.0])
        g /= np.linalg.norm(g)
        ops = orc.assemble(x, cells, phase, D=D, kappa=1e-5)
        Wo, Go = orc.strong_operators(ops, g)
        cells4 = -np.ones((nc, 4), dtype=np.int32)
        cells4[:, :nvc] = cells
        x = np.ascontiguousarray(x, dtype=np.float64)
        Dc = np.ascontiguousarray(D, dtype=np.float64)
        W = np.zeros((nc, 4, 4))
        Cm = np.zeros((nc, 4, 4))
        coef = np.zeros((nc, 4))
        lib.strong_host_cells(C.c_int64(nc), nvc, dp(x), ip(cells4), dkind, dp(Dc), dp(g), dp(W), dp(Cm))
        lib.strong_host_facets(C.c_int64(nc), nvc, dp(x), ip(cells4), dkind, dp(Dc), dp(g), dp(coef))
        n = ops.ndof
        rows = np.repeat(ops.cell_dofs, nvc, axis=1).ravel()
        cols = np.tile(ops.cell_dofs, (1, nvc)).ravel()
        Wk = sps.coo_matrix((W[:, :nvc, :nvc].ravel(), (rows, cols)), shape=(n, n)).tocsr()
        Ck = sps.coo_matrix((Cm[:, :nvc, :nvc].ravel(), (rows, cols)), shape=(n, n)).tocsr()
        # bounding facets the way k_strong_flag picks them: exterior, or the neighbour is in the other compartment
        f, cell, lf = orc.facets(cells)
        same_next = np.zeros(len(f), dtype=bool)
        same_next[:-1] = np.all(f[1:] == f[:-1], axis=1)
        same_prev = np.zeros(len(f), dtype=bool)
        same_prev[1:] = same_next[:-1]
        partner = np.where(same_next, np.roll(cell, -1), np.where(same_prev, np.roll(cell, 1), -1))
        bnd = (partner < 0) | (phase[np.maximum(partner, 0)] != phase[cell])
        rr, cc, vv = [], [], []
        for t, l in zip(cell[bnd], lf[bnd]):
            loc = [k for k in range(nvc) if k != l]                  # la = a < lf ? a : a + 1
            for a_ in range(nvc - 1):
                for b_ in range(nvc - 1):
                    rr.append(ops.cell_dofs[t, loc[a_]])
                    cc.append(ops.cell_dofs[t, loc[b_]])
                    vv.append(coef[t, l] * (2.0 if a_ == b_ else 1.0))
        Nk = sps.coo_matrix((vv, (rr, cc)), shape=(n, n)).tocsr()
        assert abs(Wk - Wo).max() <= 1e-15 * abs(Wo).max()
        assert abs((Ck - Nk) - Go).max() <= 1e-13 * abs(Go).max()
    out = np.zeros(3)
    lib.strong_host_combine(C.c_double(2.0), C.c_double(0.5), C.c_double(0.25), C.c_double(-0.75), C.c_double(4.0),
                            C.c_double(8.0), C.c_double(0.1), dp(out))
    assert np.allclose(out, [(2.0 + 0.5 + 4.0 * 0.25) * 0.1, (2.0 - 0.5 - 8.0 * 0.25) * 0.1, -0.075], rtol=1e-15)


def test_periodic_along_the_axis_gives_free_diffusion_as_the_notebooks_state():
    """T2_Relaxation.ipynb / MultilayeredStructures.ipynb / DiscontinuousInitialCondition.ipynb cell 12 quote, for the
    layered cylinder with the gradient along its axis: `mydomain.PeriodicDir = [1, 0, 0]: s=exp(-bvalue*D0)` (and the
    restricted values for PeriodicDir = [0, 0, 0]).  The strongly periodic oracle reproduces the quoted limit up to
    the time-discretisation error, membranes and all (they are parallel to the gradient); Neumann ends give the
    restricted value instead."""
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 5.0, (2, 1, 1), 12, 4)      # axis z
    ph = (marker % 2).astype(np.int32)
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tets)
    seq = orc.pgse(10000.0, 10000.0)
    g = [0, 0, 1]
    vm = orc.periodic_vertex_map(xyz, [0, 0, 1], lo, hi, 1e-2 * hmin)
    ops = orc.assemble(xyz, tets, ph, D=3e-3, kappa=1e-5, vmaster=vm)
    for b, k, tol in ((1000.0, 200.0, 3e-3), (1000.0, 50.0, 3e-4), (3000.0, 50.0, 2e-3)):
        r = orc.theta_solve_strong(ops, seq, seq.q_from_b(b), g, k)
        assert abs(r["signal"] / r["voi"] - np.exp(-b * 3e-3)) <= tol * np.exp(-b * 3e-3)
    neu = orc.theta_solve(orc.assemble(xyz, tets, ph, D=3e-3, kappa=1e-5), seq, seq.q_from_b(1000.0), g, 200.0, solver="lu")
    assert neu["signal"] / neu["voi"] > 0.9
