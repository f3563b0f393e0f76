"""GPU parity on triangle meshes (gdim-2 disks as ArbitraryTimeSequence.ipynb / T2_Relaxation.ipynb /
MultilayeredDiskVariablePermeability.ipynb run them, and a surface in 3-D as Manifolds.ipynb): the CUDA path through
btfem_set_mesh_tri against the oracle (pinned on triangles by tests/test_oracle_2d.py).  Same bars as the
tetrahedral tests: pattern bit-exact, values 1e-12, SpMV 1e-13, signals 1e-8."""
import numpy as np
import pytest
import sympy as sp

import bt_oracle as orc
from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes, periodic

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _rect(nx, ny, lx=4.0, ly=3.0):
    xs, ys = np.linspace(-lx / 2, lx / 2, nx + 1), np.linspace(-ly / 2, ly / 2, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    xy = np.column_stack([X.ravel(), Y.ravel()])
    idx = np.arange((nx + 1) * (ny + 1)).reshape(nx + 1, ny + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    return xy, np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)]).astype(np.int32)


def _cases():
    rng = np.random.default_rng(5)
    xy, tris, _ = meshes.disk_triangulation((5.0,), (6,), 24)
    xy = xy + 0.03 * rng.standard_normal(xy.shape)
    yield "disk_1c", xy, tris, None, dict(D=2e-3)
    xy2, tris2, lay = meshes.disk_triangulation((5.0, 7.5, 10.0), (4, 2, 2), 20)
    yield "layered_disk_2c", xy2, tris2, (lay % 2).astype(np.int32), dict(
        D=np.array([3e-3, 1e-3, 3e-3])[lay], kappa=1e-4, invT2=np.array([1e-16, 2.5e-5, 2.5e-5])[lay])
    xy3, tris3 = _rect(7, 5)
    c, s = np.cos(0.6), np.sin(0.6)
    Rm = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]) @ np.array([[1, 0, 0], [0, c, -s], [0, s, c]])
    xyz3 = orc.as_xyz3(xy3) @ Rm.T + np.array([0.5, -0.25, 1.0])
    Dt = np.broadcast_to(np.array([[3e-3, 1e-4, 2e-4], [1e-4, 2e-3, 0], [2e-4, 0, 1e-3]]), (len(tris3), 3, 3)).copy()
    yield "surface_in_3d_tensorD", xyz3, tris3, None, dict(D=Dt)


CASES = list(_cases())


def _setup(fem, xyz, tris, phase, co):
    fem.set_mesh(xyz, tris, phase)
    fem.set_diffusion(co.get("D", 1.0))
    fem.set_relaxation(co.get("invT2", 0.0))
    if phase is not None:
        fem.set_permeability(co.get("kappa", 0.0))
    fem.assemble()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_triangle_pattern_values_spmv(case):
    _, xyz, tris, phase, co = case
    ops = orc.assemble(xyz, tris, phase, D=co.get("D", 1.0), invT2=co.get("invT2", 0.0), kappa=co.get("kappa", 0.0))
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tris, phase, co)
        assert fem.ndof == ops.ndof and fem.nnz == ops.nnz
        rp, ci = fem.pattern()
        assert np.array_equal(rp, ops.rowptr) and np.array_equal(ci, ops.colidx)          # bit-exact
        if phase is None:
            rps, cis = orc.scalar_pattern(len(xyz), tris)
            assert np.array_equal(rp, rps) and np.array_equal(ci, cis)
        dv, dc = fem.dofmap()
        assert np.array_equal(dv, ops.dof_vertex) and np.array_equal(dc, ops.dof_comp)
        for name in ("M", "S", "R", "Jx", "Jy", "Jz", "I"):
            got, want = fem.values(name), getattr(ops, name).data
            if np.max(np.abs(want)) == 0:
                assert np.max(np.abs(got)) == 0, name
            else:
                assert _relmax(got, want) <= 1e-12, (name, _relmax(got, want))
        assert _relmax(fem.lumped_mass(), ops.lumped) <= 1e-13
        hmin, hmax = fem.mesh_stats()
        _, _, h0, h1 = orc.domain_sizes(orc.as_xyz3(xyz), tris)
        assert abs(hmin - h0) <= 1e-14 * h0 and abs(hmax - h1) <= 1e-14 * h1
        rng = np.random.default_rng(3)
        x = rng.standard_normal(ops.ndof) + 1j * rng.standard_normal(ops.ndof)
        g = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
        dt, theta, c = 200.0, 0.5, 1.5e-5
        want = (ops.M / dt + theta * (ops.S + ops.R + ops.I)) @ x + \
            1j * theta * c * ((g[0] * ops.Jx + g[1] * ops.Jy + g[2] * ops.Jz) @ x)
        for lanes in (0, 8):
            fem.set_lanes(lanes)
            assert _relmax(fem.spmv(dt, theta, c, g, x), want) <= 1e-13


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_triangle_theta_loop_signal(case):
    _, xyz, tris, phase, co = case
    ops = orc.assemble(xyz, tris, phase, D=co.get("D", 1.0), invT2=co.get("invT2", 0.0), kappa=co.get("kappa", 0.0))
    seq = orc.pgse(2000.0, 6000.0)
    q = seq.q_from_b(1500.0)
    k = 200.0
    g = np.array([0.6, 0.8, 0.0])
    ref = orc.theta_solve(ops, seq, q, g, k, solver="lu")
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tris, phase, co)
        res = fem.solve(k, 0.5, q * f, q * fp, g, rtol=1e-13, atol=1e-16)
        u = fem.solution()
    assert abs(res["voi"] - ref["voi"]) <= 1e-12 * abs(ref["voi"])
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])      # north-star tolerance
    assert _relmax(u, ref["u"]) <= 1e-8


def test_planar_weak_pseudo_periodic():
    """-pdir 1 1 0 on a two-compartment rectangle with non-matching opposite faces: boundary facets are edges."""
    xy, tris = _rect(8, 6)
    rng = np.random.default_rng(2)
    on = np.abs(xy[:, 0] - 2.0) < 1e-12
    inner = on & (np.abs(xy[:, 1]) < 1.4)
    xy[inner, 1] += 0.08 * rng.standard_normal(inner.sum())
    phase = (np.linalg.norm(xy[tris].mean(axis=1), axis=1) < 1.0).astype(np.int32)
    pdir = [1, 1, 0]
    xyz = orc.as_xyz3(xy)
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tris)
    bm = orc.periodic_marker(xyz, pdir, lo, hi, hmin)
    ops = orc.assemble(xy, tris, phase, D=2e-3, kappa=1e-5, bnd_kappa_vertex=bm)
    seq = orc.pgse(2000.0, 5000.0)
    q = seq.q_from_b(800.0)
    g = np.array([1.0, 0.5, 0.0]) / np.linalg.norm([1.0, 0.5, 0.0])
    k = 200.0
    per = orc.periodic_term(xyz, tris, ops, pdir, lo, hi, q, g, 0.5)
    ref = orc.theta_solve(ops, seq, q, g, k, solver="lu", periodic=per)
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    Fp = np.concatenate([[seq.F(0.0)], [seq.F(t) for t in ts[:-1]]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xy, tris, phase)
        fem.set_diffusion(2e-3)
        fem.set_permeability(1e-5)
        fem.set_periodic(pdir, 3e-3 / hmin, 1e-2 * hmin, lo, hi)
        fem.assemble()
        assert np.array_equal(fem.pattern()[1], ops.colidx)
        assert _relmax(fem.values("B"), ops.B.data) <= 1e-12
        dv, dc = fem.dofmap()
        bf = fem.boundary_facets()
        assert (bf[:, 2] == -1).all() and (bf[:, :2] >= 0).all()
        g_host = periodic.build_gather(xy, tris, phase, pdir, lo, hi, dv, dc)
        g_dev = periodic.build_gather(xy, tris, phase, pdir, lo, hi, dv, dc, bfacets=bf)
        for x1, x2 in zip(g_host, g_dev):
            assert np.array_equal(x1, x2)
        fem.set_periodic_gather(*g_dev)
        res = fem.solve(k, 0.5, q * f, q * fp, g, q=q, Fb=Fp, rtol=1e-13, atol=1e-16)
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    neu = orc.theta_solve(orc.assemble(xy, tris, phase, D=2e-3, kappa=1e-5), seq, q, g, k, solver="lu")
    assert abs(neu["signal"] - ref["signal"]) > 1e-3 * abs(ref["signal"])


def test_config3_on_the_2d_disk_through_the_driver(tmp_path, monkeypatch):
    """BASELINE configs[2] as the reference runs it: the multilayered DISK (gdim 2), variable permeability by
    marker pair, per-layer D and T2, cos-OGSE -- through Mesh / MyDomain / MRI_simulation / PostProcessing; and the
    published matrix-formalism value for PGSE on the same disk (T2_Relaxation.ipynb cell 12: 0.4777 at b=1000)."""
    monkeypatch.chdir(tmp_path)
    xy, tris, marker = meshes.disk_triangulation((5.0, 7.5, 10.0), (6, 3, 3), 48)
    phase = (marker % 2).astype(np.int32)
    mesh = dl.Mesh(xy, tris)
    assert mesh.geometry().dim() == 2 and mesh.topology().dim() == 2
    mp = dl.MRI_parameters()
    mp.bvalue = 1000
    mp.delta, mp.Delta = 4000.0, 4000.0
    t0 = 100.0
    Dd = mp.Delta + mp.delta
    mp.T = Dd + t0 + 100.0
    omega = 2.0 * np.pi / mp.delta
    tau = Dd / 2.0
    mp.fs_sym = sp.Piecewise((0., mp.s < t0), (sp.cos(omega * (mp.s - t0)), mp.s <= mp.delta + t0),
                             (0., mp.s <= tau + t0), (-sp.cos(omega * (mp.s - t0 - tau)), mp.s <= mp.delta + tau + t0),
                             (0., True))
    mp.set_gradient_dir(mesh, 1, 1, 5)                           # gdim 2: the z component is dropped (:813-814)
    assert abs(mp.gdir.array()[2]) == 0.0 and abs(np.linalg.norm(mp.gdir.array()) - 1.0) < 1e-15
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 100
    md = dl.MyDomain(mesh, mp)
    assert (md.gdim, md.tdim, md.zmin, md.zmax) == (2, 2, 0.0, 0.0)
    md.phase, md.IsDomainMultiple = phase, True
    kt = np.zeros((3, 3))
    kt[0, 1] = kt[1, 0] = 1e-4
    kt[1, 2] = kt[2, 1] = 1e-5
    md.kappa, md.kappa_marker = kt, marker
    md.T2_cell = np.array([4e16, 4e4, 4e4])[marker]
    md.Apply()
    Dl = np.array([3e-3, 1e-3, 3e-3])[marker]
    z = np.zeros(len(tris))
    md.ImposeDiffusionTensor(Dl, z, z, z, Dl, z, z, z, Dl)
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    ls.parameters["relative_tolerance"] = 1e-12
    ls.parameters["absolute_tolerance"] = 1e-15
    sim.solve(md, mp, ls)
    dl.PostProcessing(md, mp, sim, None, '')
    ops = orc.assemble(xy, tris, phase, D=Dl, invT2=1.0 / md.T2_cell,
                       kappa_facet=lambda fv, c0, c1: kt[marker[c0], marker[c1]])
    seq = orc.Sequence(mp.fs_sym, mp.T, mp.s)
    ref = orc.theta_solve(ops, seq, mp.qvalue, [1, 1, 0], 100.0, solver="lu")
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    # PGSE, uniform D and kappa: the matrix-formalism table
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xy, tris, phase)
        fem.set_diffusion(3e-3)
        fem.set_permeability(1e-5)
        fem.assemble()
        pg = orc.pgse(10000.0, 10000.0)
        ts = orc.time_grid(pg.T, 50.0)
        f = np.array([pg.f(t) for t in ts])
        fp = np.concatenate([[f[0]], f[:-1]])
        q = pg.q_from_b(1000.0)
        res = fem.solve(50.0, 0.5, q * f, q * fp, [0, 1, 0], rtol=1e-10, atol=1e-12)
    assert abs(res["signal"] / res["voi"] - 0.4777) <= 0.01 * 0.4777


def test_triangle_meshes_refuse_row_partitions():
    xy, tris, _ = meshes.disk_triangulation((5.0,), (3,), 12)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xy, tris)
        with pytest.raises(btfem.BTFemError):
            fem.set_partition(5, 3)


def test_segment_mesh_in_3d(tmp_path, monkeypatch):
    """Curves in 3-D (the neuron skeleton of Manifolds.ipynb: tdim 1, gdim 3) through btfem_set_mesh_seg and through
    the driver: pattern bit-exact (branch points included), values 1e-12, signal 1e-8 against the oracle."""
    from test_oracle_2d import _tree
    monkeypatch.chdir(tmp_path)
    xyz, segs = _tree(seed=9, nseg=300)
    ops = orc.assemble(xyz, segs, D=3e-3, invT2=1e-5)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, segs)
        fem.set_diffusion(3e-3)
        fem.set_relaxation(1e-5)
        fem.assemble()
        rp, ci = fem.pattern()
        rps, cis = orc.scalar_pattern(len(xyz), segs)
        assert np.array_equal(rp, rps) and np.array_equal(ci, cis)
        for name in ("M", "S", "R", "Jx", "Jy", "Jz"):
            assert _relmax(fem.values(name), getattr(ops, name).data) <= 1e-12, name
        assert np.max(np.abs(fem.values("I"))) == 0
        hmin, hmax = fem.mesh_stats()
        L = np.linalg.norm(xyz[segs[:, 1]] - xyz[segs[:, 0]], axis=1)
        assert abs(hmin - L.min()) <= 1e-14 * L.min() and abs(hmax - L.max()) <= 1e-14 * L.max()
        with pytest.raises(btfem.BTFemError):
            fem.set_phase(np.zeros(len(segs), dtype=np.int32))
    mesh = dl.Mesh(xyz, segs)
    assert (mesh.geometry().dim(), mesh.topology().dim()) == (3, 1)
    mp = dl.MRI_parameters()
    mp.bvalue = 2000
    mp.delta, mp.Delta = 2000.0, 6000.0
    mp.T = mp.delta + mp.Delta
    mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
    mp.set_gradient_dir(mesh, 1, 1, 1)
    mp.T2 = 1e5
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 200
    md = dl.MyDomain(mesh, mp)
    md.Apply()
    md.D0 = 3e-3
    md.D = md.D0
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    ls.parameters["relative_tolerance"] = 1e-13
    ls.parameters["absolute_tolerance"] = 1e-16
    sim.solve(md, mp, ls)
    seq = orc.pgse(2000.0, 6000.0)
    ref = orc.theta_solve(ops, seq, mp.qvalue, [1, 1, 1], 200.0, solver="lu")
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
