"""GPU parity tests proper: the CUDA path (through the C-ABI) against the oracle on the same
seeded inputs.  Integer pattern bit-exact; matrix values <= 1e-12 relative (north star);
SpMV <= 1e-13; final signals <= 1e-8 relative (north star) -- tolerances written at each assert."""
import os

import numpy as np
import pytest

import bt_oracle as orc
from conftest import GOLDEN
from dmri_fem_cloud_b200 import btfem, meshes

pytestmark = pytest.mark.gpu


def _relmax(a, b):
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)


def _cases():
    rng = np.random.default_rng(7)
    xyz, tets = meshes.box_mesh((-2.5,) * 3, (2.5,) * 3, 5, 4, 3)
    xyz = xyz + 0.05 * rng.standard_normal(xyz.shape)          # break the symmetry
    yield "box_1c", xyz, tets, None, dict(D=2e-3)
    xyz2, tets2 = meshes.cylinder(3.0, 10.0, nr=3, nsec=10, nz=6)
    xyz2, tets2 = meshes.shuffle_vertices(xyz2, tets2, seed=1)
    yield "cyl_1c_tensorD_T2", xyz2, tets2, None, dict(
        D=np.broadcast_to(np.array([[3e-3, 1e-4, 0], [1e-4, 2e-3, 0], [0, 0, 1e-3]]), (len(tets2), 3, 3)).copy(),
        invT2=rng.uniform(1e-5, 1e-4, len(tets2)))
    xyz3, tets3, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 5.0, (3, 2, 2), 12, 2)
    yield "layered_2c", xyz3, tets3, (marker % 2).astype(np.int32), dict(
        D=np.array([3e-3, 1e-3, 3e-3])[marker], kappa=1e-5)
    xyz4, tets4, ph4 = meshes.box_with_sphere(10.0, 6, 5.0)
    yield "sphere_in_box_2c", xyz4, tets4, ph4, dict(D=3e-3, kappa=5e-5, invT2=1.0 / 4e4)


CASES = list(_cases())


def _setup_inputs(fem, xyz, tets, phase, co):
    fem.set_mesh(xyz, tets, phase)
    fem.set_diffusion(co.get("D", 1.0))
    fem.set_relaxation(co.get("invT2", 0.0))
    if phase is not None:
        fem.set_permeability(co.get("kappa", 0.0))


def _setup(fem, xyz, tets, phase, co):
    _setup_inputs(fem, xyz, tets, phase, co)
    fem.assemble()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_pattern_and_values(case):
    _, xyz, tets, phase, co = case
    ops = orc.assemble(xyz, tets, phase, D=co.get("D", 1.0), invT2=co.get("invT2", 0.0), kappa=co.get("kappa", 0.0))
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        assert fem.ndof == ops.ndof and fem.nnz == ops.nnz
        rp, ci = fem.pattern()
        assert np.array_equal(rp, ops.rowptr)            # bit-exact
        assert np.array_equal(ci, ops.colidx)            # bit-exact
        dv, dc = fem.dofmap()
        assert np.array_equal(dv, ops.dof_vertex) and np.array_equal(dc, ops.dof_comp)
        if phase is None:                                 # one compartment: the reference's scalar P1 pattern
            rps, cis = orc.scalar_pattern(len(xyz), tets)
            assert np.array_equal(rp, rps) and np.array_equal(ci, cis)
        for name in ("M", "S", "R", "Jx", "Jy", "Jz", "I"):
            got, want = fem.values(name), getattr(ops, name).data
            if np.max(np.abs(want)) == 0:
                assert np.max(np.abs(got)) == 0, name
            else:
                assert _relmax(got, want) <= 1e-12, (name, _relmax(got, want))
        assert _relmax(fem.lumped_mass(), ops.lumped) <= 1e-13


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_fused_spmv(case):
    _, xyz, tets, phase, co = case
    ops = orc.assemble(xyz, tets, phase, D=co.get("D", 1.0), invT2=co.get("invT2", 0.0), kappa=co.get("kappa", 0.0))
    rng = np.random.default_rng(3)
    x = rng.standard_normal(ops.ndof) + 1j * rng.standard_normal(ops.ndof)
    g = np.array([0.3, -0.5, 0.8])
    g /= np.linalg.norm(g)
    dt, theta, c = 200.0, 0.5, 1.5e-5
    P = ops.M / dt + theta * (ops.S + ops.R + ops.I + ops.B)
    Jg = g[0] * ops.Jx + g[1] * ops.Jy + g[2] * ops.Jz
    want = P @ x + 1j * theta * c * (Jg @ x)
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        for lanes in (0, 4, 8, 16, 32):          # 0 = stream variant (default)
            fem.set_lanes(lanes)
            got = fem.spmv(dt, theta, c, g, x)
            assert _relmax(got, want) <= 1e-13, (lanes, _relmax(got, want))


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_theta_loop_signal(case):
    """Whole path: GPU Jacobi-BiCGStab at tight tolerance vs the oracle's exact (LU) time stepping."""
    _, xyz, tets, phase, co = case
    ops = orc.assemble(xyz, tets, phase, D=co.get("D", 1.0), invT2=co.get("invT2", 0.0), kappa=co.get("kappa", 0.0))
    seq = orc.pgse(2000.0, 6000.0)
    q = seq.q_from_b(1000.0)
    k = 200.0
    g = np.array([0.0, 1.0, 0.0])
    ref = orc.theta_solve(ops, seq, q, g, k, solver="lu")
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        res = fem.solve(k, 0.5, q * f, q * fp, g, rtol=1e-13, atol=1e-16, want_iters=True)
        u = fem.solution()
    assert res["n_steps"] == ref["n_steps"]
    assert abs(res["voi"] - ref["voi"]) <= 1e-12 * abs(ref["voi"])
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])      # north-star tolerance
    assert _relmax(u, ref["u"]) <= 1e-8
    assert res["total_iters"] > 0 and res["last_reason"] > 0


def test_bicgstab_matches_petsc_restatement():
    """Same Krylov: iteration counts of the GPU solver equal the oracle's PETSc-BCGS restatement
    at the CLI tolerances (GCloudDmriSolver.py:219-222), and the signals agree to 1e-8."""
    _, xyz, tets, phase, co = CASES[0]
    ops = orc.assemble(xyz, tets, phase, D=co["D"])
    seq = orc.pgse(2000.0, 6000.0)
    q = seq.q_from_b(1000.0)
    k = 200.0
    g = np.array([1.0, 0.0, 0.0])
    ref = orc.theta_solve(ops, seq, q, g, k, solver="bicgstab", rtol=1e-9, atol=1e-10)
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        res = fem.solve(k, 0.5, q * f, q * fp, g, rtol=1e-9, atol=1e-10, want_iters=True)
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    # identical recurrence; rounding may shift an iteration here and there
    assert np.max(np.abs(res["iters"].astype(int) - ref["iters"].astype(int))) <= 1
    assert np.mean(res["iters"] == ref["iters"]) > 0.8


def test_golden_convergence_box():
    """Reference-recorded output: ConvergenceTest.ipynb cell 10, BoxMesh n=16 -> 8.440078e-01
    (loop `t < T`, 1100 steps, dt=10, D=2e-3, delta=1000, Delta=10000, g=z, b=1000)."""
    xyz, tets = meshes.box_mesh((-2.5,) * 3, (2.5,) * 3, 16, 16, 16)
    seq = orc.pgse(1000.0, 10000.0)
    q = seq.q_from_b(1000.0)
    k = 10.0
    ts = orc.time_grid(seq.T, k, closed=False)
    assert len(ts) == 1100
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets)
        fem.set_diffusion(2e-3)
        fem.assemble()
        res = fem.solve(k, 0.5, q * f, q * fp, [0, 0, 1.0], rtol=1e-12, atol=1e-15)
    s = res["signal"] / res["voi"]
    assert "%.6e" % s == "8.440078e-01"
    gold = np.load(os.path.join(GOLDEN, "convergence_box_n16.npz"))
    assert abs(s - float(gold["normalized_signal_lu"])) <= 1e-8 * s


def test_errors_are_reported():
    with btfem.BTFem(0) as fem:
        with pytest.raises(btfem.BTFemError):
            fem.assemble()                                   # no mesh yet
        xyz, tets = meshes.box_mesh((0,) * 3, (1,) * 3, 2, 2, 2)
        bad = tets.copy()
        bad[0, 0] = 10 ** 6
        with pytest.raises(btfem.BTFemError):
            fem.set_mesh(xyz, bad)
        fem.set_mesh(xyz, tets)
        fem.set_diffusion(1e-3)
        fem.assemble()
        with pytest.raises(btfem.BTFemError) as ei:
            fem.solve(1.0, 0.5, np.ones(5) * 1e-3, np.ones(5) * 1e-3, [1, 0, 0], rtol=1e-30, atol=0.0, maxit=3)
        assert ei.value.code == -3                           # KSP_DIVERGED_ITS


def test_full_size_properties():
    """At bench size the oracle is too slow; use size-independent properties: q=0 reduces the
    path to pure diffusion with Neumann BC, which conserves total magnetisation (signal == voi),
    and the SpMV is linear."""
    xyz, tets = meshes.box_mesh((-5,) * 3, (5,) * 3, 40, 40, 40)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets)
        fem.set_diffusion(3e-3)
        fem.assemble()
        n = 50
        res = fem.solve(200.0, 0.5, np.zeros(n), np.zeros(n), [1, 0, 0], rtol=1e-12, atol=1e-15)
        assert abs(res["signal"] - res["voi"]) <= 1e-9 * res["voi"]
        assert abs(res["whole_vol"] - 1000.0) <= 1e-9 * 1000.0
        rng = np.random.default_rng(0)
        x = rng.standard_normal(fem.ndof) + 1j * rng.standard_normal(fem.ndof)
        y = rng.standard_normal(fem.ndof) + 1j * rng.standard_normal(fem.ndof)
        g = [0, 0, 1.0]
        a = fem.spmv(200.0, 0.5, 2e-5, g, x)
        b = fem.spmv(200.0, 0.5, 2e-5, g, y)
        ab = fem.spmv(200.0, 0.5, 2e-5, g, 2.0 * x - 3.0 * y)
        assert _relmax(ab, 2.0 * a - 3.0 * b) <= 1e-13


def _periodic_cases():
    rng = np.random.default_rng(11)
    xyz, tets = meshes.box_mesh((-5, -2, -2), (5, 2, 2), 10, 4, 4)
    yield "box_1c_px", xyz, tets, None, [1, 0, 0]
    # opposite x-faces made non-matching: tangential jitter of the vertices on the max face only
    xyz2 = xyz.copy()
    on = np.abs(xyz2[:, 0] - 5.0) < 1e-12
    inner = on & (np.abs(xyz2[:, 1]) < 1.9) & (np.abs(xyz2[:, 2]) < 1.9)
    xyz2[inner, 1:] += 0.15 * rng.standard_normal((inner.sum(), 2))
    yield "box_1c_px_nonmatching", xyz2, tets, None, [1, 0, 0]
    xyz3, tets3, ph3 = meshes.ecs_slab(12, 12, 2, lx=12.0, ly=12.0, lz=1.0, ncyl=4, rmin=2.0, rmax=3.0, seed=3)
    yield "ecs_2c_pxy", xyz3, tets3, ph3, [1, 1, 0]


PCASES = list(_periodic_cases())


@pytest.mark.parametrize("case", PCASES, ids=[c[0] for c in PCASES])
def test_weak_pseudo_periodic(case):
    """-pdir: artificial-permeability boundary matrix B and the lagged u_bc term (DmriFemLib.py:58-93,
    256-324) against the oracle's independent restatement."""
    from dmri_fem_cloud_b200 import periodic
    _, xyz, tets, phase, pdir = case
    lo, hi, hmin, hmax = orc.domain_sizes(xyz, tets)
    kappa_e, tol = 3e-3 / hmin, 1e-2 * hmin
    bm = orc.periodic_marker(xyz, pdir, lo, hi, hmin)
    ops = orc.assemble(xyz, tets, phase, D=2e-3, kappa=1e-5, bnd_kappa_vertex=bm)
    seq = orc.pgse(2000.0, 5000.0)
    q = seq.q_from_b(800.0)
    g = np.array([1.0, 0.5, 0.0])
    g /= np.linalg.norm(g)
    k = 200.0
    per = orc.periodic_term(xyz, tets, ops, pdir, lo, hi, q, g, 0.5)
    ref = orc.theta_solve(ops, seq, q, g, k, solver="lu", periodic=per)
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    Fp = np.concatenate([[seq.F(0.0)], [seq.F(t) for t in ts[:-1]]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        fem.set_diffusion(2e-3)
        if phase is not None:
            fem.set_permeability(1e-5)
        fem.set_periodic(pdir, kappa_e, tol, lo, hi)
        fem.assemble()
        assert np.array_equal(fem.pattern()[1], ops.colidx)
        assert _relmax(fem.values("B"), ops.B.data) <= 1e-12
        dv, dc = fem.dofmap()
        g_host = periodic.build_gather(xyz, tets, phase, pdir, lo, hi, dv, dc)
        g_dev = periodic.build_gather(xyz, tets, phase, pdir, lo, hi, dv, dc, bfacets=fem.boundary_facets())
        for x1, x2 in zip(g_host, g_dev):                    # GPU-found facets give the same operator
            assert np.array_equal(x1, x2)
        fem.set_periodic_gather(*g_dev)
        res = fem.solve(k, 0.5, q * f, q * fp, g, q=q, Fb=Fp, rtol=1e-13, atol=1e-16)
        u = fem.solution()
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert _relmax(u, ref["u"]) <= 1e-8
    # the BC does something: the Neumann answer is different
    neu = orc.theta_solve(orc.assemble(xyz, tets, phase, D=2e-3, kappa=1e-5), seq, q, g, k, solver="lu")
    assert abs(neu["signal"] - ref["signal"]) > 1e-3 * abs(ref["signal"])


@pytest.mark.parametrize("ksp,pc,nz", [("gmres", "jacobi", False), ("gmres", "none", False), ("gmres", "jacobi", True),
                                       ("bicgstab", "none", False), ("bicgstab", "jacobi", True)])
def test_other_krylov_choices(ksp, pc, nz):
    """The reference's other solver settings: GMRES(30) (RealNeurons*.ipynb, rtol 1e-4,
    nonzero initial guess), BiCGStab without PC (comri multilayer main.cpp:237).  Iteration counts follow
    the oracle's PETSc restatement; signals agree with the exact (LU) stepping within the solver tolerance."""
    _, xyz, tets, phase, co = CASES[3]
    ops = orc.assemble(xyz, tets, phase, D=co["D"], invT2=co["invT2"], kappa=co["kappa"])
    seq = orc.pgse(2000.0, 6000.0)
    q = seq.q_from_b(1000.0)
    k = 200.0
    g = np.array([0.0, 0.6, 0.8])
    name = ksp + ("_none" if pc == "none" else "")
    ref = orc.theta_solve(ops, seq, q, g, k, solver=name, rtol=1e-10, atol=1e-14, nonzero_guess=nz, restart=30)
    lu = orc.theta_solve(ops, seq, q, g, k, solver="lu")
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        res = fem.solve(k, 0.5, q * f, q * fp, g, ksp=ksp, pc=pc, rtol=1e-10, atol=1e-14, nonzero_guess=nz,
                        restart=30, want_iters=True)
    assert abs(res["signal"] - lu["signal"]) <= 1e-8 * abs(lu["signal"])
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert np.max(np.abs(res["iters"].astype(int) - ref["iters"].astype(int))) <= 2
    assert np.mean(np.abs(res["iters"].astype(int) - ref["iters"].astype(int)) <= 1) > 0.9


def test_gmres_restart_path():
    """Small restart forces several GMRES cycles (true-residual recomputation at each restart)."""
    _, xyz, tets, phase, co = CASES[0]
    ops = orc.assemble(xyz, tets, phase, D=co["D"])
    seq = orc.pgse(2000.0, 6000.0)
    q = seq.q_from_b(1000.0)
    k = 400.0
    g = np.array([1.0, 0.0, 0.0])
    ref = orc.theta_solve(ops, seq, q, g, k, solver="gmres", rtol=1e-11, atol=1e-15, restart=5)
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        res = fem.solve(k, 0.5, q * f, q * fp, g, ksp="gmres", rtol=1e-11, atol=1e-15, restart=5, want_iters=True)
    assert ref["iters"].max() > 5
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert np.max(np.abs(res["iters"].astype(int) - ref["iters"].astype(int))) <= 2


def test_batched_solves_equal_individual_solves(monkeypatch):
    """btfem_solve_batch on the kernel chain (BTFEM_BATCH_PERSIST=0; the path of batches of more than 16 members):
    members (different directions and q) advanced in lock step.  A member gets exactly the same bits whatever batch it
    travels in (same kernels, same reduction order per member); the one-at-a-time solves run the persistent kernel
    (other grouping of the dot-product partial sums), so they agree to solver tolerance."""
    monkeypatch.setenv("BTFEM_BATCH_PERSIST", "0")
    _, xyz, tets, phase, co = CASES[3]
    seq = orc.pgse(2000.0, 6000.0)
    k = 200.0
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    dirs = meshes.fibonacci_hemisphere(3)
    qs = [seq.q_from_b(b) for b in (500.0, 3000.0)]
    members = [(q * f, q * fp, d) for d in dirs for q in qs]
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        single = [fem.solve(k, 0.5, cA, cb, g, rtol=1e-10, atol=1e-14) for cA, cb, g in members]
        batch = fem.solve_batch(k, 0.5, members, rtol=1e-10, atol=1e-14)
        split = fem.solve_batch(k, 0.5, members[:2], rtol=1e-10, atol=1e-14) + \
            fem.solve_batch(k, 0.5, members[2:], rtol=1e-10, atol=1e-14)
        again = fem.solve(k, 0.5, *members[1], rtol=1e-10, atol=1e-14)      # back to a batch of one
    for sb, ss in zip(batch, split):
        assert sb["signal"] == ss["signal"]                                  # bit-identical
        assert sb["total_iters"] == ss["total_iters"]
    for s1, sb in zip(single, batch):
        assert abs(s1["signal"] - sb["signal"]) <= 1e-9 * abs(sb["signal"])
        assert abs(s1["total_iters"] - sb["total_iters"]) <= max(3, 0.02 * sb["total_iters"])
    assert again["signal"] == single[1]["signal"]                            # reproducible
    assert len({round(s["signal"], 6) for s in batch}) == len(batch)         # the members really differ


FIX = np.load(os.path.join(GOLDEN, "fixture_meshes.npz"))


@pytest.mark.parametrize("name,gdir,two", [("cyl6_r_3E_6_vol", [1, 0, 0], False), ("cyl12_r_3E_6_vol", [1, 0, 0], False),
                                          ("torus", [0, 1, 0], True)])
def test_reference_mesh_fixtures(name, gdir, two):
    """The reference's own meshes (comri/meshes/*.zip, stored in tests/golden/fixture_meshes.npz with the
    oracle's exact-stepping signals): BASELINE configs[0] (-M 0 ... -gdir 1 0 0) and configs[1]
    (-M 1 -b 1000 -p 1e-5 -k 200 -gdir 0 1 0) at the CLI's Krylov tolerances."""
    xyz, tets = FIX[name + "_xyz"], FIX[name + "_tets"]
    phase = FIX[name + "_phase"] if two else None
    seq = orc.pgse(10600.0, 43100.0)
    q = seq.q_from_b(1000.0)
    ts = orc.time_grid(seq.T, 200.0)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        fem.set_diffusion(3e-3)
        fem.set_relaxation(1e-16)
        if two:
            fem.set_permeability(1e-5)
        fem.assemble()
        if two:
            assert fem.ndof == int(FIX["torus_ndof"])
        else:
            rp, ci = fem.pattern()
            rps, cis = orc.scalar_pattern(len(xyz), tets)
            assert np.array_equal(rp, rps) and np.array_equal(ci, cis)
        tight = fem.solve(200.0, 0.5, q * f, q * fp, gdir, rtol=1e-13, atol=1e-16)
        cli = fem.solve(200.0, 0.5, q * f, q * fp, gdir, rtol=1e-9, atol=1e-10)      # GCloudDmriSolver.py:219-222
    want = float(FIX[name + "_signal"])
    assert abs(tight["voi"] - float(FIX[name + "_voi"])) <= 1e-12 * tight["voi"]
    assert abs(tight["signal"] - want) <= 1e-8 * abs(want)
    assert abs(cli["signal"] - want) <= 1e-6 * abs(want)


def test_edge_cases():
    """Smallest and degenerate inputs: one tet; zero time steps; q = 0; an interface-free two-compartment mesh."""
    xyz = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    tets = np.array([[0, 1, 2, 3]], dtype=np.int32)
    ops = orc.assemble(xyz, tets, D=1e-3)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets)
        fem.set_diffusion(1e-3)
        fem.assemble()
        assert fem.ndof == 4 and fem.nnz == 16
        assert _relmax(fem.values("S"), ops.S.data) <= 1e-13
        res = fem.solve(10.0, 0.5, np.zeros(0), np.zeros(0), [1, 0, 0])           # no steps: u = IC
        assert res["n_steps"] == 0 and abs(res["signal"] - 1.0 / 6.0) <= 1e-15
        res = fem.solve(10.0, 0.5, np.full(5, 1e-3), np.full(5, 1e-3), [1, 0, 0], rtol=1e-13, atol=1e-18)
        seq_c = 1e-3
        u = np.ones(4, dtype=complex)
        Jg = ops.Jx
        for _ in range(5):
            A = (ops.M / 10.0 + 0.5 * ops.S + 0.5j * seq_c * Jg).toarray()
            b = (ops.M / 10.0 - 0.5 * ops.S - 0.5j * seq_c * Jg) @ u
            u = np.linalg.solve(A, b)
        assert abs(res["signal"] - float(ops.lumped @ u.real)) <= 1e-10 * abs(res["signal"])
    # two "compartments" that never touch: all cells phase 1 -> no interface facets, compartment 0 empty
    xyz2, tets2 = meshes.box_mesh((0,) * 3, (1,) * 3, 2, 2, 2)
    ph = np.ones(len(tets2), dtype=np.int32)
    ops2 = orc.assemble(xyz2, tets2, ph, D=1e-3, kappa=1e-5)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz2, tets2, ph)
        fem.set_diffusion(1e-3)
        fem.set_permeability(1e-5)
        fem.assemble()
        assert fem.n_iface == 0 and fem.ndof == ops2.ndof == len(xyz2)
        res = fem.solve(10.0, 0.5, np.zeros(3), np.zeros(3), [0, 0, 1])
        assert res["signal_comp"][0] == 0.0 and abs(res["signal_comp"][1] - 1.0) <= 1e-9      # default rtol 1e-9


def test_layered_sphere_parity_and_matrix_formalism():
    """Three-layer sphere (curved conforming interfaces, meshes.layered_sphere): pattern / values / signal against the
    oracle, and the matrix-formalism value of T2_Relaxation.ipynb cell 12 (D=3e-3, kappa=5e-5, delta=Delta=40000,
    b=1000 -> .7886) on the fine mesh, which the GPU solves in a blink (the CPU oracle needs 25 s per b-value)."""
    xyz, tets, marker = meshes.layered_sphere((5.0, 7.5, 10.0), (3, 2, 2), 2)
    ph = (marker % 2).astype(np.int32)
    D = np.array([3e-3, 1e-3, 3e-3])[marker]
    ops = orc.assemble(xyz, tets, ph, D=D, kappa=5e-5)
    seq = orc.pgse(2000.0, 6000.0)
    q, k = seq.q_from_b(1500.0), 200.0
    g = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    ref = orc.theta_solve(ops, seq, q, g, k, solver="lu")
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, ph)
        fem.set_diffusion(D)
        fem.set_permeability(5e-5)
        fem.assemble()
        rp, ci = fem.pattern()
        assert np.array_equal(rp, ops.rowptr) and np.array_equal(ci, ops.colidx)
        for name in ("M", "S", "Jx", "Jy", "Jz", "I"):
            assert _relmax(fem.values(name), getattr(ops, name).data) <= 1e-12, name
        res = fem.solve(k, 0.5, q * f, q * fp, g, rtol=1e-13, atol=1e-16)
    assert abs(res["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    xyz, tets, marker = meshes.layered_sphere((5.0, 7.5, 10.0), (8, 4, 4), 4)          # 41 k vertices
    seq = orc.pgse(40000.0, 40000.0)
    ts = orc.time_grid(seq.T, 200.0)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, (marker % 2).astype(np.int32))
        fem.set_diffusion(3e-3)
        fem.set_permeability(5e-5)
        fem.assemble()
        for b, want in ((1000.0, .7886), (3000.0, .4932)):
            q = seq.q_from_b(b)
            res = fem.solve(200.0, 0.5, q * f, q * fp, [0, 0, 1], rtol=1e-10, atol=1e-12)
            assert abs(res["signal"] / res["voi"] - want) <= 3e-3 * want


def test_persistent_batch_kernels(monkeypatch):
    """Batches of up to 16 members run the whole time loop as ONE cooperative launch.  Default: the many-warp kernel on
    the SELL copies (k_bicgstab_coop_batch, one operator copy per direction); BTFEM_BATCH_PERSIST=ring: the TMA-ring
    kernel (k_bicgstab_persistent_batch; members of one direction travel through the passes in groups of up to 4:
    groups of 4, 3, 2 and 1 units here).  Members drop out at different iterations; both agree with the one-at-a-time
    persistent solves and with the kernel chain to solver tolerance; a failing member ends the batch."""
    _, xyz, tets, phase, co = CASES[3]
    seq = orc.pgse(2000.0, 6000.0)
    k = 200.0
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    dirs = meshes.fibonacci_hemisphere(3)
    q = [seq.q_from_b(b) for b in (300.0, 1000.0, 2000.0, 3000.0, 4000.0)]
    members = [(qq * f, qq * fp, dirs[0]) for qq in q]              # 5 members of one direction: groups of 4 + 1
    members += [(qq * f, qq * fp, dirs[1]) for qq in q[:3]]         # 3 members
    members += [(qq * f, qq * fp, dirs[2]) for qq in q[:2]]         # 2 members
    par = dict(rtol=1e-10, atol=1e-14)
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        single = [fem.solve(k, 0.5, cA, cb, g, **par) for cA, cb, g in members]
        coop = fem.solve_batch(k, 0.5, members, **par)
        assert coop[0]["n_kernels"] <= 3                              # one launch for the whole time loop (+ signal)
        coop2 = fem.solve_batch(k, 0.5, members, **par)
        coop_split = fem.solve_batch(k, 0.5, members[:3], **par) + fem.solve_batch(k, 0.5, members[3:], **par)
        with pytest.raises(btfem.BTFemError):
            fem.solve_batch(k, 0.5, members[:6], rtol=1e-30, atol=0.0, maxit=3)
        coop_after = fem.solve_batch(k, 0.5, members[:6], **par)      # the handle is usable after the failure
        monkeypatch.setenv("BTFEM_BATCH_PERSIST", "0")
        chain = fem.solve_batch(k, 0.5, members, **par)
        assert chain[0]["n_kernels"] > 100
        monkeypatch.setenv("BTFEM_BATCH_PERSIST", "ring")
        ring = fem.solve_batch(k, 0.5, members, **par)
        assert ring[0]["n_kernels"] <= 3
        ring2 = fem.solve_batch(k, 0.5, members, **par)
        ring_split = fem.solve_batch(k, 0.5, members[:3], **par) + fem.solve_batch(k, 0.5, members[3:], **par)
        with pytest.raises(btfem.BTFemError):
            fem.solve_batch(k, 0.5, members[:6], rtol=1e-30, atol=0.0, maxit=3)
        ring_after = fem.solve_batch(k, 0.5, members[:6], **par)
        monkeypatch.setenv("BTFEM_BATCH_PERSIST", "hb")               # member-interleaved layout, cooperative kernel
        chb = fem.solve_batch(k, 0.5, members, **par)                 # 10 members = 2 groups of 8 (the second partly empty)
        assert chb[0]["n_kernels"] <= 3
        chb2 = fem.solve_batch(k, 0.5, members, **par)
        chb_split = fem.solve_batch(k, 0.5, members[:3], **par) + fem.solve_batch(k, 0.5, members[3:], **par)
        with pytest.raises(btfem.BTFemError):
            fem.solve_batch(k, 0.5, members[:6], rtol=1e-30, atol=0.0, maxit=3)
        chb_after = fem.solve_batch(k, 0.5, members[:6], **par)

    def close(x, y):
        return abs(x["signal"] - y["signal"]) <= 1e-9 * abs(y["signal"]) and \
            abs(x["total_iters"] - y["total_iters"]) <= max(3, 0.02 * y["total_iters"])

    for i, s1 in enumerate(single):
        assert coop[i]["signal"] == coop2[i]["signal"] and coop[i]["total_iters"] == coop2[i]["total_iters"]   # reproducible
        assert ring[i]["signal"] == ring2[i]["signal"] and ring[i]["total_iters"] == ring2[i]["total_iters"]
        assert ring[i]["signal"] == ring_split[i]["signal"]            # ring kernel: same bits whatever the batch
        assert close(coop[i], s1) and close(ring[i], s1) and close(chain[i], s1) and close(coop_split[i], s1)
        assert coop[i]["last_reason"] > 0 and ring[i]["last_reason"] > 0 and chb[i]["last_reason"] > 0
        assert chb[i]["signal"] == chb2[i]["signal"] and chb[i]["total_iters"] == chb2[i]["total_iters"]
        assert close(chb[i], s1) and close(chb_split[i], s1)
    for i in range(6):
        assert close(coop_after[i], single[i]) and ring_after[i]["signal"] == ring[i]["signal"]
        assert close(chb_after[i], single[i])
    assert len({round(s["signal"], 6) for s in coop}) == len(coop)
    assert len({s["total_iters"] for s in coop}) > 3                    # members leave the iteration at different times


@pytest.mark.parametrize("share", [4, 8])
def test_batch_on_an_sm_share(share):
    """A handle confined to a few SMs (btfem_set_sm_partition): a batch of 6 runs the cooperative batch kernel on 8
    blocks and falls back to the kernel chain on 4 (a block's chunk of the vector phases may span two members at most)."""
    _, xyz, tets, phase, co = CASES[3]
    seq = orc.pgse(2000.0, 6000.0)
    k = 200.0
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    dirs = meshes.fibonacci_hemisphere(3)
    qs = [seq.q_from_b(b) for b in (500.0, 3000.0)]
    members = [(q * f, q * fp, d) for d in dirs for q in qs]
    par = dict(rtol=1e-10, atol=1e-14)
    with btfem.BTFem(0) as fem:
        fem.set_sm_partition(share)
        _setup(fem, xyz, tets, phase, co)
        single = [fem.solve(k, 0.5, cA, cb, g, **par) for cA, cb, g in members]
        batch = fem.solve_batch(k, 0.5, members, **par)
    assert (batch[0]["n_kernels"] <= 3) == (share >= len(members))
    for s1, sb in zip(single, batch):
        assert abs(s1["signal"] - sb["signal"]) <= 1e-9 * abs(sb["signal"])
        assert abs(s1["total_iters"] - sb["total_iters"]) <= max(3, 0.02 * sb["total_iters"])


def test_interleaved_batch_layout(monkeypatch):
    """BTFEM_BATCH_LAYOUT=interleaved (k_hb_*: groups of 8 members, member-innermost vectors, one shared
    direction-independent operator): a member gets the same bits whatever batch it travels in -- also across a group
    boundary (9 members = 2 groups) -- and agrees with the default layout to solver tolerance."""
    _, xyz, tets, phase, co = CASES[3]
    seq = orc.pgse(2000.0, 6000.0)
    k = 200.0
    ts = orc.time_grid(seq.T, k)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    dirs = meshes.fibonacci_hemisphere(3)
    qs = [seq.q_from_b(b) for b in (500.0, 1500.0, 3000.0)]
    members = [(q * f, q * fp, d) for d in dirs for q in qs]           # 9 members
    with btfem.BTFem(0) as fem:
        _setup(fem, xyz, tets, phase, co)
        default = fem.solve_batch(k, 0.5, members, rtol=1e-10, atol=1e-14)
        monkeypatch.setenv("BTFEM_BATCH_LAYOUT", "interleaved")
        inter = fem.solve_batch(k, 0.5, members, rtol=1e-10, atol=1e-14)
        split = fem.solve_batch(k, 0.5, members[:2], rtol=1e-10, atol=1e-14) + \
            fem.solve_batch(k, 0.5, members[2:], rtol=1e-10, atol=1e-14)
    for a, b in zip(inter, split):
        assert a["signal"] == b["signal"] and a["total_iters"] == b["total_iters"]      # bit-identical
    for a, d in zip(inter, default):
        assert abs(a["signal"] - d["signal"]) <= 1e-9 * abs(d["signal"])
        assert abs(a["total_iters"] - d["total_iters"]) <= max(3, 0.02 * d["total_iters"])


def test_sm_partition_concurrent_handles():
    """btfem_set_sm_partition + sweep.run_sweep_concurrent: handles confined to disjoint SM shares solve concurrently
    (one host thread and one persistent-kernel launch each); same bits whichever handle takes a unit, solver-tolerance
    agreement with a full-device solve."""
    import sympy as sp
    from dmri_fem_cloud_b200 import dmrifemlib as dl, sweep
    _, xyz, tets, phase, co = CASES[3]

    def make_fem(fem):
        _setup_inputs(fem, xyz, tets, phase, co)

    mp = dl.MRI_parameters()
    mp.delta, mp.Delta = 2000.0, 6000.0
    mp.T = mp.delta + mp.Delta
    mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
    mp.bvalue = 1000.0
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 200.0
    dirs = meshes.fibonacci_hemisphere(3)
    bvals = [500.0, 3000.0]
    par = dict(rtol=1e-10, atol=1e-14, maxit=100000)
    fems = sweep.make_concurrent_handles(make_fem, 3, n_sms=12)          # 3 handles x 4 SMs
    try:
        assert all(f.spmv_kernel == 2 for f in fems)
        mine, a = sweep.run_sweep_concurrent(fems, mp, sim, dirs, bvals, par)
        _, b = sweep.run_sweep_concurrent(fems[::-1], mp, sim, dirs, bvals, par)
    finally:
        for f in fems:
            f.close()
    assert np.array_equal(a, b)                                          # bit-identical, whoever solved what
    with btfem.BTFem(0) as fem:
        make_fem(fem)
        fem.assemble()
        _, ref = sweep.run_sweep(fem, mp, sim, dirs, bvals, par)
    assert mine == list(range(6)) and np.max(np.abs(a - ref) / np.abs(ref)) <= 1e-9
