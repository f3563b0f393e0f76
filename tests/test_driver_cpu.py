"""CPU: the host layers (GCloudDmriSolver-compatible CLI, DmriFemLib mirror, comri demos, sweeps) driven end to end
with the GPU handle replaced by tests/fake_btfem.py (an oracle-backed TEST DOUBLE).  What is checked is host logic:
flags, sequence scalars, the call order into the handle, printed result lines.  CUDA parity is `-m gpu`."""
import io

import numpy as np
import pytest
import sympy as sp

import bt_oracle as orc
from fake_btfem import FakeBTFem
from dmri_fem_cloud_b200 import btfem, cli, comri, dmrifemlib as dl, meshes, sweep


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(btfem, "BTFem", FakeBTFem)
    return FakeBTFem


def test_cli_config0_flow(tmp_path, monkeypatch, fake):
    """BASELINE configs[0] through cli.main: 270 steps, q printed like the reference, signal of the exact stepping."""
    monkeypatch.chdir(tmp_path)
    xyz, tets = meshes.cylinder(3.0, 25.0, nr=2, nsec=8, nz=6)
    np.savez("cyl.npz", xyz=xyz, tets=tets)
    text = cli.main(["prog", "-f", "cyl.npz", "-M", "0", "-b", "1000", "-d", "10600", "-D", "43100", "-k", "200",
                     "-K", "3e-3", "-gdir", "2", "0", "0", "-N", "7", "-q", "5"])
    ops = orc.assemble(xyz, tets, D=3e-3, invT2=1e-16)
    seq = orc.pgse(10600.0, 43100.0)
    ref = orc.theta_solve(ops, seq, seq.q_from_b(1000.0), [1, 0, 0], 200.0, solver="lu")
    got = float(text.split("Normalized signal: ")[1].split(",")[0])
    assert abs(got - ref["signal"] / ref["voi"]) <= 1e-6 * got           # %.6e print
    assert "b: 1000.000, g: 0.056, q: 1.500e-05" in text


def test_cli_two_compartments_from_marker_and_tensor_file(tmp_path, monkeypatch, fake):
    """-M 1 with phase = marker % 2, per-cell T2 and the diffusion tensor from the input file (is_kcoeff_from_file)."""
    monkeypatch.chdir(tmp_path)
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 2.0, (2, 1, 1), 10, 1)
    nc = len(tets)
    Dl = np.array([3e-3, 1e-3, 3e-3])[marker]
    z = np.zeros(nc)
    T2 = np.array([1e6, 4e4, 1e6])[marker]
    np.savez("in.npz", xyz=xyz, tets=tets, phase=(marker % 2), T2=T2, d00=Dl, d01=z, d02=z, d10=z, d11=Dl, d12=z,
             d20=z, d21=z, d22=Dl)
    flags = ["prog", "-M", "1", "-b", "2000", "-p", "5e-5", "-d", "2000", "-D", "6000", "-k", "200", "-gdir", "0", "1", "0"]
    seq = orc.pgse(2000.0, 6000.0)
    # the reference reads the T2 dataset but never applies it (GCloudDmriSolver.py:167-171 vs DmriFemLib.py:807):
    # identical flags -> T2 = 1e16; `-applyT2 1` (extension) applies the per-cell values.  Same from `.npz` and `.h5`.
    from dmri_fem_cloud_b200 import hdf5io
    fields = dict(phase=(marker % 2), T2=T2, ic=np.ones(nc), d00=Dl, d01=z, d02=z, d10=z, d11=Dl, d12=z, d20=z, d21=z, d22=Dl)
    hdf5io.write_dolfin_h5("in.h5", xyz, tets, fields)
    for extra, invT2 in (([], 1e-16), (["-applyT2", "1"], 1.0 / T2)):
        ops = orc.assemble(xyz, tets, (marker % 2).astype(np.int32), D=Dl, invT2=invT2, kappa=5e-5)
        ref = orc.theta_solve(ops, seq, seq.q_from_b(2000.0), [0, 1, 0], 200.0, solver="lu")
        for infile in ("in.npz", "in.h5"):
            text = cli.main(flags + ["-f", infile] + extra)
            got = float(text.split("Normalized signal: ")[1].split(",")[0])
            assert abs(got - ref["signal"] / ref["voi"]) <= 1e-6 * got, (infile, extra)
            assert "kappa: 5.000e-05" in text or "kappa:" in text


def test_driver_weak_periodic_and_ogse(tmp_path, monkeypatch, fake):
    """MyDomain with PeriodicDir = [1,0,0] (weakly imposed) and a sine-OGSE profile: F(t_{n-1}) reaches the handle."""
    monkeypatch.chdir(tmp_path)
    xyz, tets = meshes.box_mesh((-3, -1, -1), (3, 1, 1), 6, 2, 2)
    mesh = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.bvalue = 500
    mp.delta, mp.Delta = 1000.0, 2000.0
    mp.T = mp.delta + mp.Delta
    om = 2 * np.pi / mp.delta
    mp.fs_sym = sp.Piecewise((sp.sin(om * mp.s), mp.s < mp.delta), (0., mp.s < mp.Delta),
                             (-sp.sin(om * (mp.s - mp.Delta)), mp.s < mp.T), (0., True))
    mp.set_gradient_dir(mesh, 1, 0, 0)
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 100
    sim.verbose = False
    md = dl.MyDomain(mesh, mp)
    md.PeriodicDir = [1, 0, 0]
    md.Apply()
    md.D0 = 2e-3
    md.D = md.D0
    sim.solve(md, mp, dl.KrylovSolver("bicgstab", "petsc_amg"))
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tets)
    assert (md.hmin, md.kappa_e_scalar) == (hmin, 3e-3 / hmin)
    ops = orc.assemble(xyz, tets, D=2e-3, invT2=1e-16, bnd_kappa_vertex=orc.periodic_marker(xyz, [1, 0, 0], lo, hi, hmin))
    seq = orc.Sequence(mp.fs_sym, mp.T, mp.s)
    per = orc.periodic_term(xyz, tets, ops, [1, 0, 0], lo, hi, mp.qvalue, [1, 0, 0], 0.5)
    ref = orc.theta_solve(ops, seq, mp.qvalue, [1, 0, 0], 100.0, solver="lu", periodic=per)
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-10 * abs(ref["signal"])
    assert ("solve", "bicgstab", "jacobi") in sim.fem.calls


def test_comri_flows(fake, monkeypatch):
    monkeypatch.setattr(comri._bt, "BTFem", FakeBTFem)
    xyz, tets = meshes.cylinder(3.0, 10.0, nr=2, nsec=8, nz=4)
    p = comri.parse("one-comp", ["demo", "-b", "1000", "-d", "2000", "-D", "6000", "-N", "40", "-v", "0", "0", "3"])
    p["mesh"] = (xyz, tets)
    buf = io.StringIO()
    r = comri.run("one-comp", p, out=buf)
    assert r["dt"] == 200.0 and len(r["ts"]) == 41 and "gdir: (0.000000, 0.000000, 1.000000), s: %f" % r["s"] in buf.getvalue()
    with pytest.raises(RuntimeError):
        comri.run("one-comp", dict(p, nrefine=1), out=io.StringIO())
    q = comri.parse("one-comp", ["demo", "-q", "0.5", "-d", "2000", "-D", "6000"])
    q["mesh"] = (xyz, tets)
    r2 = comri.run("one-comp", q, out=io.StringIO())
    assert abs(r2["gnorm"] - 0.5 * comri.G_RATIO * 1e-12) < 1e-20 and r2["bvalue"] > 0


def test_sweep_batches_and_sharding(fake):
    """run_sweep: every unit solved once, q per b-value from convert_b2q, members grouped in batches."""
    xyz, tets = meshes.box_mesh((-2,) * 3, (2,) * 3, 3, 3, 3)
    fem = FakeBTFem()
    fem.set_mesh(xyz, tets)
    fem.set_diffusion(2e-3)
    fem.assemble()
    mp = dl.MRI_parameters()
    mp.delta, mp.Delta = 1000.0, 3000.0
    mp.T = mp.delta + mp.Delta
    mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
    mp.bvalue = 100.0
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 200.0
    dirs = meshes.fibonacci_hemisphere(4)
    bvals = [500.0, 2000.0]
    par = dict(rtol=1e-9, atol=1e-10, maxit=1000)
    full = np.zeros(8)
    for rank in range(2):
        mine, sig = sweep.run_sweep(fem, mp, sim, dirs, bvals, par, rank=rank, world=2, batch=3)
        full[mine] = sig
    assert [c for c in fem.calls if c[0] == "solve_batch"] == [("solve_batch", 3), ("solve_batch", 1)] * 2
    one, s1 = sweep.run_sweep(fem, mp, sim, dirs, bvals, par, batch=1)
    assert one == list(range(8)) and np.allclose(full, s1, rtol=1e-12)
    seq = orc.pgse(1000.0, 3000.0)
    ref = orc.theta_solve(fem.ops, seq, seq.q_from_b(2000.0), dirs[1], 200.0, solver="lu")
    assert abs(full[1 * 2 + 1] - ref["signal"] / ref["voi"]) <= 1e-10
    assert full[1] < full[0] < 1.0                      # higher b, lower signal


def test_driver_strong_periodic_flow(tmp_path, monkeypatch, fake):
    """IsDomainPeriodic = True with PeriodicDir = [1,1,0]: the driver identifies the vertices of opposite faces
    (periodic.vertex_map == the oracle's), hands the INTEGRATED profile to the solver, and the result equals the
    oracle's transformed-equation stepping."""
    from dmri_fem_cloud_b200 import periodic
    monkeypatch.chdir(tmp_path)
    xyz, tets, ph = meshes.box_with_sphere(4.0, 5, 2.5)
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tets)
    vm = periodic.vertex_map(xyz, [1, 1, 0], lo, hi, 1e-2 * hmin)
    assert np.array_equal(vm, orc.periodic_vertex_map(xyz, [1, 1, 0], lo, hi, 1e-2 * hmin))
    assert np.array_equal(periodic.vertex_map(xyz, [1, 1, 1], lo, hi, 1e-2 * hmin),
                          orc.periodic_vertex_map(xyz, [1, 1, 1], lo, hi, 1e-2 * hmin))
    with pytest.raises(RuntimeError):
        bad = xyz.copy()
        bad[np.argmax(bad[:, 0] + 1e-3 * bad[:, 1])] += [0.0, 0.3, 0.0]
        periodic.vertex_map(bad, [1, 0, 0], lo, hi, 1e-2 * hmin)
    mesh = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.bvalue = 1000
    mp.delta, mp.Delta = 1000.0, 3000.0
    mp.T = mp.delta + mp.Delta
    mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
    mp.set_gradient_dir(mesh, 1, 0.5, 0)
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 100
    sim.verbose = False
    md = dl.MyDomain(mesh, mp)
    md.phase, md.IsDomainMultiple, md.kappa = ph, True, 5e-5
    md.PeriodicDir, md.IsDomainPeriodic = [1, 1, 0], True
    md.Apply()
    md.D0 = 2e-3
    md.D = md.D0
    sim.solve(md, mp, dl.KrylovSolver("bicgstab", "jacobi"))
    ops = orc.assemble(xyz, tets, ph, D=2e-3, invT2=1e-16, kappa=5e-5, vmaster=vm)
    seq = orc.pgse(1000.0, 3000.0)
    ref = orc.theta_solve_strong(ops, seq, mp.qvalue, [1, 0.5, 0], 100.0)
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-10 * abs(ref["signal"])
    assert sim.fem.periodic is None and sim.fem.vmaster is not None      # no weak marker in this mode


def test_bench_hardi_sweep_helper(fake):
    """bench.py's HARDI figure: the helper shards 4 directions x 4 b-values over two ranks and every unit is solved
    once (tiny mesh, test double)."""
    import importlib.util
    import os
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    tot = 0.0
    for rank in range(2):
        dt, n_sig, n_vert, checksum = b.hardi_sweep(0, rank, 2, batch=3, h=4.0, ndir=4)
        assert dt > 0 and n_sig == 16 and n_vert > 50 and 0 < checksum < 8
        tot += checksum
    one = b.hardi_sweep(0, 0, 1, batch=16, h=4.0, ndir=4)
    assert abs(one[3] - tot) <= 1e-9 * tot
