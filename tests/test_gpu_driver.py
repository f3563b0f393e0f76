"""GPU tests of the reference-facing layer: the DmriFemLib mirror and the GCloudDmriSolver CLI."""
import os

import numpy as np
import pytest
import sympy as sp

import bt_oracle as orc
from conftest import REF_MESH_DIR
from dmri_fem_cloud_b200 import cli, dmrifemlib as dl, meshes

pytestmark = pytest.mark.gpu


def test_mesh_stats_match_numpy():
    rng = np.random.default_rng(0)
    xyz, tets = meshes.cylinder(3.0, 10.0, nr=3, nsec=10, nz=6)
    xyz = xyz + 0.02 * rng.standard_normal(xyz.shape)
    m = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.qvalue = 0.0
    md = dl.MyDomain(m, mp)
    assert md.hmin == m.hmin() and md.hmax == m.hmax()          # same fp64 operations, exact


def _pgse(mp, delta, Delta):
    mp.delta, mp.Delta = delta, Delta
    mp.T = Delta + delta
    mp.fs_sym = sp.Piecewise((1., mp.s < delta), (0., mp.s < Delta), (-1., mp.s < mp.T), (0., True))


def test_driver_two_compartment_matches_oracle(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 5.0, (3, 2, 2), 12, 2)
    phase = (marker % 2).astype(np.int32)
    mesh = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.bvalue = 1000
    _pgse(mp, 2000.0, 6000.0)
    mp.set_gradient_dir(mesh, 0, 1, 0)
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 200
    md = dl.MyDomain(mesh, mp)
    md.phase, md.IsDomainMultiple, md.kappa = phase, True, 1e-5
    md.Apply()
    md.D0 = 3e-3
    md.D = md.D0
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    ls.parameters["relative_tolerance"] = 1e-12
    ls.parameters["absolute_tolerance"] = 1e-15
    sim.solve(md, mp, ls)
    text = dl.PostProcessing(md, mp, sim, None, '')
    ops = orc.assemble(xyz, tets, phase, D=3e-3, invT2=1e-16, kappa=1e-5)
    seq = orc.pgse(2000.0, 6000.0)
    ref = orc.theta_solve(ops, seq, seq.q_from_b(1000.0), [0, 1, 0], 200.0, solver="lu")
    assert abs(mp.qvalue - seq.q_from_b(1000.0)) <= 1e-15 * mp.qvalue
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert "Normalized signal: %.6e" % (ref["signal"] / ref["voi"]) in text
    assert os.path.exists("log.txt")
    # solution in the reference's blocked layout
    want = orc.expand_to_reference_layout(ops, ref["u"]).reshape(4, -1)
    assert np.max(np.abs(sim.u_0 - want)) <= 1e-8 * np.max(np.abs(want))


def test_cli_config0_npz(tmp_path, monkeypatch, capsys):
    """BASELINE configs[0]: -M 0 -b 1000 -d 10600 -D 43100 -k 200 -K 3e-3 -gdir 1 0 0 on a small cylinder."""
    monkeypatch.chdir(tmp_path)
    xyz, tets = meshes.cylinder(3.0, 25.0, nr=2, nsec=8, nz=8)
    np.savez("cyl.npz", xyz=xyz, tets=tets)
    text = cli.main(["prog", "-f", "cyl.npz", "-M", "0", "-b", "1000", "-d", "10600", "-D", "43100", "-k", "200",
                     "-K", "3e-3", "-gdir", "1", "0", "0"])
    ops = orc.assemble(xyz, tets, D=3e-3, invT2=1e-16)
    seq = orc.pgse(10600.0, 43100.0)
    ref = orc.theta_solve(ops, seq, seq.q_from_b(1000.0), [1, 0, 0], 200.0, solver="lu")
    assert ref["n_steps"] == 270
    got = float(text.split("Normalized signal: ")[1].split(",")[0])
    assert abs(got - ref["signal"] / ref["voi"]) <= 2e-6 * got      # CLI tolerances (rtol 1e-9) + %.6e print
    assert "b: 1000.000, g: 0.056, q: 1.500e-05" in text            # ExplicitImplementation.ipynb cell 10 prints q=1.499786e-05


def test_cli_reference_msh_fixture(tmp_path, monkeypatch):
    """-f <gmsh v2 file>: the reference's cyl6 fixture.  On the GPU box (no /root/reference) the same mesh is
    written as `.msh` from the committed golden arrays (tests/golden/fixture_meshes.npz), so the test never skips."""
    monkeypatch.chdir(tmp_path)
    path = os.path.join(REF_MESH_DIR, "cyl6_r_3E_6_vol.msh.zip")
    if not os.path.exists(path):
        from conftest import GOLDEN
        z = np.load(os.path.join(GOLDEN, "fixture_meshes.npz"))
        path = str(tmp_path / "cyl6_r_3E_6_vol.msh")
        meshes.write_gmsh2(path, z["cyl6_r_3E_6_vol_xyz"], z["cyl6_r_3E_6_vol_tets"])
    text = cli.main(["prog", "-f", path, "-M", "0", "-b", "1000", "-k", "200", "-K", "3e-3", "-gdir", "1", "0", "0"])
    xyz, tets, _ = meshes.read_gmsh2(path)
    assert len(xyz) == 54 and len(tets) == 123
    ops = orc.assemble(xyz, tets, D=3e-3, invT2=1e-16)
    seq = orc.pgse(10600.0, 43100.0)
    ref = orc.theta_solve(ops, seq, seq.q_from_b(1000.0), [1, 0, 0], 200.0, solver="lu")
    got = float(text.split("Normalized signal: ")[1].split(",")[0])
    assert abs(got - ref["signal"] / ref["voi"]) <= 2e-6 * got


def test_cli_h5_input_two_compartments(tmp_path, monkeypatch):
    """The documented command line (README.md:89-94: `-f files.h5 -M 1 -b 1000 -p 1e-5 ...`) on a DOLFIN HDF5
    container as PreprocessingMultiCompt.py:148-152 writes it (mesh, T2 = 1e6, ic, phase, d00..d22): tensor and phase
    come from the file, T2 is read and -- like the reference -- not applied."""
    from dmri_fem_cloud_b200 import hdf5io, preprocess
    monkeypatch.chdir(tmp_path)
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 2.0, (2, 1, 1), 10, 1)
    fields = preprocess.cell_fields(marker, [3e-3, 1e-3, 3e-3], [1e6] * 3, [1.0] * 3)
    hdf5io.write_dolfin_h5("files.h5", xyz, tets, dict(phase=(marker % 2).astype(float), **fields))
    text = cli.main(["prog", "-f", "files.h5", "-M", "1", "-b", "1000", "-p", "1e-5", "-d", "2000", "-D", "6000",
                     "-k", "200", "-gdir", "0", "1", "0"])
    ops = orc.assemble(xyz, tets, (marker % 2).astype(np.int32), D=np.array([3e-3, 1e-3, 3e-3])[marker], invT2=1e-16,
                       kappa=1e-5)
    seq = orc.pgse(2000.0, 6000.0)
    ref = orc.theta_solve(ops, seq, seq.q_from_b(1000.0), [0, 1, 0], 200.0, solver="lu")
    got = float(text.split("Normalized signal: ")[1].split(",")[0])
    assert abs(got - ref["signal"] / ref["voi"]) <= 2e-6 * got


def test_config3_layered_variable_kappa_T2_ogse(tmp_path, monkeypatch):
    """BASELINE configs[2]: multilayered cylinder/disk, variable permeability (kappa_tensor by marker pair,
    MultilayeredDiskVariablePermeability.ipynb cell 10), per-layer D through ImposeDiffusionTensor, per-layer
    T2 (T2_Relaxation.ipynb cell 10), cos-OGSE profile (ArbitraryTimeSequence.ipynb cell 10, profile 3)."""
    monkeypatch.chdir(tmp_path)
    xyz, tets, marker = meshes.layered_cylinder((5.0, 7.5, 10.0), 5.0, (3, 2, 2), 12, 2)
    phase = (marker % 2).astype(np.int32)
    nc = len(tets)
    mesh = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.bvalue = 1000
    mp.delta, mp.Delta = 4000.0, 4000.0
    t0, seq_ext = 100.0, 100.0
    Dd = mp.Delta + mp.delta
    mp.T = Dd + t0 + seq_ext
    mp.nperiod = 1
    omega = 2.0 * mp.nperiod * np.pi / mp.delta      # dolfin `pi` is a float (ArbitraryTimeSequence.ipynb)
    tau = Dd / 2.0
    mp.fs_sym = sp.Piecewise((0., mp.s < t0), (sp.cos(omega * (mp.s - t0)), mp.s <= mp.delta + t0),
                             (0., mp.s <= tau + t0), (-sp.cos(omega * (mp.s - t0 - tau)), mp.s <= mp.delta + tau + t0),
                             (0., True))
    mp.set_gradient_dir(mesh, 1, 1, 0)
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 100
    md = dl.MyDomain(mesh, mp)
    md.phase, md.IsDomainMultiple = phase, True
    kt = np.zeros((3, 3))
    kt[0, 1] = kt[1, 0] = 1e-4
    kt[1, 2] = kt[2, 1] = 1e-5
    md.kappa, md.kappa_marker = kt, marker
    md.T2_cell = np.array([4e16, 4e4, 4e4])[marker]
    md.Apply()
    Dl = np.array([3e-3, 1e-3, 3e-3])[marker]
    z = np.zeros(nc)
    md.ImposeDiffusionTensor(Dl, z, z, z, Dl, z, z, z, Dl)
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    ls.parameters["relative_tolerance"] = 1e-12
    ls.parameters["absolute_tolerance"] = 1e-15
    sim.solve(md, mp, ls)
    text = dl.PostProcessing(md, mp, sim, None, '')
    assert "kappa:" not in text                                  # non-scalar kappa: the line without kappa (:935)
    ops = orc.assemble(xyz, tets, phase, D=Dl, invT2=1.0 / md.T2_cell,
                       kappa_facet=lambda fv, c0, c1: kt[marker[c0], marker[c1]])
    seq = orc.Sequence(mp.fs_sym, mp.T, mp.s)
    assert abs(seq.q_from_b(1000.0) - mp.qvalue) <= 1e-14 * mp.qvalue
    ref = orc.theta_solve(ops, seq, mp.qvalue, [1, 1, 0], 100.0, solver="lu")
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-8 * abs(ref["signal"])
    assert abs(sim.stats["signal_comp"][1] - float(ops.lumped[ops.dof_comp == 1] @ ref["u"].real[ops.dof_comp == 1])) \
        <= 1e-8 * abs(ref["signal"])


def test_lu_solver_stand_in(tmp_path, monkeypatch):
    """`linsolver = PETScLUSolver("mumps")` (ConvergenceTest.ipynb / T2_Relaxation.ipynb cell 10): the stand-in
    (BiCGStab to rounding level) reproduces the exact discrete solve of the oracle (sparse LU) to 1e-10."""
    monkeypatch.chdir(tmp_path)
    xyz, tets = meshes.box_mesh((-2.5,) * 3, (2.5,) * 3, 6, 6, 6)
    mesh = dl.Mesh(xyz, tets)
    mp = dl.MRI_parameters()
    mp.bvalue = 1000
    _pgse(mp, 1000.0, 3000.0)
    mp.set_gradient_dir(mesh, 0, 0, 1)
    mp.Apply()
    sim = dl.MRI_simulation()
    sim.k = 100
    md = dl.MyDomain(mesh, mp)
    md.Apply()
    md.D0 = 2e-3
    md.D = md.D0
    sim.solve(md, mp, dl.PETScLUSolver("mumps"))
    dl.PostProcessing(md, mp, sim, None, '')
    ops = orc.assemble(xyz, tets, D=2e-3, invT2=1e-16)
    seq = orc.pgse(1000.0, 3000.0)
    ref = orc.theta_solve(ops, seq, seq.q_from_b(1000.0), [0, 0, 1], 100.0, solver="lu")
    assert abs(sim.stats["signal"] - ref["signal"]) <= 1e-10 * abs(ref["signal"])
    assert dl.LUSolver is dl.PETScLUSolver


def test_comri_drivers_match_oracle(tmp_path, monkeypatch):
    """The comri mains on libbtfem (comri.py): f(t_n) on both sides, FT closed at Delta+delta, s/s0 -- against the
    oracle in the same variant (rhs_uses_current_f), on the reference's own cylinder / torus meshes."""
    import io
    from conftest import GOLDEN
    from dmri_fem_cloud_b200 import comri
    fix = np.load(os.path.join(GOLDEN, "fixture_meshes.npz"))
    monkeypatch.setattr(comri, "KRYLOV", {"rtol": 1e-11, "atol": 1e-16, "maxit": 100000})

    class Seq:                                   # FT of the C++ drivers as an oracle sequence
        def __init__(self, delta, Delta):
            self.delta, self.Delta, self.T = delta, Delta, delta + Delta

        def f(self, t):
            return comri.FT(t, self.delta, self.Delta)

        def F(self, t):
            return 0.0

    # one-comp: -m cyl12 -b 1000 -d 2000 -D 6000 -k 200 -v 1 0 0 -K 3e-3
    xyz, tets = fix["cyl12_r_3E_6_vol_xyz"], fix["cyl12_r_3E_6_vol_tets"]
    p = comri.parse("one-comp", ["demo", "-b", "1000", "-d", "2000", "-D", "6000", "-k", "200", "-v", "1", "0", "0",
                                 "-K", "3e-3"])
    p["mesh"] = (xyz, tets)
    buf = io.StringIO()
    r = comri.run("one-comp", p, out=buf)
    ops = orc.assemble(xyz, tets, D=3e-3)
    ref = orc.theta_solve(ops, Seq(2000.0, 6000.0), r["gnorm"], [1, 0, 0], 200.0, solver="lu", rhs_uses_current_f=True)
    assert len(r["ts"]) == ref["n_steps"] == 41
    assert abs(r["gnorm"] - np.sqrt(1000.0) / np.sqrt(2000.0 ** 2 * (6000.0 - 2000.0 / 3))) <= 1e-18
    assert abs(r["s"] - ref["signal"] / ref["voi"]) <= 1e-8 * r["s"]
    assert ("b: %f, gnorm: %f, q: %f, gdir: (%f, %f, %f), s: %f" % (1000.0, r["gnorm"], r["qvalue"], 1, 0, 0, r["s"])) \
        in buf.getvalue()
    # lagging the right-hand side (DmriFemLib) gives a different number: the variant matters
    lag = orc.theta_solve(ops, Seq(2000.0, 6000.0), r["gnorm"], [1, 0, 0], 200.0, solver="lu")
    assert abs(lag["signal"] - ref["signal"]) > 1e-6 * abs(ref["signal"])

    # two-comp with the compartment given as a sub-mesh (-c): multi_layer_torus + its compt1 cells
    xyz, tets, phase = fix["torus_xyz"], fix["torus_tets"], fix["torus_phase"]
    sub_cells = tets[phase == 1]
    used = np.unique(sub_cells)
    remap = -np.ones(len(xyz), dtype=np.int64)
    remap[used] = np.arange(len(used))
    p = comri.parse("two-comp", ["demo", "-b", "1000", "-d", "2000", "-D", "6000", "-N", "20", "-v", "0", "1", "0",
                                 "-p", "1e-5"])
    p["mesh"], p["cell"] = (xyz, tets), (xyz[used], remap[sub_cells].astype(np.int32))
    r = comri.run("two-comp", p, out=io.StringIO())
    assert np.array_equal(r["phase"], phase) and r["dt"] == 400.0
    ops = orc.assemble(xyz, tets, phase, D=3e-3, kappa=1e-5)
    ref = orc.theta_solve(ops, Seq(2000.0, 6000.0), r["gnorm"], [0, 1, 0], 400.0, solver="lu", rhs_uses_current_f=True)
    assert abs(r["s"] - ref["signal"] / ref["voi"]) <= 1e-8 * r["s"]

    # multilayer: torus shells by formula, loop t < T, no preconditioner
    p = comri.parse("multilayer", ["demo", "-b", "500", "-d", "1000", "-D", "3000", "-N", "16"])
    p["mesh"] = (xyz, tets)
    r = comri.run("multilayer", p, out=io.StringIO())
    assert len(r["ts"]) == 16 and set(np.unique(r["phase"])) == {0, 1}
    ops = orc.assemble(xyz, tets, r["phase"], D=3e-3, kappa=5e-5)
    ref = orc.theta_solve(ops, Seq(1000.0, 3000.0), r["gnorm"], [0, 0, 1], 250.0, solver="lu", closed=False,
                          rhs_uses_current_f=True)
    assert ref["n_steps"] == 16
    assert abs(r["s"] - ref["signal"] / ref["voi"]) <= 1e-8 * r["s"]
