"""CPU: host-side logic of the product (no GPU): C-ABI library loads and exports every symbol the
header declares; the DmriFemLib mirror's scalar logic; mesh readers/generators; sweep sharding."""
import os
import re

import numpy as np
import pytest
import sympy as sp

import bt_oracle as orc
from conftest import REF_MESH_DIR, ROOT
from dmri_fem_cloud_b200 import btfem, dmrifemlib as dl, meshes, sweep


def test_cabi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "btfem.h")).read()
    declared = set(re.findall(r"\b(btfem_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(built_lib, name), name
    assert set(built_lib._btfem_symbols) == declared       # the binding covers the whole header
    assert built_lib.btfem_version() >= 100


def test_no_cpu_fallback_without_gpu(built_lib):
    import ctypes
    h = ctypes.c_void_p()
    rc = built_lib.btfem_create(0, ctypes.byref(h))
    if rc == 0:                                            # a GPU is present: fine, just clean up
        built_lib.btfem_destroy(h)
    else:
        assert rc == -2                                    # BTFEM_ECUDA, and the Python layer raises
        with pytest.raises(btfem.BTFemError):
            btfem.BTFem(0)


def test_struct_layouts_have_no_padding():
    import ctypes
    assert ctypes.sizeof(btfem.SolveArgs) == 8 * 17
    assert ctypes.sizeof(btfem.SolveOut) == 8 * 14


def test_mri_parameters_match_reference_numbers():
    mp = dl.MRI_parameters()
    mp.bvalue = 1000
    mp.delta, mp.Delta = 10600, 43100
    mp.T = mp.delta + mp.Delta
    mp.fs_sym = sp.Piecewise((1., mp.s < mp.delta), (0., mp.s < mp.Delta), (-1., mp.s < mp.T), (0., True))
    mp.set_gradient_dir(None, 0, 2, 0)
    mp.Apply()
    assert "%.6e" % mp.qvalue == "1.499786e-05"            # ExplicitImplementation.ipynb cell 10
    assert "%.3f" % mp.gvalue == "0.056"
    assert np.allclose(mp.gdir.array(), [0, 1, 0])         # normalised
    sim = dl.MRI_simulation()
    sim.k = 200
    ts = sim.time_grid(mp)
    assert len(ts) == 270
    f, F = mp.profiles_on_grid(ts)
    for i in (0, 52, 53, 54, 215, 216, 268, 269):
        assert f[i] == mp.time_profile(ts[i]) and abs(F[i] - mp.itime_profile(ts[i])) <= 1e-9 * max(1, abs(F[i]))
    seq = orc.pgse(10600.0, 43100.0)
    assert np.array_equal(ts, orc.time_grid(seq.T, 200.0))
    # g <-> q round trip and the gvalue entry point
    mp2 = dl.MRI_parameters()
    mp2.gvalue = mp.gvalue
    mp2.T, mp2.fs_sym = mp.T, mp.fs_sym
    mp2.Apply()
    assert abs(mp2.bvalue - 1000) <= 1e-9 * 1000


def test_zero_gradient_direction_exits():
    mp = dl.MRI_parameters()
    with pytest.raises(SystemExit):
        mp.set_gradient_dir(None, 0, 0, 0)                 # DmriFemLib.py:816,823


def test_krylov_solver_defaults():
    ls = dl.KrylovSolver("bicgstab", "jacobi")
    assert ls.parameters["relative_tolerance"] == 1e-6 and ls.parameters["nonzero_initial_guess"] is False
    with pytest.raises(RuntimeError):
        dl.KrylovSolver("cg", "jacobi")
    with pytest.raises(RuntimeError):
        dl.KrylovSolver("bicgstab", "no_such_pc")
    # PETSc's default preconditioner is ILU(0) on one process; "ilu" is implemented (comri fenics-cpp main.cpp:180-183)
    assert dl.KrylovSolver("bicgstab").preconditioner == "ilu" and dl.KrylovSolver("bicgstab").requested_preconditioner == "default"
    assert dl.KrylovSolver("gmres", "ilu").preconditioner == "ilu"
    # multigrid & co. (ECS_226Cylinders.ipynb / RealNeurons.ipynb cell 10) run with Jacobi -- loudly
    with pytest.warns(RuntimeWarning, match="NOT available"):
        ls = dl.KrylovSolver("bicgstab", "petsc_amg")
    assert ls.preconditioner == "jacobi" and ls.requested_preconditioner == "petsc_amg"
    assert dl.KrylovSolver("gmres", "none").preconditioner == "none"
    lu = dl.PETScLUSolver("mumps")
    assert (lu.method, lu.preconditioner, lu.parameters["relative_tolerance"]) == ("bicgstab", "jacobi", 1e-13)


@pytest.mark.skipif(not os.path.isdir(REF_MESH_DIR), reason="reference meshes not present")
def test_readers_on_reference_fixtures():
    sizes = {"cyl6_r_3E_6_vol.msh.zip": (54, 123), "cyl12_r_3E_6_vol.msh.zip": (178, 567)}
    for name, (nv, nc) in sizes.items():
        xyz, tets, marker = meshes.read_gmsh2(os.path.join(REF_MESH_DIR, name))
        assert xyz.shape == (nv, 3) and tets.shape == (nc, 4) and set(marker) == {0}
        assert tets.min() == 0 and tets.max() == nv - 1
        rp, ci = orc.scalar_pattern(nv, tets)
        assert rp[-1] == {54: 496, 178: 1936}[nv]          # nnz recorded in SURVEY section 8
    xyz, tets = meshes.read_dolfin_xml(os.path.join(REF_MESH_DIR, "multi_layer_torus.xml.zip"))
    sx, st = meshes.read_dolfin_xml(os.path.join(REF_MESH_DIR, "multi_layer_torus_compt1.xml.zip"))
    assert xyz.shape == (8160, 3) and tets.shape == (42840, 4) and st.shape == (12960, 4)
    phase = meshes.phase_from_submesh(xyz, tets, sx, st)
    assert phase.sum() == 12960
    rp, ci = orc.scalar_pattern(8160, tets)
    assert rp[-1] == 114000


def test_generators_are_conforming_and_deterministic():
    for xyz, tets in (meshes.box_mesh((0,) * 3, (1, 2, 3), 3, 4, 5), meshes.cylinder(3.0, 10.0, 3, 10, 4),
                      meshes.layered_cylinder()[:2], meshes.neuron_like(n_dend=2, soma_r=4, dend_len=20, h=1.0)):
        det, vol, _ = orc.tet_geometry(xyz, tets)
        assert vol.min() > 0
        f, cell, _ = orc.facets(tets)
        same = np.all(f[1:] == f[:-1], axis=1)
        # every facet is shared by at most two cells (conforming mesh)
        assert not np.any(same[1:] & same[:-1])
    xyz, tets = meshes.box_mesh((-2.5,) * 3, (2.5,) * 3, 4, 4, 4)
    assert abs(orc.tet_geometry(xyz, tets)[1].sum() - 125.0) < 1e-12
    a = meshes.ecs_slab(8, 8, 1, ncyl=5)
    b = meshes.ecs_slab(8, 8, 1, ncyl=5)
    assert np.array_equal(a[2], b[2])
    x2, t2 = meshes.rcm_order(*meshes.shuffle_vertices(xyz, tets, 3))
    assert abs(orc.tet_geometry(x2, t2)[1].sum() - 125.0) < 1e-12


def test_sweep_sharding_covers_every_unit_once():
    units = sweep.sweep_units(range(64), range(4))
    assert len(units) == 256
    for world in (1, 2, 4, 8):
        owned = sorted(u for r in range(world) for u in sweep.shard_units(len(units), r, world))
        assert owned == list(range(256))
        sizes = [len(sweep.shard_units(len(units), r, world)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
        # the batched sweep's sharding: same cover, every rank holds every b-value equally often
        shards = [sweep.shard_balanced(64, 4, r, world) for r in range(world)]
        assert sorted(u for s in shards for u in s) == list(range(256))
        for s in shards:
            bs = [units[u][1] for u in s]
            assert s == sorted(s) and all(bs.count(j) == 64 // world for j in range(4))
    assert sorted(u for r in range(3) for u in sweep.shard_balanced(7, 2, r, 3)) == list(range(14))   # ragged


def test_readers_triangle_meshes_and_markers(tmp_path):
    """gmsh v2 triangles (a 2-D mesh: dolfin-convert writes gdim 2), DOLFIN XML celltype triangle, and the cell
    marker file msh2xml writes (DmriFemLib.py:725-746)."""
    msh = tmp_path / "sq.msh"
    msh.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n5\n1 0 0 0\n2 1 0 0\n3 1 1 0\n4 0 1 0\n5 9 9 0\n$EndNodes\n"
                   "$Elements\n3\n1 1 2 7 7 1 2\n2 2 2 3 1 1 2 3\n3 2 2 4 2 1 3 4\n$EndElements\n")
    xy, tris, mk = meshes.read_gmsh2(str(msh))
    assert xy.shape == (4, 2) and tris.tolist() == [[0, 1, 2], [0, 2, 3]] and mk.tolist() == [3, 4]
    xml = tmp_path / "sq.xml"
    xml.write_text('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n'
                   '  <mesh celltype="triangle" dim="2">\n    <vertices size="4">\n'
                   '      <vertex index="0" x="0" y="0" />\n      <vertex index="1" x="1" y="0" />\n'
                   '      <vertex index="2" x="1" y="1" />\n      <vertex index="3" x="0" y="1" />\n    </vertices>\n'
                   '    <cells size="2">\n      <triangle index="0" v0="0" v1="1" v2="2" />\n'
                   '      <triangle index="1" v0="0" v1="2" v2="3" />\n    </cells>\n  </mesh>\n</dolfin>\n')
    xy2, tris2 = meshes.read_dolfin_xml(str(xml))
    assert np.array_equal(xy2, xy) and np.array_equal(tris2, tris)
    ops = orc.assemble(xy2, tris2, D=1.0)
    assert abs(ops.lumped.sum() - 1.0) < 1e-15
    pmk = tmp_path / "pmk_sq.xml"
    pmk.write_text('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n  <mesh_function>\n'
                   '    <mesh_value_collection type="uint" dim="2" size="2">\n'
                   '      <value cell_index="0" local_entity="0" value="3" />\n'
                   '      <value cell_index="1" local_entity="0" value="4" />\n'
                   '    </mesh_value_collection>\n  </mesh_function>\n</dolfin>\n')
    assert meshes.read_dolfin_markers(str(pmk)).tolist() == [3, 4]


def test_planar_periodic_gather_matches_oracle():
    """periodic.build_gather on a triangle mesh (boundary facets = edges) reproduces the oracle's independent
    restatement of WeakPseudoPeriodic_*.eval on non-matching opposite faces, pdir = (1,1,0), two compartments."""
    from dmri_fem_cloud_b200 import periodic
    n = 6
    xs, ys = np.linspace(-2, 2, n + 1), np.linspace(-1, 1.5, n + 1)
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    xy = np.column_stack([X.ravel(), Y.ravel()])
    idx = np.arange((n + 1) ** 2).reshape(n + 1, n + 1)
    a, b, c, d = idx[:-1, :-1].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel(), idx[:-1, 1:].ravel()
    tris = np.concatenate([np.stack([a, b, c], 1), np.stack([a, c, d], 1)])
    on = np.abs(xy[:, 0] - 2) < 1e-9
    inner = on & (np.abs(xy[:, 1] + 1) > 1e-9) & (np.abs(xy[:, 1] - 1.5) > 1e-9)
    xy[inner, 1] += 0.07
    xyz = orc.as_xyz3(xy)
    lo, hi, hmin, _ = orc.domain_sizes(xyz, tris)
    pdir = [1, 1, 0]
    ph = (np.linalg.norm(xy[tris].mean(axis=1), axis=1) < 0.9).astype(np.int32)
    ops = orc.assemble(xy, tris, ph, D=3e-3, kappa=1e-5, bnd_kappa_vertex=orc.periodic_marker(xyz, pdir, lo, hi, hmin))
    q, F = 0.3, 2.0
    g = np.array([0.6, 0.8, 0.0])
    rng = np.random.default_rng(0)
    u = rng.normal(size=ops.ndof) + 1j * rng.normal(size=ops.ndof)
    want = orc.periodic_term(xyz, tris, ops, pdir, lo, hi, q, g, 0.5)(u, F)
    dof, src, w, dx = periodic.build_gather(xy, tris, ph, pdir, lo, hi, ops.dof_vertex, ops.dof_comp)
    ubc = np.zeros(ops.ndof, complex)
    ubc[dof] = (np.where(src >= 0, u[np.maximum(src, 0)], 0) * w).sum(axis=1) * np.exp(1j * q * (dx @ g) * F)
    got = 0.5 * (ops.B @ ubc)
    assert np.abs(want).max() > 1e-4 and np.abs(got - want).max() <= 1e-15


def test_comri_cli_parsing_and_scheme_quirks():
    """comri drivers (comri/*/hpc-fenics-cpp/main.cpp): flags, defaults, gnorm/qvalue, FT closed at Delta+delta,
    loop `t < T + dt` vs `t < T`, shell phases."""
    from dmri_fem_cloud_b200 import comri
    p = comri.parse("one-comp", ["demo", "-m", "cyl.xml", "-b", "1000", "-d", "10600", "-D", "43100", "-k", "200",
                                 "-v", "2", "0", "0", "-K", "3e-3", "-N", "50", "-j", "10"])
    assert (p["mesh"], p["b"], p["delta"], p["Delta"], p["dt"], p["is_dt"], p["K"], p["N"], p["nskip"]) == \
        ("cyl.xml", 1000.0, 10600.0, 43100.0, 200.0, True, 3e-3, 50, 10)
    assert p["g"] == (1.0, 0.0, 0.0)                                   # normalised (main.cpp:176-177)
    d = comri.parse("one-comp", ["demo"])
    assert (d["b"], d["delta"], d["Delta"], d["K"], d["g"], d["is_dt"]) == (1000.0, 40000.0, 40000.0, 2.4e-3,
                                                                           (0.0, 1.0, 0.0), False)
    t2 = comri.parse("two-comp", ["demo", "-c", "cell.xml", "-p", "1e-5"])
    assert (t2["cell"], t2["kappa"], t2["b"], t2["K"]) == ("cell.xml", 1e-5, 4000.0, 3e-3)
    ml = comri.parse("multilayer", ["demo", "-k", "5", "-v", "1", "0", "0", "-p", "9"])
    assert (ml["is_dt"], ml["g"], ml["kappa"]) == (False, (0.0, 0.0, 1.0), 5e-5)     # those flags do not exist there
    assert comri.FT(0.0, 10.0, 30.0) == 1.0 and comri.FT(10.0, 10.0, 30.0) == 0.0
    assert comri.FT(30.0, 10.0, 30.0) == -1.0 and comri.FT(40.0, 10.0, 30.0) == -1.0      # `t <= Delta + delta`
    assert len(comri.time_grid(40.0, 10.0, "one-comp")) == 5 and len(comri.time_grid(40.0, 10.0, "multilayer")) == 4
    mid = np.array([[0.5, 0, 0], [1.2, 0, 0], [1.7, 0, 0], [3.0, 0, 0]])
    assert comri.shell_phase(mid, "two-comp").tolist() == [0, 1, 0, 0]
    mid = np.array([[20.0, 0, 0], [26.0, 0, 0], [28.0, 0, 0], [40.0, 0, 0]])
    assert comri.shell_phase(mid, "multilayer").tolist() == [0, 1, 0, 0]


def test_preprocess_phase_and_cell_fields(tmp_path):
    """CreatePhaseFunc (DmriFemLib.py:748-799) + the DG0 fields of PreprocessingMultiCompt.py:118-152 -> .npz."""
    from dmri_fem_cloud_b200 import preprocess
    xyz, tets, lay = meshes.layered_cylinder((5.0, 7.5, 10.0), 2.0, (2, 1, 1), 8, 1)
    ph, plist = preprocess.create_phase_func(xyz, tets, partition_marker=lay)
    assert np.array_equal(ph, lay % 2) and sorted(plist) == [0, 1, 2]
    sub = lambda k: (xyz, tets[lay == k])                 # a compartment as a cell subset of the mesh
    ph2, plist2, pm2 = preprocess.create_phase_func(xyz, tets, evengroup=[sub(2)], oddgroup=[sub(1)])
    assert np.array_equal(pm2[lay == 1], np.full((lay == 1).sum(), 1)) and set(pm2[lay == 2]) == {2} and set(pm2[lay == 0]) == {0}
    assert np.array_equal(ph2, (lay == 1).astype(np.int32)) and plist2 == [0, 2, 1]
    f = preprocess.cell_fields(pm2, [3e-3, 1e-3, 2e-3], [1e6, 4e4, 5e4], [1, 0, 1])
    assert set(f["d11"][lay == 1]) == {1e-3} and set(f["T2"][lay == 2]) == {5e4} and not f["d01"].any()
    assert set(f["ic"][lay == 1]) == {0.0}
    # command line: marker file written the way msh2xml does, output readable by the solver CLI
    np.savez(tmp_path / "m.npz", xyz=xyz, tets=tets)
    pmk = tmp_path / "pmk.xml"
    pmk.write_text('<dolfin><mesh_function><mesh_value_collection type="uint" dim="3" size="%d">\n' % len(tets) +
                   "".join('<value cell_index="%d" local_entity="0" value="%d" />\n' % (c, v) for c, v in enumerate(lay)) +
                   '</mesh_value_collection></mesh_function></dolfin>\n')
    monkey_load = preprocess._load
    preprocess._load = lambda p: (xyz, tets)
    try:
        rc = preprocess.main(["preprocess", "-m", "mesh.xml", "-pmk", str(pmk), "-D0", "3e-3", "1e-3", "3e-3", "-T2", "1e6",
                              "4e4", "4e4", "-o", str(tmp_path / "files.h5")])
    finally:
        preprocess._load = monkey_load
    from dmri_fem_cloud_b200 import cli
    z = cli.load_input(str(tmp_path / "files.h5"))         # DOLFIN HDF5 container, like the reference's output
    assert rc == 0 and np.array_equal(z["phase"], lay % 2) and set(z["d00"][lay == 1]) == {1e-3} and set(z["ic"]) == {1.0}
    assert all(k in z for k in ("xyz", "tets", "phase", "T2", "ic", "d00", "d22"))
    preprocess._load = lambda p: (xyz, tets)
    try:                                                   # `-o x.npz` keeps the NumPy side format
        assert preprocess.main(["preprocess", "-m", "mesh.xml", "-pmk", str(pmk), "-o", str(tmp_path / "files.npz")]) == 0
    finally:
        preprocess._load = monkey_load
    data = cli.load_input(str(tmp_path / "files.npz"))
    assert all(k in data for k in ("xyz", "tets", "phase", "T2", "ic", "d00", "d22")) and np.array_equal(data["phase"], z["phase"])


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints ONE JSON line with the
    contract's keys; under torchrun only rank 0 prints.  Tiny workload, CPU only."""
    import json
    import subprocess
    import sys as _sys
    cmd = [_sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
           "--cpu-sample-steps", "2", "--n-box", "8"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, check=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "DOF-steps/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    silent = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert silent.returncode == 0 and silent.stdout.strip() == ""


def test_bench_loop_roofline_arithmetic():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    r = b.loop_roofline(nnz=1000, n=100, iters=10, nsteps=2, loop_ms=1.0, peak_gbs=1.0)
    spmv = 20 * 1000 + 36 * 100
    assert r["algorithmic_bytes_per_iteration"] == 2 * spmv + 224 * 100
    assert r["algorithmic_bytes_per_solve"] == 10 * (2 * spmv + 224 * 100) + 2 * spmv
    assert abs(r["achieved"] - r["algorithmic_bytes_per_solve"] / 1e-3 / 1e9) < 1e-12 and r["us_per_iteration"] == 100.0


def test_periodic_locate_falls_back_to_exhaustive_search():
    """periodic._locate: on a strongly graded face the containing triangle is not among the nearest centroids; the
    point must still get its true triangle and weights (the reference evaluates in the containing cell)."""
    from dmri_fem_cloud_b200 import periodic
    big = [[(0.0, 0.0), (100.0, 0.0), (0.0, 1.0)], [(100.0, 0.0), (100.0, 1.0), (0.0, 1.0)]]
    rng = np.random.default_rng(3)
    small = []
    for _ in range(40):                      # a cloud of tiny triangles next to the query point, none containing it
        c = np.array([90.0, 3.0]) + rng.uniform(-0.5, 0.5, 2)
        small.append([tuple(c), tuple(c + [0.05, 0.0]), tuple(c + [0.0, 0.05])])
    tri_xy = np.array(big + small)
    pts = np.array([[90.0, 0.05], [10.0, 0.95]])
    tri, w = periodic._locate(pts, tri_xy, ncand=16)
    assert tri.tolist() == [0, 1]
    assert (w >= -1e-12).all() and np.allclose(w.sum(axis=1), 1.0)
    assert np.allclose((w[:, :, None] * tri_xy[tri]).sum(axis=1), pts)
