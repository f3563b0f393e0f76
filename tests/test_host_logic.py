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
