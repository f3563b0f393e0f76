"""Row-partitioned solve of ONE mesh through the C-ABI.

* local operator: every rank's handle (sub-mesh, owned rows) against the ORACLE's global operator;
* the whole protocol (halo pushes into peer vectors, in-kernel all-reduce, lazy halo wait) with the ranks as
  THREADS on one GPU -- small meshes, so that every rank's grids are co-resident -- against the single-handle
  solve: signals <= 1e-10 relative (the dot products are summed in a different order), same step count;
* one process per GPU over CUDA IPC when the box has >= 2 GPUs (skipped otherwise)."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

import bt_oracle as orc
from dmri_fem_cloud_b200 import btfem, meshes, partition

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mesh(two_comp, n=7):
    xyz, tets, phase = meshes.box_with_sphere(half=10.0, n=n, radius=5.0)
    xyz, tets = meshes.shuffle_vertices(xyz, tets, seed=3)
    xyz, tets = meshes.rcm_order(xyz, tets)
    return xyz, tets, (phase.astype(np.int32) if two_comp else None)


COEF = dict(D=2e-3, invT2=1e-4, kappa=1e-3)


def _pgse(k=200.0, delta=2000.0, Delta=5000.0, b=1000.0):
    seq = orc.pgse(delta, Delta)
    ts = orc.time_grid(seq.T, k)
    q = seq.q_from_b(b)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[seq.f(0.0)], f[:-1]])
    return k, q * f, q * fp


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("two_comp", [False, True])
def test_local_operator_rows(world, two_comp):
    xyz, tets, phase = _mesh(two_comp)
    gops = orc.assemble(xyz, tets, phase, D=COEF["D"], invT2=COEF["invT2"], kappa=COEF["kappa"])
    g = np.array([0.3, -0.5, 0.8])
    dt, theta, c = 5.0, 0.5, 0.37
    A = (gops.M / dt + theta * (gops.S + gops.R + gops.I) + 1j * theta * c * (g[0] * gops.Jx + g[1] * gops.Jy + g[2] * gops.Jz))
    rng = np.random.default_rng(1)
    x = rng.standard_normal(gops.ndof) + 1j * rng.standard_normal(gops.ndof)
    y = A @ x
    bounds = partition.block_bounds(len(xyz), world, tets)
    owned = 0
    for rank in range(world):
        part = partition.local_part(tets, bounds, rank)
        with btfem.BTFem(0) as fem:
            fem.set_mesh(xyz[part.l2g], part.tets, None if phase is None else phase[part.cells])
            fem.set_diffusion(COEF["D"])
            fem.set_relaxation(COEF["invT2"])
            if two_comp:
                fem.set_permeability(COEF["kappa"])
            fem.set_partition(part.nv_own, part.nv_int)
            fem.assemble()
            n_own, n_int, shift = fem.partition_sizes()
            assert (n_own + shift) % 8 == 0 and 0 <= shift < 8
            dv, dc = fem.dofmap()
            gd = gops.vc2dof[part.l2g[dv], dc]
            yl = fem.spmv(dt, theta, c, g, x[gd])
            assert np.max(np.abs(yl[:n_own] - y[gd[:n_own]])) <= 1e-13 * np.max(np.abs(y))   # SpMV tolerance
            assert np.all(yl[n_own:] == 0)
            owned += n_own
    assert owned == gops.ndof


def _single(xyz, tets, phase, k, cA, cb, g, **kw):
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        fem.set_diffusion(COEF["D"])
        fem.set_relaxation(COEF["invT2"])
        if phase is not None:
            fem.set_permeability(COEF["kappa"])
        fem.assemble()
        res = fem.solve(k, 0.5, cA, cb, g, want_iters=True, **kw)
        dv, dc = fem.dofmap()
        return res, fem.solution(), dv, dc


def _thread_ranks(world, xyz, tets, phase, k, cA, cb, g, **kw):
    comms = partition.ThreadComm.make(world)
    out, err = [None] * world, [None] * world

    def run(comm):
        try:
            d = partition.DistBTFem(xyz, tets, comm, device=0, phase=phase)
            d.set_diffusion(COEF["D"])
            d.set_relaxation(COEF["invT2"])
            if phase is not None:
                d.set_permeability(COEF["kappa"])
            d.assemble()
            res = d.solve(k, 0.5, cA, cb, g, want_iters=True, **kw)
            res2 = d.solve(k, 0.5, cA, cb, g, **kw)          # sequence numbers carry over between solves
            sol = d.global_solution()
            comm.barrier()
            out[comm.rank] = (res, res2, sol, d.n_send, d.n_own, d.n_int)
            d.close()
        except Exception as e:   # noqa: BLE001
            err[comm.rank] = e
            try:
                comm.sh.barrier.abort()
            except Exception:
                pass

    th = [threading.Thread(target=run, args=(c,)) for c in comms]
    [t.start() for t in th]
    [t.join(300) for t in th]
    for e in err:
        if e is not None:
            raise e
    return out


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("two_comp", [False, True])
def test_thread_ranks_match_single_handle(world, two_comp, monkeypatch):
    monkeypatch.setenv("BTFEM_COMM_TIMEOUT_MS", "5000")
    xyz, tets, phase = _mesh(two_comp)
    k, cA, cb = _pgse()
    g = np.array([0.0, 1.0, 0.0])
    ref, uref, dv, dc = _single(xyz, tets, phase, k, cA, cb, g)
    out = _thread_ranks(world, xyz, tets, phase, k, cA, cb, g)
    for rank, (res, res2, sol, n_send, n_own, n_int) in enumerate(out):
        assert abs(res["signal"] - ref["signal"]) <= 1e-10 * abs(ref["signal"])      # signal tolerance (1e-8 north star)
        assert abs(res["voi"] - ref["voi"]) <= 1e-12 * ref["voi"]
        assert abs(res["whole_vol"] - ref["whole_vol"]) <= 1e-12 * ref["whole_vol"]
        assert res2["signal"] == res["signal"]                                        # deterministic, re-entrant
        assert res["signal"] == out[0][0]["signal"]                                   # bit-identical on every rank
        assert abs(int(res["total_iters"]) - int(ref["total_iters"])) <= max(2, ref["total_iters"] // 50)
        gv, cp, u = sol
        assert np.array_equal(gv, dv) and np.array_equal(cp, dc)
        assert np.max(np.abs(u - uref)) <= 1e-9 * np.max(np.abs(uref))
        if world > 1:
            assert n_send > 0 and n_int < n_own


def test_nonzero_guess_partitioned(monkeypatch):
    monkeypatch.setenv("BTFEM_COMM_TIMEOUT_MS", "5000")
    xyz, tets, phase = _mesh(True)
    k, cA, cb = _pgse()
    g = np.array([1.0, 0.0, 0.0])
    ref, _, _, _ = _single(xyz, tets, phase, k, cA, cb, g, nonzero_guess=True)
    out = _thread_ranks(2, xyz, tets, phase, k, cA, cb, g, nonzero_guess=True)
    assert abs(out[0][0]["signal"] - ref["signal"]) <= 1e-10 * abs(ref["signal"])


@pytest.mark.parametrize("world,two_comp,pdir", [(2, False, [1, 0, 0]), (3, True, [1, 1, 0]), (2, True, [0, 1, 1])])
def test_weak_periodic_partitioned(world, two_comp, pdir, monkeypatch):
    """configs[3] shape: weak pseudo-periodic BC on a partitioned mesh -- the mirrored sources of the gather live
    on other ranks and travel with the per-step u push."""
    from dmri_fem_cloud_b200 import periodic
    monkeypatch.setenv("BTFEM_COMM_TIMEOUT_MS", "5000")
    xyz, tets, phase = _mesh(two_comp)
    lo, hi = xyz.min(axis=0), xyz.max(axis=0)
    seq = orc.pgse(2000.0, 5000.0)
    k = 200.0
    ts = orc.time_grid(seq.T, k)
    q = seq.q_from_b(800.0)
    f = np.array([seq.f(t) for t in ts])
    fp = np.concatenate([[f[0]], f[:-1]])
    Fp = np.concatenate([[seq.F(0.0)], [seq.F(t) for t in ts[:-1]]])
    g = np.array([1.0, 0.5, 0.25])
    g /= np.linalg.norm(g)
    with btfem.BTFem(0) as fem:
        fem.set_mesh(xyz, tets, phase)
        hmin, _ = fem.mesh_stats()
        kappa_e, tol = 3e-3 / hmin, 1e-2 * hmin
        fem.set_diffusion(COEF["D"])
        if two_comp:
            fem.set_permeability(COEF["kappa"])
        fem.set_periodic(pdir, kappa_e, tol, lo, hi)
        fem.assemble()
        dv, dc = fem.dofmap()
        fem.set_periodic_gather(*periodic.build_gather(xyz, tets, phase, pdir, lo, hi, dv, dc))
        ref = fem.solve(k, 0.5, q * f, q * fp, g, q=q, Fb=Fp)
        uref = fem.solution()
        neu_sig = None
    comms = partition.ThreadComm.make(world)
    out, err = [None] * world, [None] * world

    def run(comm):
        try:
            d = partition.DistBTFem(xyz, tets, comm, device=0, phase=phase)
            assert abs(d.mesh_stats()[0] - hmin) <= 1e-15 * hmin
            d.set_diffusion(COEF["D"])
            if two_comp:
                d.set_permeability(COEF["kappa"])
            d.set_periodic(pdir, kappa_e, tol, lo, hi)
            d.assemble()
            res = d.solve(k, 0.5, q * f, q * fp, g, q=q, Fb=Fp)
            out[comm.rank] = (res, d.global_solution(), d.n_send_u)
            comm.barrier()
            d.close()
        except Exception as e:   # noqa: BLE001
            err[comm.rank] = e
            comm.sh.barrier.abort()

    th = [threading.Thread(target=run, args=(c,)) for c in comms]
    [t.start() for t in th]
    [t.join(300) for t in th]
    for e in err:
        if e is not None:
            raise e
    assert sum(o[2] for o in out) > 0, "the test mesh must put mirrored sources on other ranks"
    for res, sol, _ in out:
        assert abs(res["signal"] - ref["signal"]) <= 1e-10 * abs(ref["signal"])
        assert np.max(np.abs(sol[2] - uref)) <= 1e-9 * np.max(np.abs(uref))


def test_lost_peer_times_out(monkeypatch):
    """A rank whose peer never shows up must fail with BTFEM_ECOMM, not hang."""
    monkeypatch.setenv("BTFEM_COMM_TIMEOUT_MS", "300")
    xyz, tets, phase = _mesh(False)
    k, cA, cb = _pgse()
    comms = partition.ThreadComm.make(2)
    res = [None, None]

    def run(comm):
        d = partition.DistBTFem(xyz, tets, comm, device=0, phase=None)
        d.set_diffusion(COEF["D"])
        d.set_relaxation(0.0)
        d.assemble()
        if comm.rank == 0:
            try:
                d.fem.solve(k, 0.5, cA, cb, np.array([1.0, 0, 0]))
                res[0] = "no error"
            except btfem.BTFemError as e:
                res[0] = e.code
        comm.barrier()
        d.close()

    th = [threading.Thread(target=run, args=(c,)) for c in comms]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert res[0] == -8


def _gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.skipif(_gpu_count() < 2, reason="needs >= 2 GPUs (one process per GPU over CUDA IPC)")
def test_two_processes_two_gpus():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "scripts", "dist_solve.py"), "--nbox", "12", "--check"]
    env = dict(os.environ, BTFEM_COMM_TIMEOUT_MS="10000")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    assert r["rel_signal_err"] <= 1e-10 and r["world"] == 2
