"""Row partition of one mesh (partition.py), host logic only: the local sub-mesh operators (built here with the
ORACLE standing in for the GPU assembly) plus the halo send lists must reproduce the global operator.
Also the world_size-2 gloo run of the same plumbing through torch.distributed."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as entry  # noqa: E402

entry.load_package()
from dmri_fem_cloud_b200 import meshes, partition  # noqa: E402
import bt_oracle as orc  # noqa: E402


def _mesh(two_comp=True):
    xyz, tets, phase = meshes.box_with_sphere(half=10.0, n=7, radius=5.0)
    xyz, tets = meshes.shuffle_vertices(xyz, tets, seed=3)
    xyz, tets = meshes.rcm_order(xyz, tets)
    return xyz, tets, (phase if two_comp else None)


def _operator(ops, c=0.37, g=(0.3, -0.5, 0.8), dt=5.0, theta=0.5):
    P = ops.M / dt + theta * (ops.S + ops.R + ops.I)
    J = g[0] * ops.Jx + g[1] * ops.Jy + g[2] * ops.Jz
    return (P + 1j * theta * c * J).tocsr()


def _rank_state(xyz, tets, phase, bounds, rank):
    part = partition.local_part(tets, bounds, rank)
    ph = None if phase is None else phase[part.cells]
    ops = orc.assemble(xyz[part.l2g], part.tets, ph, D=2e-3, invT2=1e-3, kappa=1e-2)
    dv, dc = ops.dof_vertex, ops.dof_comp
    n_own = int(np.searchsorted(dv, part.nv_own))
    n_int = int(np.searchsorted(dv, part.nv_int))
    return dict(part=part, ops=ops, dv=dv, dc=dc, n_own=n_own, n_int=n_int)


@pytest.mark.parametrize("world", [1, 2, 3, 5])
@pytest.mark.parametrize("two_comp", [False, True])
def test_partitioned_operator_equals_global(world, two_comp):
    xyz, tets, phase = _mesh(two_comp)
    gops = orc.assemble(xyz, tets, phase, D=2e-3, invT2=1e-3, kappa=1e-2)
    A = _operator(gops)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(gops.ndof) + 1j * rng.standard_normal(gops.ndof)
    y = A @ x
    bounds = partition.block_bounds(len(xyz), world, tets)
    assert bounds[0] == 0 and bounds[-1] == len(xyz) and (np.diff(bounds) > 0).all()
    st = [_rank_state(xyz, tets, phase, bounds, r) for r in range(world)]
    shift = 5   # any halo shift must come back in the slots
    req = [partition.halo_requests(s["part"], bounds, s["dv"], s["dc"], s["n_own"], shift) for s in st]
    xl = []
    for s in st:
        gd = gops.vc2dof[s["part"].l2g[s["dv"]], s["dc"]]
        assert (gd[:s["n_own"]] >= 0).all()
        v = np.zeros(s["ops"].ndof + shift, dtype=complex)
        v[:s["n_own"]] = x[gd[:s["n_own"]]]
        xl.append(v)
        s["gd"] = gd
    n_send_total = 0
    for s in st:
        src, dr, ds, recv = partition.send_list(s["part"], req, s["dv"], s["dc"], s["n_own"])
        assert (src >= s["n_int"]).all(), "interior dofs must never be sent"
        for e in range(len(src)):
            xl[dr[e]][ds[e]] = xl[s["part"].rank][src[e]]
        n_send_total += len(src)
        want = np.zeros(world, dtype=np.int32)
        if len(req[s["part"].rank]):
            want[np.unique(req[s["part"].rank][:, 0])] = 1
        assert np.array_equal(recv, want)
    assert n_send_total == sum(len(q) for q in req)
    owned_total, vol = 0, 0.0
    for s, v in zip(st, xl):
        n_own, n_int, ops = s["n_own"], s["n_int"], s["ops"]
        Al = _operator(ops)
        assert Al[:n_int, n_own:].nnz == 0, "rows before n_int must not reference halo columns"
        xloc = np.concatenate([v[:n_own], v[n_own + shift:]])
        yl = (Al @ xloc)[:n_own]
        np.testing.assert_allclose(yl, y[s["gd"][:n_own]], rtol=1e-12, atol=1e-13)
        owned_total += n_own
        vol += ops.lumped[:n_own].sum()
    assert owned_total == gops.ndof
    np.testing.assert_allclose(vol, gops.lumped.sum(), rtol=1e-13)


def test_thread_comm_collectives():
    import threading
    comms = partition.ThreadComm.make(3)
    out = [None] * 3

    def run(c):
        g = c.allgather(("r", c.rank))
        s = c.sum([c.rank + 1.0, 2.0])
        c.barrier()
        out[c.rank] = (g, s.tolist())

    th = [threading.Thread(target=run, args=(c,)) for c in comms]
    [t.start() for t in th]
    [t.join(30) for t in th]
    for r in range(3):
        assert out[r] == ([("r", 0), ("r", 1), ("r", 2)], [6.0, 6.0])


def _gloo_worker(rank, world, port, q, kind="gloo"):
    dist = None
    if kind == "socket":      # the torch-free plumbing: plain TCP through rank 0
        comm = partition.SocketComm(rank, world, "127.0.0.1", port)
    else:
        import torch.distributed as dist
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        comm = partition.TorchComm(dist)
    xyz, tets, phase = _mesh(True)
    gops = orc.assemble(xyz, tets, phase, D=2e-3, invT2=1e-3, kappa=1e-2)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(gops.ndof) + 1j * rng.standard_normal(gops.ndof)
    y = _operator(gops) @ x
    bounds = partition.block_bounds(len(xyz), world, tets)
    s = _rank_state(xyz, tets, phase, bounds, rank)
    req = partition.halo_requests(s["part"], bounds, s["dv"], s["dc"], s["n_own"], 0)
    allreq = comm.allgather(req)
    src, dr, ds, recv = partition.send_list(s["part"], allreq, s["dv"], s["dc"], s["n_own"])
    gd = gops.vc2dof[s["part"].l2g[s["dv"]], s["dc"]]
    xl = np.zeros(s["ops"].ndof, dtype=complex)
    xl[:s["n_own"]] = x[gd[:s["n_own"]]]
    for r, (pr, pslot, pval) in enumerate(comm.allgather((dr, ds, xl[src]))):   # the halo exchange
        m = pr == rank
        xl[pslot[m]] = pval[m]
        assert bool(recv[r]) == bool(m.any())
    yl = (_operator(s["ops"]) @ xl)[:s["n_own"]]
    err = float(np.abs(yl - y[gd[:s["n_own"]]]).max())
    tot = comm.sum([s["n_own"]])
    from dmri_fem_cloud_b200 import sweep                      # the signal table of a sharded sweep through the same object
    full = sweep.gather_signals(world, [rank], np.array([rank + 1.0]), comm)
    assert np.array_equal(full, np.arange(1.0, world + 1.0))
    comm.barrier()
    q.put((rank, err, int(tot[0]), gops.ndof))
    if dist is not None:
        dist.destroy_process_group()
    else:
        comm.close()


@pytest.mark.parametrize("kind", ["gloo", "socket"])
def test_two_rank_partition_gloo(kind):
    """The N > 1 host plumbing on two CPU processes: torch.distributed (gloo) and the torch-free SocketComm."""
    import multiprocessing as mp
    sk = socket.socket()
    sk.bind(("127.0.0.1", 0))
    port = sk.getsockname()[1]
    sk.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q, kind)) for r in range(2)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, tot, ndof in out:
        assert err < 1e-12 and tot == ndof
