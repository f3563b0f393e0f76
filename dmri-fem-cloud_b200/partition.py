"""One mesh row-partitioned over several GPUs: the host logic.

The reference scales one big mesh (the 504 k-vertex extracellular space, realistic neurons) with
`mpirun -n N python3 GCloudDmriSolver.py ...` (README.md:89-94): DOLFIN distributes the mesh, PETSc's MatMult
scatters ghost values and KSPSolve all-reduces the dot products (third party).  Here every rank (one process per
GPU) owns a contiguous block of vertices of a locality-ordered numbering and builds a btfem handle on its LOCAL
sub-mesh; libbtfem exchanges halo entries and dot products through peer memory inside its kernels
(include/btfem.h, "one mesh row-partitioned over several GPUs").  This module only does bookkeeping with numpy:

    block_bounds   contiguous vertex blocks of (about) equal work
    local_part     the sub-mesh of a rank: cells touching an owned vertex, local vertex numbering
                   [owned, no peer needs it | owned, peers need it | halo], local <-> global maps
    halo_requests / send_list   who sends which dof into which vector element of whom
    DistBTFem      btfem.BTFem-like object (set_* / assemble / solve) on top of the above
    TorchComm / ThreadComm      the tiny collective interface the set-up needs (all-gather of python objects,
                   barrier, sum): torch.distributed for one process per GPU, threads for in-process tests
"""
import threading

import numpy as np

from . import btfem as _bt
from . import periodic as _periodic


# ------------------------------------------------------------------------------------------------ partition

def block_bounds(nv, world, tets=None):
    """Vertex blocks [b[r], b[r+1]) with about equal work.  Work of a vertex ~ its row length, estimated by
    1 + number of incident cells when `tets` is given, else 1."""
    if tets is None:
        w = np.ones(nv)
    else:
        w = 1.0 + np.bincount(np.asarray(tets).ravel(), minlength=nv)
    c = np.concatenate([[0.0], np.cumsum(w)])
    targets = c[-1] * np.arange(1, world) / world
    inner = np.searchsorted(c, targets, side="left")
    b = np.concatenate([[0], inner, [nv]]).astype(np.int64)
    for r in range(1, world + 1):          # every rank owns at least one vertex
        b[r] = max(b[r], b[r - 1] + 1)
    if b[world] != nv:
        raise ValueError("more ranks than vertices")
    return b


class Part:
    """Sub-mesh of one rank.  Local vertex l is global vertex l2g[l]; [0,nv_int) owned and unseen by peers,
    [nv_int,nv_own) owned and needed by peers, [nv_own,nv) halo.  cells: global ids of the local cells."""

    def __init__(self, rank, lo, hi, cells, l2g, nv_own, nv_int, tets, nv_global):
        self.rank, self.lo, self.hi = rank, lo, hi
        self.cells, self.l2g, self.nv_own, self.nv_int, self.tets = cells, l2g, nv_own, nv_int, tets
        self.nv_global = nv_global

    def g2l(self):
        m = np.full(self.nv_global, -1, dtype=np.int64)
        m[self.l2g] = np.arange(len(self.l2g))
        return m


def local_part(tets, bounds, rank):
    tets = np.asarray(tets)
    nv = int(bounds[-1])
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    mine = (tets >= lo) & (tets < hi)
    touch = mine.any(axis=1)
    cells = np.nonzero(touch)[0]
    lt = tets[cells]
    foreign_cell = ~mine[cells].all(axis=1)
    seen = np.unique(lt)
    halo = seen[(seen < lo) | (seen >= hi)]
    bnd = np.unique(lt[foreign_cell])
    bnd = bnd[(bnd >= lo) & (bnd < hi)]
    is_bnd = np.zeros(hi - lo, dtype=bool)
    is_bnd[bnd - lo] = True
    own = np.arange(lo, hi)
    l2g = np.concatenate([own[~is_bnd], own[is_bnd], halo]).astype(np.int64)
    g2l = np.full(nv, -1, dtype=np.int64)
    g2l[l2g] = np.arange(len(l2g))
    return Part(rank, lo, hi, cells, l2g, hi - lo, int((~is_bnd).sum()), g2l[lt].astype(np.int32), nv)


def halo_requests(part, bounds, dof_vertex, dof_comp, n_own, halo_shift):
    """One row per halo dof of this rank: (owner rank, global vertex, compartment, vector element here)."""
    j = np.arange(n_own, len(dof_vertex))
    gv = part.l2g[dof_vertex[j]]
    owner = np.searchsorted(bounds, gv, side="right") - 1
    return np.stack([owner, gv, dof_comp[j].astype(np.int64), j + halo_shift], axis=1).astype(np.int64)


def send_list(part, requests, dof_vertex, dof_comp, n_own):
    """requests[r] = halo_requests of rank r.  Returns (src, dst_rank, dst_slot, recv_from) for part.rank."""
    world = len(requests)
    vc2dof = np.full((len(part.l2g), 2), -1, dtype=np.int64)
    vc2dof[dof_vertex, dof_comp] = np.arange(len(dof_vertex))
    g2l = part.g2l()
    src, dst_rank, dst_slot = [], [], []
    for r in range(world):
        if r == part.rank:
            continue
        q = requests[r]
        q = q[q[:, 0] == part.rank]
        if len(q) == 0:
            continue
        lv = g2l[q[:, 1]]
        d = vc2dof[lv, q[:, 2]]
        if (lv < 0).any() or (d < 0).any() or (d >= n_own).any():
            raise RuntimeError("rank %d: peer %d asks for a dof this rank does not own" % (part.rank, r))
        src.append(d)
        dst_rank.append(np.full(len(q), r))
        dst_slot.append(q[:, 3])
    recv_from = np.zeros(world, dtype=np.int32)
    mine = requests[part.rank]
    if len(mine):
        recv_from[np.unique(mine[:, 0])] = 1
    cat = lambda a: np.concatenate(a).astype(np.int32) if a else np.zeros(0, dtype=np.int32)
    return cat(src), cat(dst_rank), cat(dst_slot), recv_from


def global_dofmap(nv, tets, phase=None):
    """The single-handle numbering of the whole mesh: active (vertex, compartment) pairs, vertex-major
    (csrc/setup.cu:bt_build_dofmap).  Returns dof_vertex, dof_comp, vc2dof (nv,2)."""
    tets = np.asarray(tets)
    active = np.zeros((nv, 2), dtype=bool)
    if phase is None:
        active[np.unique(tets), 0] = True
    else:
        phase = np.asarray(phase)
        for c in (0, 1):
            active[np.unique(tets[phase == c]), c] = True
    flat = active.ravel()
    vc2dof = np.where(flat, np.cumsum(flat) - 1, -1).reshape(nv, 2).astype(np.int64)
    dv, dc = np.nonzero(active)
    return dv.astype(np.int64), dc.astype(np.int64), vc2dof


def periodic_plan(part, bounds, gather, gvc2dof, gdof_vertex, gdof_comp, dof_vertex, dof_comp):
    """Restrict the GLOBAL periodic gather (periodic.build_gather on the whole mesh, global dof ids) to the
    boundary dofs this rank holds.  Sources held locally become local dofs; sources held only by a peer become
    entry k of this rank's periodic source buffer (src = -2-k) and a request row (owner, vertex, comp, k)."""
    dof_g, src_g, w, dx = gather
    lg = gvc2dof[part.l2g[dof_vertex], dof_comp]            # global id of every local dof
    g2l = np.full(len(gdof_vertex), -1, dtype=np.int64)
    g2l[lg] = np.arange(len(lg))
    sel = g2l[dof_g] >= 0
    dof_l = g2l[dof_g[sel]]
    sg = src_g[sel].astype(np.int64)
    sl = np.where(sg >= 0, g2l[np.maximum(sg, 0)], -1)
    missing = (sg >= 0) & (sl < 0)
    need = np.unique(sg[missing])
    k_of = np.full(len(gdof_vertex), -1, dtype=np.int64)
    k_of[need] = np.arange(len(need))
    sl[missing] = -2 - k_of[sg[missing]]
    gv = gdof_vertex[need]
    owner = np.searchsorted(bounds, gv, side="right") - 1
    req = np.stack([owner, gv, gdof_comp[need], np.arange(len(need))], axis=1).astype(np.int64) if len(need) \
        else np.zeros((0, 4), dtype=np.int64)
    return (dof_l.astype(np.int32), sl.astype(np.int32), w[sel], dx[sel]), req


# ------------------------------------------------------------------------------------------------ collectives

class TorchComm:
    """torch.distributed (any backend for the data path; python objects travel through a gloo group)."""

    def __init__(self, dist, group=None):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.group = group
        if group is None and dist.get_backend() != "gloo":
            self.group = dist.new_group(backend="gloo")

    def allgather(self, obj):
        out = [None] * self.world
        self.dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self):
        self.dist.barrier(group=self.group)

    def sum(self, a):
        import torch
        t = torch.from_numpy(np.array(a, dtype=np.float64, copy=True))
        self.dist.all_reduce(t, group=self.group)
        return t.numpy()

    def max(self, a):
        import torch
        t = torch.from_numpy(np.array(a, dtype=np.float64, copy=True))
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t.numpy()


class ThreadComm:
    """Ranks as threads of one process (tests; several handles on one or more GPUs)."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots = [None] * world

    @staticmethod
    def make(world):
        sh = ThreadComm._Shared(world)
        return [ThreadComm(sh, r) for r in range(world)]

    def __init__(self, shared, rank):
        self.sh, self.rank, self.world = shared, rank, shared.world

    def allgather(self, obj):
        self.sh.slots[self.rank] = obj
        self.sh.barrier.wait()
        out = list(self.sh.slots)
        self.sh.barrier.wait()
        return out

    def barrier(self):
        self.sh.barrier.wait()

    def sum(self, a):
        parts = self.allgather(np.array(a, dtype=np.float64, copy=True))
        tot = parts[0].copy()
        for p in parts[1:]:      # rank order: identical on every rank
            tot = tot + p
        return tot

    def max(self, a):
        return np.max(np.stack(self.allgather(np.array(a, dtype=np.float64, copy=True))), axis=0)


class SocketComm:
    """The same tiny collective interface over plain TCP sockets: one process per GPU WITHOUT torch in the host
    plumbing (the data path never used it: halo entries and dot products travel through peer memory inside libbtfem's
    kernels).  Rank 0 listens on (addr, port); every collective is an all-gather of pickled python objects through it.
    Launch with any launcher that sets RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun, mpirun wrappers, a
    shell loop): `SocketComm.from_env()`."""

    def __init__(self, rank, world, addr="127.0.0.1", port=29555, timeout=600.0):
        import socket
        self.rank, self.world = int(rank), int(world)
        self._peers = []          # rank 0: sockets of ranks 1..world-1, in rank order
        self._sock = None         # other ranks: socket to rank 0
        if self.world == 1:
            return
        if self.rank == 0:
            srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
            srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            srv.bind((addr, int(port)))
            srv.listen(self.world)
            srv.settimeout(timeout)
            got = {}
            while len(got) < self.world - 1:
                c, _ = srv.accept()
                c.settimeout(timeout)
                c.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                got[self._recv(c)] = c
            srv.close()
            self._peers = [got[r] for r in range(1, self.world)]
        else:
            import time
            deadline = time.time() + timeout
            while True:
                try:
                    self._sock = socket.create_connection((addr, int(port)), timeout=timeout)
                    break
                except OSError:
                    if time.time() > deadline:
                        raise
                    time.sleep(0.05)
            self._sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            self._send(self._sock, self.rank)

    @classmethod
    def from_env(cls, port_offset=17, **kw):
        import os
        return cls(os.environ.get("RANK", "0"), os.environ.get("WORLD_SIZE", "1"),
                   os.environ.get("MASTER_ADDR", "127.0.0.1"), int(os.environ.get("MASTER_PORT", "29538")) + port_offset, **kw)

    @staticmethod
    def _send(sock, obj):
        import pickle
        import struct
        data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
        sock.sendall(struct.pack("<Q", len(data)) + data)

    @staticmethod
    def _recv(sock):
        import pickle
        import struct

        def exactly(n):
            buf = bytearray()
            while len(buf) < n:
                part = sock.recv(min(1 << 20, n - len(buf)))
                if not part:
                    raise ConnectionError("peer closed the connection")
                buf.extend(part)
            return bytes(buf)

        (n,) = struct.unpack("<Q", exactly(8))
        return pickle.loads(exactly(n))

    def allgather(self, obj):
        if self.world == 1:
            return [obj]
        if self.rank == 0:
            out = [obj] + [self._recv(c) for c in self._peers]
            for c in self._peers:
                self._send(c, out)
            return out
        self._send(self._sock, obj)
        return self._recv(self._sock)

    def barrier(self):
        self.allgather(None)

    def sum(self, a):
        parts = self.allgather(np.array(a, dtype=np.float64, copy=True))
        tot = parts[0].copy()
        for p in parts[1:]:      # rank order: identical on every rank
            tot = tot + p
        return tot

    def max(self, a):
        return np.max(np.stack(self.allgather(np.array(a, dtype=np.float64, copy=True))), axis=0)

    def close(self):
        for c in self._peers:
            c.close()
        if self._sock is not None:
            self._sock.close()
        self._peers, self._sock = [], None


class SingleComm:
    rank, world = 0, 1

    def allgather(self, obj):
        return [obj]

    def barrier(self):
        pass

    def sum(self, a):
        return np.array(a, dtype=np.float64, copy=True)

    max = sum


# ------------------------------------------------------------------------------------------------ the object

class DistBTFem:
    """btfem.BTFem for ONE mesh spread over comm.world GPUs.  Every rank passes the same global arrays
    (vertices in a locality-preserving order, e.g. meshes.rcm_order); per-cell coefficients are given per
    GLOBAL cell and sliced here.  solve() returns the global signal / voi on every rank."""

    def __init__(self, xyz, tets, comm, device=0, phase=None, bounds=None):
        self.comm = comm
        self.rank, self.world = comm.rank, comm.world
        xyz = np.asarray(xyz, dtype=np.float64)
        tets = np.asarray(tets, dtype=np.int32)
        self.nv_global, self.nc_global = len(xyz), len(tets)
        self.bounds = block_bounds(len(xyz), self.world, tets) if bounds is None else np.asarray(bounds)
        self.part = local_part(tets, self.bounds, self.rank)
        self.fem = _bt.BTFem(device)
        ph = None if phase is None else np.asarray(phase, dtype=np.int32)[self.part.cells]
        self.two_comp = phase is not None
        self.fem.set_mesh(xyz[self.part.l2g], self.part.tets, ph)
        self.fem.set_partition(self.part.nv_own, self.part.nv_int)
        self._global = (xyz, tets, None if phase is None else np.asarray(phase, dtype=np.int32))
        self.pdir = None

    def close(self):
        self.fem.close()

    def _cells(self, a):
        return np.ascontiguousarray(np.asarray(a)[self.part.cells])

    def set_diffusion(self, D):
        D = np.asarray(D, dtype=np.float64)
        self.fem.set_diffusion(D if D.ndim == 0 or (D.ndim == 2 and D.shape == (3, 3)) else self._cells(D))

    def set_relaxation(self, inv_t2):
        a = np.asarray(inv_t2, dtype=np.float64)
        self.fem.set_relaxation(a if a.ndim == 0 else self._cells(a))

    def set_permeability(self, kappa, marker=None):
        self.fem.set_permeability(kappa, None if marker is None else self._cells(marker))

    def set_periodic(self, pdir, kappa_e, tol, lo, hi):
        """Weak pseudo-periodic BC; lo/hi: bounding box of the WHOLE mesh (MyDomain prints it, DmriFemLib.py:593)."""
        self.pdir = [int(p) for p in pdir]
        self.lo, self.hi = np.asarray(lo, dtype=float), np.asarray(hi, dtype=float)
        self.fem.set_periodic(self.pdir, kappa_e, tol, self.lo, self.hi)

    def mesh_stats(self):
        """hmin / hmax of the whole mesh (cells are spread over the ranks; a cut cell is seen by several)."""
        lo, hi = self.fem.mesh_stats()
        return (-float(self.comm.max([-lo])[0]), float(self.comm.max([hi])[0]))

    def set_initial(self, ic=None):
        self.fem.set_initial(None if ic is None else np.asarray(ic, dtype=np.float64)[self.part.l2g])

    def assemble(self):
        fem, part = self.fem, self.part
        fem.assemble()
        self.n_own, self.n_int, self.halo_shift = fem.partition_sizes()
        self.dof_vertex, self.dof_comp = fem.dofmap()
        req = halo_requests(part, self.bounds, self.dof_vertex, self.dof_comp, self.n_own, self.halo_shift)
        req_u = np.zeros((0, 4), dtype=np.int64)
        if self.pdir is not None and sum(self.pdir) > 0:
            xyz, tets, phase = self._global
            gdv, gdc, gvc2dof = global_dofmap(len(xyz), tets, phase)
            gather = _periodic.build_gather(xyz, tets, phase, self.pdir, self.lo, self.hi, gdv, gdc)
            local_gather, req_u = periodic_plan(part, self.bounds, gather, gvc2dof, gdv, gdc, self.dof_vertex,
                                                self.dof_comp)
            fem.set_periodic_gather(*local_gather)       # before the export: it sizes the source buffer
        blob = fem.dist_export()
        gathered = self.comm.allgather((req, blob, req_u))
        src, dst_rank, dst_slot, recv_from = send_list(part, [g[0] for g in gathered], self.dof_vertex,
                                                       self.dof_comp, self.n_own)
        usrc, urank, uidx, urecv = send_list(part, [g[2] for g in gathered], self.dof_vertex, self.dof_comp,
                                             self.n_own)
        fem.dist_connect(self.rank, self.world, np.stack([g[1] for g in gathered]), src, dst_rank, dst_slot,
                         np.maximum(recv_from, urecv), u_only=(usrc, urank, uidx))
        self.n_send, self.n_send_u = len(src), len(usrc)
        self.ndof_global = int(self.comm.sum([self.n_own])[0])
        self.comm.barrier()      # every slab is allocated, zeroed and mapped before the first halo push

    def solve(self, *args, **kw):
        res = self.fem.solve(*args, **kw)
        loc = np.array([res["voi"], res["voi_comp"][0], res["voi_comp"][1], res["whole_vol"]])
        tot = self.comm.sum(loc)
        res["voi"], res["whole_vol"] = float(tot[0]), float(tot[3])
        res["voi_comp"] = (float(tot[1]), float(tot[2]))
        return res

    def owned_solution(self):
        """(global vertex, compartment, u) of the owned dofs."""
        u = self.fem.solution()[:self.n_own]
        return self.part.l2g[self.dof_vertex[:self.n_own]], self.dof_comp[:self.n_own].copy(), u

    def global_solution(self):
        """Solution on all (global vertex, compartment) dofs, vertex-major like a single handle numbers them."""
        pieces = self.comm.allgather(self.owned_solution())
        gv = np.concatenate([p[0] for p in pieces])
        cp = np.concatenate([p[1] for p in pieces])
        u = np.concatenate([p[2] for p in pieces])
        order = np.lexsort((cp, gv))
        return gv[order], cp[order], u[order]
