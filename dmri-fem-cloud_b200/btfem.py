"""ctypes binding of libbtfem.so (include/btfem.h) -- the only way the Python host layer
reaches the GPU.  There is no CPU fallback: if the shared library is missing or no CUDA
device is usable, construction raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbtfem.so")

MAT_IDS = {"M": 0, "S": 1, "R": 2, "Jx": 3, "Jy": 4, "Jz": 5, "I": 6, "B": 7}
DIST_BLOB_BYTES = 192   # BTFEM_DIST_BLOB_BYTES
KSP_IDS = {"bicgstab": 0, "gmres": 1}
PC_IDS = {"jacobi": 0, "none": 1, "ilu": 2}

_c_double_p = C.POINTER(C.c_double)
_c_int32_p = C.POINTER(C.c_int32)
_c_int64_p = C.POINTER(C.c_int64)


class SolveArgs(C.Structure):
    _fields_ = [("nsteps", C.c_int64), ("dt", C.c_double), ("theta", C.c_double),
                ("cA", _c_double_p), ("cb", _c_double_p), ("Fb", _c_double_p),
                ("gdir", C.c_double * 3), ("q", C.c_double),
                ("ksp", C.c_int64), ("pc", C.c_int64), ("rtol", C.c_double), ("atol", C.c_double),
                ("maxit", C.c_int64), ("nonzero_guess", C.c_int64), ("restart", C.c_int64)]


class SolveOut(C.Structure):
    _fields_ = [("signal", C.c_double), ("signal_comp", C.c_double * 2), ("voi", C.c_double),
                ("voi_comp", C.c_double * 2), ("whole_vol", C.c_double), ("loop_ms", C.c_double),
                ("setup_ms", C.c_double), ("total_iters", C.c_int64), ("max_iters", C.c_int64),
                ("n_spmv", C.c_int64), ("n_kernels", C.c_int64), ("last_reason", C.c_int64)]


class BTFemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbtfem error %d: %s" % (code, msg))
        self.code = code


_lib = None


def load_library(path=None):
    """Load libbtfem.so and declare every prototype of include/btfem.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("libbtfem.so not found at %s -- build it with `python __graft_entry__.py` "
                           "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    H = C.c_void_p
    proto = {
        "btfem_create": (C.c_int, [C.c_int, C.POINTER(H)]),
        "btfem_destroy": (None, [H]),
        "btfem_last_error": (C.c_char_p, [H]),
        "btfem_version": (C.c_int, []),
        "btfem_set_mesh": (C.c_int, [H, C.c_int64, _c_double_p, C.c_int64, _c_int32_p, _c_int32_p]),
        "btfem_set_mesh_tri": (C.c_int, [H, C.c_int64, _c_double_p, C.c_int64, _c_int32_p, _c_int32_p]),
        "btfem_set_mesh_seg": (C.c_int, [H, C.c_int64, _c_double_p, C.c_int64, _c_int32_p]),
        "btfem_set_periodic_map": (C.c_int, [H, _c_int32_p]),
        "btfem_get_strong_operators": (C.c_int, [H, _c_double_p, _c_double_p, _c_double_p]),
        "btfem_set_phase": (C.c_int, [H, _c_int32_p]),
        "btfem_get_mesh_stats": (C.c_int, [H, _c_double_p, _c_double_p]),
        "btfem_get_bbox": (C.c_int, [H, _c_double_p, _c_double_p]),
        "btfem_set_diffusion": (C.c_int, [H, C.c_int, _c_double_p]),
        "btfem_set_relaxation": (C.c_int, [H, C.c_int, _c_double_p]),
        "btfem_set_permeability": (C.c_int, [H, C.c_int, _c_double_p, C.c_int32, _c_int32_p]),
        "btfem_set_periodic": (C.c_int, [H, _c_int32_p, C.c_double, C.c_double, _c_double_p, _c_double_p]),
        "btfem_get_boundary_facets": (C.c_int, [H, _c_int32_p]),
        "btfem_set_periodic_gather": (C.c_int, [H, C.c_int64, _c_int32_p, _c_int32_p, _c_double_p, _c_double_p]),
        "btfem_set_initial": (C.c_int, [H, _c_double_p]),
        "btfem_assemble": (C.c_int, [H]),
        "btfem_get_sizes": (C.c_int, [H, _c_int64_p, _c_int64_p, _c_int64_p, _c_int64_p]),
        "btfem_get_pattern": (C.c_int, [H, _c_int32_p, _c_int32_p]),
        "btfem_get_dofmap": (C.c_int, [H, _c_int32_p, _c_int32_p]),
        "btfem_get_values": (C.c_int, [H, C.c_int, _c_double_p]),
        "btfem_get_lumped_mass": (C.c_int, [H, _c_double_p]),
        "btfem_spmv": (C.c_int, [H, C.c_double, C.c_double, C.c_double, _c_double_p, _c_double_p, _c_double_p]),
        "btfem_spmv_bench": (C.c_int, [H, C.c_double, C.c_double, C.c_double, _c_double_p, C.c_int32, C.c_int32,
                                       C.c_int32, _c_double_p]),
        "btfem_set_lanes": (C.c_int, [H, C.c_int32]),
        "btfem_get_spmv_kernel": (C.c_int, [H, _c_int32_p]),
        "btfem_get_ilu_factors": (C.c_int, [H, _c_double_p]),
        "btfem_set_sm_partition": (C.c_int, [H, C.c_int32]),
        "btfem_solve": (C.c_int, [H, C.POINTER(SolveArgs), C.POINTER(SolveOut), _c_int32_p]),
        "btfem_solve_batch": (C.c_int, [H, C.c_int32, C.POINTER(SolveArgs), C.POINTER(SolveOut)]),
        "btfem_get_solution": (C.c_int, [H, _c_double_p]),
        "btfem_set_partition": (C.c_int, [H, C.c_int64, C.c_int64]),
        "btfem_get_partition": (C.c_int, [H, _c_int64_p, _c_int64_p, _c_int64_p]),
        "btfem_dist_export": (C.c_int, [H, C.c_void_p]),
        "btfem_dist_connect": (C.c_int, [H, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, _c_int32_p, _c_int32_p,
                                         _c_int32_p, C.c_int64, _c_int32_p, _c_int32_p, _c_int32_p, _c_int32_p]),
        "btfem_dist_trace": (C.c_int, [H, C.c_int64]),
        "btfem_dist_get_trace": (C.c_int, [H, C.c_void_p, C.c_int64, _c_int64_p]),
    }
    for name, (res, args) in proto.items():
        f = getattr(lib, name)
        f.restype = res
        f.argtypes = args
    lib._btfem_symbols = sorted(proto)
    if path == LIB_PATH:
        _lib = lib
    return lib


def _dp(a):
    return a.ctypes.data_as(_c_double_p)


def _ip(a):
    return a.ctypes.data_as(_c_int32_p)


class BTFem:
    """One problem (mesh + coefficients) on one GPU."""

    def __init__(self, device=0, lib=None):
        self.lib = lib or load_library()
        self.h = C.c_void_p()
        rc = self.lib.btfem_create(int(device), C.byref(self.h))
        if rc != 0:
            raise BTFemError(rc, "btfem_create failed: no usable CUDA device %d (libbtfem has no CPU fallback)" % device)
        self.nv = self.nc = 0
        self.ndof = self.nnz = 0

    def close(self):
        if getattr(self, "h", None) and self.h.value:
            self.lib.btfem_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc != 0:
            raise BTFemError(rc, (self.lib.btfem_last_error(self.h) or b"").decode())

    # ---- problem definition
    def set_mesh(self, xyz, tets, phase=None):
        """tets: (nc,4) tetrahedra, (nc,3) triangles or (nc,2) segments; xyz: (nv,3), or (nv,2) for a gdim-2 mesh."""
        xyz = np.asarray(xyz, dtype=np.float64)
        if xyz.ndim != 2 or xyz.shape[1] not in (2, 3):
            raise ValueError("xyz must be (nv,3) or (nv,2)")
        if xyz.shape[1] == 2:
            xyz = np.hstack([xyz, np.zeros((len(xyz), 1))])
        xyz = np.ascontiguousarray(xyz)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        if tets.ndim != 2 or tets.shape[1] not in (2, 3, 4):
            raise ValueError("cells must be (nc,4) tetrahedra, (nc,3) triangles or (nc,2) segments")
        if tets.shape[1] == 2 and phase is not None:
            raise ValueError("segment meshes are one-compartment")
        ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.int32)
        self.nv, self.nc = len(xyz), len(tets)
        self.two_comp = ph is not None
        self.cell_nv = tets.shape[1]
        if self.cell_nv == 2:
            self._ck(self.lib.btfem_set_mesh_seg(self.h, self.nv, _dp(xyz), self.nc, _ip(tets)))
        else:
            fn = self.lib.btfem_set_mesh if self.cell_nv == 4 else self.lib.btfem_set_mesh_tri
            self._ck(fn(self.h, self.nv, _dp(xyz), self.nc, _ip(tets), None if ph is None else _ip(ph)))
        self.h2d_bytes = xyz.nbytes + tets.nbytes + (0 if ph is None else ph.nbytes)

    def set_phase(self, phase=None):
        ph = None if phase is None else np.ascontiguousarray(phase, dtype=np.int32)
        if ph is not None:
            assert len(ph) == self.nc
        self.two_comp = ph is not None
        self._ck(self.lib.btfem_set_phase(self.h, None if ph is None else _ip(ph)))
        self.h2d_bytes += 0 if ph is None else ph.nbytes

    def mesh_stats(self):
        lo, hi = C.c_double(), C.c_double()
        self._ck(self.lib.btfem_get_mesh_stats(self.h, C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def bbox(self):
        """Bounding box of the vertices (btfem_get_bbox): (lo[3], hi[3])."""
        lo, hi = np.zeros(3), np.zeros(3)
        self._ck(self.lib.btfem_get_bbox(self.h, _dp(lo), _dp(hi)))
        return lo, hi

    def set_diffusion(self, D):
        D = np.asarray(D, dtype=np.float64)
        if D.ndim == 0:
            kind, D = 0, D.reshape(1).copy()
        elif D.ndim == 1:
            kind, D = 1, np.ascontiguousarray(D)
            assert len(D) == self.nc
        else:
            D = np.ascontiguousarray(np.broadcast_to(D, (self.nc, 3, 3)))
            kind = 2
        self._ck(self.lib.btfem_set_diffusion(self.h, kind, _dp(D)))

    def set_relaxation(self, inv_t2):
        a = np.asarray(inv_t2, dtype=np.float64)
        kind = 0 if a.ndim == 0 else 1
        a = a.reshape(1).copy() if kind == 0 else np.ascontiguousarray(a)
        self._ck(self.lib.btfem_set_relaxation(self.h, kind, _dp(a)))

    def set_permeability(self, kappa, marker=None):
        k = np.asarray(kappa, dtype=np.float64)
        if k.ndim == 0:
            k = k.reshape(1).copy()
            self._ck(self.lib.btfem_set_permeability(self.h, 0, _dp(k), 0, None))
        else:
            k = np.ascontiguousarray(k)
            m = np.ascontiguousarray(marker, dtype=np.int32)
            assert k.ndim == 2 and k.shape[0] == k.shape[1]
            self._ck(self.lib.btfem_set_permeability(self.h, 1, _dp(k), k.shape[0], _ip(m)))

    def set_periodic(self, pdir, kappa_e, tol, lo, hi):
        p = np.ascontiguousarray(pdir, dtype=np.int32)
        lo = np.ascontiguousarray(lo, dtype=np.float64)
        hi = np.ascontiguousarray(hi, dtype=np.float64)
        self._ck(self.lib.btfem_set_periodic(self.h, _ip(p), float(kappa_e), float(tol), _dp(lo), _dp(hi)))

    def set_periodic_map(self, vmaster=None):
        """Strongly imposed periodicity: vmaster[v] = master vertex of v (periodic.vertex_map); None switches it off.
        btfem_solve then steps the transformed equation: pass cA = q*F(t_n), cb = q*F(t_{n-1})."""
        if vmaster is None:
            self._ck(self.lib.btfem_set_periodic_map(self.h, None))
        else:
            vm = np.ascontiguousarray(vmaster, dtype=np.int32)
            assert len(vm) == self.nv
            self._ck(self.lib.btfem_set_periodic_map(self.h, _ip(vm)))

    def strong_operators(self, gdir):
        g = np.ascontiguousarray(gdir, dtype=np.float64)
        W, G = np.empty(self.nnz), np.empty(self.nnz)
        self._ck(self.lib.btfem_get_strong_operators(self.h, _dp(g), _dp(W), _dp(G)))
        return W, G

    def boundary_facets(self):
        out = np.empty((self.n_bfacet, 3), dtype=np.int32)
        self._ck(self.lib.btfem_get_boundary_facets(self.h, _ip(out)))
        return out

    def set_periodic_gather(self, dof, src, w, dx):
        dof = np.ascontiguousarray(dof, dtype=np.int32)
        src = np.ascontiguousarray(src, dtype=np.int32)
        w = np.ascontiguousarray(w, dtype=np.float64)
        dx = np.ascontiguousarray(dx, dtype=np.float64)
        self._ck(self.lib.btfem_set_periodic_gather(self.h, len(dof), _ip(dof), _ip(src), _dp(w), _dp(dx)))

    def set_initial(self, ic=None):
        if ic is None:
            self._ck(self.lib.btfem_set_initial(self.h, None))
        else:
            a = np.ascontiguousarray(ic, dtype=np.float64)
            assert len(a) == self.nv
            self._ck(self.lib.btfem_set_initial(self.h, _dp(a)))

    def set_lanes(self, lanes):
        self._ck(self.lib.btfem_set_lanes(self.h, int(lanes)))

    def set_sm_partition(self, nblocks):
        """The persistent kernel of this handle runs on `nblocks` SMs (0 = all): several handles with disjoint shares
        solve concurrently on one GPU (btfem_set_sm_partition).  Call before assemble()."""
        self._ck(self.lib.btfem_set_sm_partition(self.h, int(nblocks)))

    @property
    def spmv_kernel(self):
        """0 CSR / 1 SELL-32 register-staged / 2 SELL-32 through per-warp TMA rings (btfem_get_spmv_kernel)."""
        kind = C.c_int32(0)
        self._ck(self.lib.btfem_get_spmv_kernel(self.h, C.byref(kind)))
        return int(kind.value)

    @property
    def stream_kernel(self):
        return self.spmv_kernel == 2

    def ilu_factors(self):
        """Complex ILU(0) factors of the last pc="ilu" solve, CSR order (btfem_get_ilu_factors)."""
        out = np.zeros(2 * self.nnz)
        self._ck(self.lib.btfem_get_ilu_factors(self.h, _dp(out)))
        return out[0::2] + 1j * out[1::2]

    # ---- assembly + parity hooks
    def assemble(self):
        self._ck(self.lib.btfem_assemble(self.h))
        s = [C.c_int64() for _ in range(4)]
        self._ck(self.lib.btfem_get_sizes(self.h, *[C.byref(x) for x in s]))
        self.ndof, self.nnz, self.n_iface, self.n_bfacet = (int(x.value) for x in s)

    def pattern(self):
        rp = np.empty(self.ndof + 1, dtype=np.int32)
        ci = np.empty(self.nnz, dtype=np.int32)
        self._ck(self.lib.btfem_get_pattern(self.h, _ip(rp), _ip(ci)))
        return rp, ci

    def dofmap(self):
        dv = np.empty(self.ndof, dtype=np.int32)
        dc = np.empty(self.ndof, dtype=np.int32)
        self._ck(self.lib.btfem_get_dofmap(self.h, _ip(dv), _ip(dc)))
        return dv, dc

    def values(self, which):
        out = np.empty(self.nnz, dtype=np.float64)
        self._ck(self.lib.btfem_get_values(self.h, MAT_IDS[which], _dp(out)))
        return out

    def lumped_mass(self):
        out = np.empty(self.ndof, dtype=np.float64)
        self._ck(self.lib.btfem_get_lumped_mass(self.h, _dp(out)))
        return out

    def spmv(self, dt, theta, c, gdir, x):
        x = np.ascontiguousarray(x, dtype=np.complex128)
        assert len(x) == self.ndof
        y = np.empty(self.ndof, dtype=np.complex128)
        g = np.ascontiguousarray(gdir, dtype=np.float64)
        self._ck(self.lib.btfem_spmv(self.h, dt, theta, c, _dp(g), x.ctypes.data_as(_c_double_p),
                                     y.ctypes.data_as(_c_double_p)))
        return y

    def spmv_bench(self, dt, theta, c, gdir, lanes=8, nrep=20, flush_l2=False):
        g = np.ascontiguousarray(gdir, dtype=np.float64)
        ms = C.c_double()
        self._ck(self.lib.btfem_spmv_bench(self.h, dt, theta, c, _dp(g), lanes, nrep, int(flush_l2), C.byref(ms)))
        return ms.value

    # ---- theta loop
    def solve(self, dt, theta, cA, cb, gdir, q=0.0, Fb=None, ksp="bicgstab", pc="jacobi", rtol=1e-9, atol=1e-10,
              maxit=100000, nonzero_guess=False, restart=30, want_iters=False):
        cA = np.ascontiguousarray(cA, dtype=np.float64)
        cb = np.ascontiguousarray(cb, dtype=np.float64)
        assert len(cA) == len(cb)
        a = SolveArgs()
        a.nsteps = len(cA)
        a.dt, a.theta = float(dt), float(theta)
        a.cA, a.cb = _dp(cA), _dp(cb)
        if Fb is not None:
            Fb = np.ascontiguousarray(Fb, dtype=np.float64)
            a.Fb = _dp(Fb)
        g = np.asarray(gdir, dtype=np.float64)
        a.gdir[0], a.gdir[1], a.gdir[2] = g
        a.q = float(q)
        a.ksp, a.pc = KSP_IDS[ksp], PC_IDS[pc]
        a.rtol, a.atol, a.maxit = float(rtol), float(atol), int(maxit)
        a.nonzero_guess, a.restart = int(bool(nonzero_guess)), int(restart)
        o = SolveOut()
        iters = np.zeros(len(cA), dtype=np.int32) if want_iters else None
        self._ck(self.lib.btfem_solve(self.h, C.byref(a), C.byref(o), None if iters is None else _ip(iters)))
        res = {k: getattr(o, k) for k, _ in SolveOut._fields_ if k not in ("signal_comp", "voi_comp")}
        res["signal_comp"] = (o.signal_comp[0], o.signal_comp[1])
        res["voi_comp"] = (o.voi_comp[0], o.voi_comp[1])
        res["n_steps"] = len(cA)
        if want_iters:
            res["iters"] = iters
        return res

    def solve_batch(self, dt, theta, members, ksp="bicgstab", pc="jacobi", rtol=1e-9, atol=1e-10, maxit=100000):
        """members: list of (cA, cb, gdir) -- independent solves advanced in lock step on the GPU."""
        nb = len(members)
        args = (SolveArgs * nb)()
        outs = (SolveOut * nb)()
        keep = []
        for a, (cA, cb, g) in zip(args, members):
            cA = np.ascontiguousarray(cA, dtype=np.float64)
            cb = np.ascontiguousarray(cb, dtype=np.float64)
            keep += [cA, cb]
            a.nsteps = len(cA)
            a.dt, a.theta = float(dt), float(theta)
            a.cA, a.cb = _dp(cA), _dp(cb)
            g = np.asarray(g, dtype=np.float64)
            a.gdir[0], a.gdir[1], a.gdir[2] = g
            a.ksp, a.pc = KSP_IDS[ksp], PC_IDS[pc]
            a.rtol, a.atol, a.maxit = float(rtol), float(atol), int(maxit)
            a.nonzero_guess, a.restart = 0, 30
        self._ck(self.lib.btfem_solve_batch(self.h, nb, args, outs))
        res = []
        for o in outs:
            d = {k: getattr(o, k) for k, _ in SolveOut._fields_ if k not in ("signal_comp", "voi_comp")}
            d["signal_comp"] = (o.signal_comp[0], o.signal_comp[1])
            d["voi_comp"] = (o.voi_comp[0], o.voi_comp[1])
            res.append(d)
        return res

    # ---- one mesh row-partitioned over several GPUs (host logic: partition.py)
    def set_partition(self, nv_own, nv_interior):
        self._ck(self.lib.btfem_set_partition(self.h, int(nv_own), int(nv_interior)))

    def partition_sizes(self):
        """(owned dofs, owned dofs no peer needs, element shift of halo dofs in the vectors)"""
        s = [C.c_int64() for _ in range(3)]
        self._ck(self.lib.btfem_get_partition(self.h, *[C.byref(x) for x in s]))
        return tuple(int(x.value) for x in s)

    def dist_export(self):
        blob = np.zeros(DIST_BLOB_BYTES, dtype=np.uint8)
        self._ck(self.lib.btfem_dist_export(self.h, blob.ctypes.data_as(C.c_void_p)))
        return blob

    def dist_connect(self, rank, world, blobs, src, dst_rank, dst_slot, recv_from, u_only=None):
        """u_only: optional (src, dst_rank, dst_index) of the entries that travel with u only (periodic sources)."""
        blobs = np.ascontiguousarray(blobs, dtype=np.uint8).reshape(world, DIST_BLOB_BYTES)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        src, dst_rank, dst_slot, recv_from = i32(src), i32(dst_rank), i32(dst_slot), i32(recv_from)
        assert len(src) == len(dst_rank) == len(dst_slot) and len(recv_from) == world
        us, ur, ui = (i32(a) for a in (u_only if u_only is not None else ([], [], [])))
        assert len(us) == len(ur) == len(ui)
        self._ck(self.lib.btfem_dist_connect(self.h, int(rank), int(world), blobs.ctypes.data_as(C.c_void_p),
                                             len(src), _ip(src), _ip(dst_rank), _ip(dst_slot), len(us), _ip(us),
                                             _ip(ur), _ip(ui), _ip(recv_from)))

    def dist_trace(self, max_entries):
        self._trace_cap = int(max_entries)
        self._ck(self.lib.btfem_dist_trace(self.h, int(max_entries)))

    def dist_get_trace(self):
        """(n,8) uint64: first block start, local work done, collective done, longest halo wait [ns], kind, 0, 0, 0."""
        out = np.zeros((self._trace_cap, 8), dtype=np.uint64)
        n = C.c_int64()
        self._ck(self.lib.btfem_dist_get_trace(self.h, out.ctypes.data_as(C.c_void_p), self._trace_cap, C.byref(n)))
        return out[:n.value]

    def solution(self):
        u = np.empty(self.ndof, dtype=np.complex128)
        self._ck(self.lib.btfem_get_solution(self.h, u.ctypes.data_as(_c_double_p)))
        return u
