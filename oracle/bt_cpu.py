"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes wrapper of oracle/bt_cpu.c (the C/OpenMP
restatement of the reference's time loop).  Built by __graft_entry__.build_oracle()."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_build", "libbtcpu.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        _lib.btcpu_num_threads.restype = C.c_int
        _lib.btcpu_theta_loop.restype = C.c_int
    return _lib


def num_threads():
    return lib().btcpu_num_threads()


def set_threads(n):
    """Use n OpenMP threads (torchrun exports OMP_NUM_THREADS=1; the CPU arm wants every host core)."""
    lib().btcpu_set_threads(C.c_int(int(n)))


def theta_loop(ops, gdir, k, theta, cA, cb, rtol=1e-9, atol=1e-10, maxit=100000, mode=0, ic=None):
    """Run len(cA) theta steps from `ic` (default 1).  Returns u (complex), iters (per step)."""
    g = np.asarray(gdir, dtype=float)
    g = g / np.linalg.norm(g)
    Jg = (g[0] * ops.Jx.data + g[1] * ops.Jy.data + g[2] * ops.Jz.data)
    K0 = ops.S.data + ops.R.data + ops.I.data
    n = ops.ndof
    u = np.zeros(n, dtype=np.complex128)
    u[:] = 1.0 if ic is None else ic
    cA = np.ascontiguousarray(cA, dtype=float)
    cb = np.ascontiguousarray(cb, dtype=float)
    iters = np.zeros(len(cA), dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rp = np.ascontiguousarray(ops.rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(ops.colidx, dtype=np.int32)
    M = np.ascontiguousarray(ops.M.data)
    B = np.ascontiguousarray(ops.B.data)
    rc = lib().btcpu_theta_loop(C.c_int(n), ip(rp), ip(ci), dp(M), dp(K0), dp(B), dp(Jg), C.c_double(k),
                                C.c_double(theta), C.c_int(len(cA)), dp(cA), dp(cb), C.c_double(rtol),
                                C.c_double(atol), C.c_int(maxit), C.c_int(mode), dp(u), ip(iters))
    if rc != 0:
        raise RuntimeError("bt_cpu: Krylov failure %d" % rc)
    return u, iters


def theta_loop_reassemble(ops, xyz, tets, gdir, k, theta, cA, cb, D, invT2, kappa, rtol=1e-9, atol=1e-10,
                          maxit=100000, ic=None):
    """DmriFemLib.solve work pattern (re-assemble A and b every step).  Scalar D, 1/T2, kappa."""
    import bt_oracle as orc
    g = np.asarray(gdir, dtype=float)
    g = np.ascontiguousarray(g / np.linalg.norm(g))
    n = ops.ndof
    u = np.zeros(n, dtype=np.complex128)
    u[:] = 1.0 if ic is None else ic
    cA = np.ascontiguousarray(cA, dtype=float)
    cb = np.ascontiguousarray(cb, dtype=float)
    iters = np.zeros(len(cA), dtype=np.int32)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    rp = np.ascontiguousarray(ops.rowptr, dtype=np.int32)
    ci = np.ascontiguousarray(ops.colidx, dtype=np.int32)
    xyz = np.ascontiguousarray(xyz, dtype=float)
    tets = np.ascontiguousarray(tets, dtype=np.int32)
    cd = np.ascontiguousarray(ops.cell_dofs, dtype=np.int32)
    if ops.phase is not None and len(ops.iface[0]):
        fv = ops.iface[0]
        ifd = np.ascontiguousarray(np.concatenate([ops.vc2dof[fv, 0], ops.vc2dof[fv, 1]], axis=1), dtype=np.int32)
        coef = np.ascontiguousarray(kappa * orc.tri_area(xyz, fv) / 12.0)
    else:
        ifd = np.zeros((0, 6), dtype=np.int32)
        coef = np.zeros(0)
    f = lib().btcpu_theta_loop_reassemble
    f.restype = C.c_int
    rc = f(C.c_int(n), ip(rp), ip(ci), C.c_int(len(tets)), dp(xyz), ip(tets), ip(cd), C.c_double(D), C.c_double(invT2),
           C.c_int(len(ifd)), ip(ifd), dp(coef), dp(g), C.c_double(k), C.c_double(theta), C.c_int(len(cA)), dp(cA),
           dp(cb), C.c_double(rtol), C.c_double(atol), C.c_int(maxit), dp(u), ip(iters))
    if rc != 0:
        raise RuntimeError("bt_cpu: Krylov failure %d" % rc)
    return u, iters
