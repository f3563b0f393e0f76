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
