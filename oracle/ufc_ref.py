"""TEST INFRASTRUCTURE ONLY (oracle/): ctypes access to oracle/_ref/libufcref.so, i.e. the
reference's OWN FFC-generated element kernels (built by oracle/build_ref.py from
/root/reference/comri/*/hpc-fenics-cpp/ufc/*.cpp), plus helpers that scatter their element
tensors into the reference's blocked global layout so they can be compared with the oracle.

Coefficient (`w`) orders, from the reference's call sites:
  one-comp NoTime3D cell   : GX, K, mmk, smk, jmk          comri/one-comp/hpc-fenics-cpp/main.cpp:267
  one-comp Bloch_Torrey3D a: GX, ft(TH), gnorm, K, theta, dt   (form 0 of Bloch_Torrey3D.ufl)
  one-comp Bloch_Torrey3D L: u(TH), GX, ft(TH), gnorm, K, theta, dt   main.cpp:239
  two-comp NoTime3D        : phase(DG0), GX, K, kappa, mmk, smk, jmk, imk, kappa_e, h   two-comp main.cpp:832
  Comp_Sig3D               : u(TH)  |  phase(DG0), u(TH)
Local dof numbering of the mixed P1 element is blocked: local = field*4 + vertex; the interior
facet macro tensor is [cell0's 16, cell1's 16].
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_ref", "libufcref.so")
WSTRIDE = 32

_lib = None


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(SO)
        assert _lib.ufcref_wstride() == WSTRIDE
    return _lib


def _pack(ws):
    flat = np.zeros((len(ws), WSTRIDE))
    for i, w in enumerate(ws):
        w = np.atleast_1d(np.asarray(w, dtype=float))
        flat[i, :len(w)] = w
    return flat


_dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))


def cell(name, nA, ws, x):
    A = np.zeros(nA)
    flat = _pack(ws)
    x = np.ascontiguousarray(x, dtype=float)
    getattr(lib(), "ufcref_" + name)(_dp(A), _dp(flat), C.c_int(len(ws)), _dp(x))
    return A


def ext_facet(name, nA, ws, x, facet):
    A = np.zeros(nA)
    flat = _pack(ws)
    x = np.ascontiguousarray(x, dtype=float)
    getattr(lib(), "ufcref_" + name)(_dp(A), _dp(flat), C.c_int(len(ws)), _dp(x), C.c_int(facet))
    return A


def int_facet(name, nA, ws, x0, x1, f0, f1):
    A = np.zeros(nA)
    flat = _pack(ws)
    x0 = np.ascontiguousarray(x0, dtype=float)
    x1 = np.ascontiguousarray(x1, dtype=float)
    getattr(lib(), "ufcref_" + name)(_dp(A), _dp(flat), C.c_int(len(ws)), _dp(x0), _dp(x1), C.c_int(f0), C.c_int(f1))
    return A


OC_NOTIME = "oc_notime_bloch_torrey_notime3d_cell_integral_0_0"
OC_A = "oc_bt_bloch_torrey3d_cell_integral_0_0"
OC_L = "oc_bt_bloch_torrey3d_cell_integral_1_0"
OC_SIG = "oc_sig_comp_sig3d_cell_integral_0_0"
TC_CELL = "tc_notime_bloch_torrey_notime3d_cell_integral_0_0"
TC_EXT = "tc_notime_bloch_torrey_notime3d_exterior_facet_integral_0_0"
TC_INT = "tc_notime_bloch_torrey_notime3d_interior_facet_integral_0_0"
TC_SIG = "tc_sig_comp_sig3d_cell_integral_0_0"


def random_tet(rng, scale=1.0):
    while True:
        x = rng.standard_normal((4, 3)) * scale + rng.standard_normal(3) * 3 * scale
        J = (x[1:] - x[0]).T
        if abs(np.linalg.det(J)) > 0.05 * scale ** 3:
            return x


def golden_element_tensors(seed=2024, ncell=24):
    """Run the reference kernels on seeded inputs; returns dict of arrays (inputs + outputs)."""
    rng = np.random.default_rng(seed)
    out = {}
    X = np.stack([random_tet(rng) for _ in range(ncell)])
    Kc = rng.uniform(1e-3, 3e-3, ncell)
    one = np.ones(4)
    out["x"] = X
    out["K"] = Kc
    # one-comp no-time: mass only, stiffness only, J only (GX = x.g with g random)
    g = rng.standard_normal((ncell, 3))
    out["g"] = g
    A_m, A_s, A_j = [], [], []
    for c in range(ncell):
        GX = X[c] @ g[c]
        A_m.append(cell(OC_NOTIME, 64, [GX, Kc[c] * one, one, 0 * one, 0 * one], X[c]))
        A_s.append(cell(OC_NOTIME, 64, [GX, Kc[c] * one, 0 * one, one, 0 * one], X[c]))
        A_j.append(cell(OC_NOTIME, 64, [GX, Kc[c] * one, 0 * one, 0 * one, one], X[c]))
    out["oc_mass"], out["oc_stiff"], out["oc_j"] = np.array(A_m), np.array(A_s), np.array(A_j)
    # one-comp theta-scheme bilinear + linear forms
    U = rng.standard_normal((ncell, 8))
    ft, gn, th, dt = rng.uniform(-1, 1, ncell), rng.uniform(1e-5, 1e-4, ncell), 0.5, 200.0
    out["u"], out["ft"], out["gnorm"], out["theta"], out["dt"] = U, ft, gn, th, dt
    A_a, b_L, sig = [], [], []
    for c in range(ncell):
        GX = X[c] @ g[c]
        ftw = np.concatenate([ft[c] * one, 0 * one])
        A_a.append(cell(OC_A, 64, [GX, ftw, gn[c] * one, Kc[c] * one, th * one, dt * one], X[c]))
        b_L.append(cell(OC_L, 8, [U[c], GX, ftw, gn[c] * one, Kc[c] * one, th * one, dt * one], X[c]))
        sig.append(cell(OC_SIG, 1, [U[c]], X[c]))
    out["oc_a"], out["oc_L"], out["oc_sig"] = np.array(A_a), np.array(b_L), np.array(sig)[:, 0]
    # two-comp cell integrals for phase 0 and 1
    tc0, tc1 = [], []
    for c in range(ncell):
        GX = X[c] @ g[c]
        ws = lambda ph: [[ph], GX, Kc[c] * one, 0 * one, one, one, one, 0 * one, 0 * one, one]
        tc0.append(cell(TC_CELL, 256, ws(0.0), X[c]))
        tc1.append(cell(TC_CELL, 256, ws(1.0), X[c]))
    out["tc_cell_ph0"], out["tc_cell_ph1"] = np.array(tc0), np.array(tc1)
    # two-comp interior facet: pairs of UFC-ordered tets sharing the facet (v0,v1,v2)|...
    P5 = []
    IF = []
    kap = rng.uniform(1e-5, 1e-4, ncell)
    perms = []
    for c in range(ncell):
        x = random_tet(rng)
        # fifth point: mirror of one vertex through the opposite face, jittered
        opp = rng.integers(0, 4)
        others = [i for i in range(4) if i != opp]
        cen = x[others].mean(axis=0)
        p5 = cen + (cen - x[opp]) * rng.uniform(0.5, 1.5) + 0.1 * rng.standard_normal(3)
        pts = np.vstack([x, p5])                        # cell A = {0,1,2,3}, cell B = others + {4}
        perm = rng.permutation(5)                        # random global numbering
        gid = perm                                       # point i has global id gid[i]
        cellA = sorted([0, 1, 2, 3], key=lambda i: gid[i])
        cellB = sorted(others + [4], key=lambda i: gid[i])
        f0 = cellA.index(opp)
        f1 = cellB.index(4)
        # phase: cell A = 0, cell B = 1 ; coefficient restrictions are [cell0 dofs, cell1 dofs]
        def both(fA, fB):
            return np.concatenate([fA, fB])
        z4 = np.zeros(4)
        ws = [both([0.0], [1.0]), both(z4, z4), both(z4, z4), both(kap[c] * one, kap[c] * one),
              both(z4, z4), both(z4, z4), both(z4, z4), both(one, one), both(z4, z4), both(one, one)]
        A = int_facet(TC_INT, 32 * 32, ws, pts[cellA], pts[cellB], f0, f1)
        P5.append(pts)
        perms.append(np.array([gid[i] for i in range(5)]))
        IF.append(np.concatenate([[opp, f0, f1], cellA, cellB, A]))
    out["if_points"] = np.array(P5)
    out["if_gid"] = np.array(perms)
    out["if_kappa"] = kap
    out["if_record"] = np.array(IF)       # opp, f0, f1, cellA[4], cellB[4], A[1024]
    # two-comp exterior facet with kappa_e/h = P1 field: ws kappa_e nodal values, h = 1
    EX = []
    ke = rng.uniform(0, 1, (ncell, 4))
    out["ext_kappa_e"] = ke
    for c in range(ncell):
        recs = []
        for ph in (0.0, 1.0):
            for facet in range(4):
                ws = [[ph], 0 * one, 0 * one, 0 * one, 0 * one, 0 * one, 0 * one, 0 * one, ke[c], one]
                recs.append(ext_facet(TC_EXT, 256, ws, X[c], facet))
        EX.append(np.array(recs))
    out["tc_ext"] = np.array(EX)          # (ncell, 8 = phase*4+facet, 256)
    return out
