"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy/scipy) of the reference's
Bloch-Torrey theta-scheme path.  Nothing in the product path may import this module;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm.

What it restates (all citations are files under /root/reference):

* weak forms              DmriFemLib.py:41-50 (FuncF_wBC, icondition_wBC),
                          :58-145 (ThetaMethodL/F_wBC1c/2c), :240-254 (mass, ident_zeros)
* time loop               DmriFemLib.py:878-915 (MRI_simulation.solve)
* signal                  DmriFemLib.py:917-931, 970-971
* sequence scalars        DmriFemLib.py:826-858, GCloudDmriSolver.py:188-193
* weak pseudo-periodic BC DmriFemLib.py:256-324, 386-450, 599-610
* pre-assembled variant   comri/one-comp/fenics-python/Theta_solver_BT_one_comp.py:162-236,
                          comri/one-comp/hpc-fenics-cpp/main.cpp:263-328

Third-party arithmetic that is NOT under /root/reference and is restated from its
published algorithm: DOLFIN 2019.1.0 `assemble` (P1 element integrals, exact for these
polynomial integrands -> closed forms below, cross-checked against the reference's own
FFC-generated `tabulate_tensor` via oracle/_ref, see tests/test_oracle_vs_ufc.py) and
PETSc 3.7.7 `KSPSolve_BCGS` + `PCJACOBI` + `KSPConvergedDefault` (left-preconditioned
BiCGStab, real arithmetic on the re/im-split system).

Parity pinning: element level against oracle/_ref (reference code, built here); whole
path against the recorded `ConvergenceTest.ipynb cell 10` signal 8.440078e-01 and the
analytic value 0.84389487095614 (tests/test_oracle_golden.py).

Discrete layout used here (and by the CUDA library): the unknowns are the ACTIVE
(vertex, compartment) pairs, one complex number each, numbered vertex-major.  The
reference instead carries 2 (1c) or 4 (2c) real fields on every vertex and pins the
inactive ones with ident_zeros (DmriFemLib.py:246); `expand_to_reference_layout`
converts.  On that numbering
    A_n = P + i*theta*c_n*Jg,     P = M/k + theta*(S + R + I + B)
    b_n = (Q - i*(1-theta)*c_{n-1}*Jg) u^n + (1-theta)*B*u_bc,
    Q  = M/k - (1-theta)*(S + R + I)
with c_n = q*f(t_n) and all of M,S,R,Jx,Jy,Jz,I,B real on one shared CSR pattern.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

# --------------------------------------------------------------------------- geometry


def tet_geometry(xyz, tets):
    """Signed 6*volume, |T| and P1 gradients (nc,4,3) of each tet."""
    x = xyz[tets]                                   # (nc,4,3)
    J = np.stack([x[:, 1] - x[:, 0], x[:, 2] - x[:, 0], x[:, 3] - x[:, 0]], axis=2)  # columns = edges
    det = np.linalg.det(J)
    Jinv = np.linalg.inv(J)                         # rows = grad of barycentric 1..3
    g = np.empty((len(tets), 4, 3))
    g[:, 1:, :] = Jinv
    g[:, 0, :] = -Jinv.sum(axis=1)
    return det, np.abs(det) / 6.0, g


def tri_geometry(xyz, tris):
    """Area and P1 gradients (nc,3,3) of triangles embedded in R^3 (gdim 2 meshes carry z = 0; manifolds keep
    theirs): with e1 = x1-x0, e2 = x2-x0, n = e1 x e2 the gradients are the in-plane vectors
    grad l1 = (e2 x n)/|n|^2, grad l2 = (n x e1)/|n|^2, grad l0 = -(grad l1 + grad l2)."""
    x = xyz[tris]
    e1, e2 = x[:, 1] - x[:, 0], x[:, 2] - x[:, 0]
    n = np.cross(e1, e2)
    n2 = (n * n).sum(axis=1)
    g = np.empty((len(tris), 3, 3))
    g[:, 1] = np.cross(e2, n) / n2[:, None]
    g[:, 2] = np.cross(n, e1) / n2[:, None]
    g[:, 0] = -(g[:, 1] + g[:, 2])
    return 0.5 * np.sqrt(n2), g


def seg_geometry(xyz, segs):
    """Length and P1 gradients (nc,2,3) of segments embedded in R^3 (neuron skeletons, Manifolds.ipynb: tdim 1,
    gdim 3): grad l1 = e/|e|^2, grad l0 = -grad l1."""
    e = xyz[segs[:, 1]] - xyz[segs[:, 0]]
    l2 = (e * e).sum(axis=1)
    g = np.empty((len(segs), 2, 3))
    g[:, 1] = e / l2[:, None]
    g[:, 0] = -g[:, 1]
    return np.sqrt(l2), g


def as_xyz3(xyz):
    """Coordinates as (nv,3): gdim-2 meshes get z = 0 (GdotX then reduces to x*g0 + y*g1, DmriFemLib.py:33-39)."""
    xyz = np.asarray(xyz, dtype=float)
    if xyz.shape[1] == 2:
        xyz = np.hstack([xyz, np.zeros((len(xyz), 1))])
    return xyz


def element_matrices(xyz, tets, D=1.0, invT2=0.0):
    """Closed-form P1 element matrices.

    D: scalar, (nc,) per-cell scalar, or (nc,3,3) / (3,3) tensor (DmriFemLib.py:611-616).
    invT2: scalar or (nc,) per-cell 1/T2 (DmriFemLib.py:43-44).
    Returns dict of (nc,n,n) arrays M,S,R,Jx,Jy,Jz and vol (nc,), n = vertices per cell: 4 (tetrahedra) or
    3 (triangles: 2-D meshes and surfaces in 3-D, DmriFemLib.py:34-38,591-592) or 2 (segments: curves in 3-D,
    Manifolds.ipynb).  On a d-simplex
    int phi_i phi_j = |T|(1+d_ij)/((d+1)(d+2)) and int x phi_i phi_j = |T| w_ij/((d+1)(d+2)(d+3)),
    w_ij = sum_k x_k + x_i + x_j (i != j), 2 sum_k x_k + 4 x_i (i == j).
    """
    nc = len(tets)
    nvc = np.asarray(tets).shape[1]
    d = nvc - 1
    if nvc == 4:
        _, vol, g = tet_geometry(xyz, tets)
    elif nvc == 3:
        vol, g = tri_geometry(xyz, tets)
    else:
        vol, g = seg_geometry(xyz, tets)
    I4 = np.eye(nvc)
    M = vol[:, None, None] * (1.0 + I4)[None] / float((d + 1) * (d + 2))
    D = np.asarray(D, dtype=float)
    if D.ndim == 0:
        Dg = D * g
    elif D.ndim == 1:
        Dg = D[:, None, None] * g
    elif D.ndim == 2:
        Dg = np.einsum("ab,ckb->cka", D, g)
    else:
        Dg = np.einsum("cab,ckb->cka", D, g)
    S = vol[:, None, None] * np.einsum("cia,cja->cij", g, Dg)
    it2 = np.broadcast_to(np.asarray(invT2, dtype=float), (nc,))
    R = it2[:, None, None] * M
    x = xyz[tets]
    out = {"M": M, "S": S, "R": R, "vol": vol}
    jden = float((d + 1) * (d + 2) * (d + 3))
    for d, name in enumerate(("Jx", "Jy", "Jz")):
        xd = x[:, :, d]                              # (nc,n)
        sx = xd.sum(axis=1)
        Jm = sx[:, None, None] + xd[:, :, None] + xd[:, None, :]          # i != j
        Jd = 2.0 * sx[:, None] + 4.0 * xd                                  # i == j
        Jm = Jm * (1 - I4)[None] + Jd[:, :, None] * I4[None]
        out[name] = vol[:, None, None] * Jm / jden
    return out


# --------------------------------------------------------------------------- dof map / pattern


def dof_map(nv, tets, phase=None, vmaster=None):
    """Number the active (vertex, compartment) pairs vertex-major.

    phase: None (one compartment) or (nc,) int in {0,1} = marker % 2 (DmriFemLib.py:764).
    vmaster: None, or (nv,) the master vertex of every vertex (itself if it is not a periodic slave): strongly
    imposed periodicity (`constrained_domain=PeriodicBD`, DmriFemLib.py:327-375, 478-483) -- a slave vertex
    carries the dofs of its master, only masters are numbered.
    Returns cell_dofs (nc,4) int32, ndof, dof_vertex (ndof,), dof_comp (ndof,),
    vc2dof (nv,2) int32 (-1 where inactive).
    """
    tets = np.asarray(tets)
    if phase is None:
        phase = np.zeros(len(tets), dtype=np.int32)
    phase = np.asarray(phase).astype(np.int32)
    vm = np.arange(nv) if vmaster is None else np.asarray(vmaster)
    active = np.zeros((nv, 2), dtype=bool)
    for c in (0, 1):
        active[vm[np.unique(tets[phase == c])], c] = True
    flat = active.ravel()
    ids = np.cumsum(flat) - 1
    vc2dof = np.where(flat, ids, -1).reshape(nv, 2).astype(np.int32)
    ndof = int(flat.sum())
    dv, dc = np.nonzero(active)
    vc2dof = vc2dof[vm]                       # slaves point at their master's dofs
    cell_dofs = vc2dof[tets, phase[:, None]].astype(np.int32)
    return cell_dofs, ndof, dv.astype(np.int32), dc.astype(np.int32), vc2dof


_FACES = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])  # facet i is opposite vertex i (UFC)


_EDGES = np.array([[1, 2], [0, 2], [0, 1]])


def facets(tets):
    """All (cell, local facet) sorted by vertex tuple.  Returns key (nf,3) (tets) or (nf,2) (triangles), cell, lf."""
    nc = len(tets)
    nvc = np.asarray(tets).shape[1]
    loc = _FACES if nvc == 4 else _EDGES
    f = np.sort(tets[:, loc], axis=2).reshape(nc * nvc, nvc - 1)
    cell = np.repeat(np.arange(nc), nvc)
    lf = np.tile(np.arange(nvc), nc)
    order = np.lexsort(tuple(f[:, k] for k in range(nvc - 2, -1, -1)))
    return f[order], cell[order], lf[order]


def interface_facets(tets, phase):
    """Interior facets whose two cells have different phase (|jump(phase)| = 1).

    Returns verts (ni,3), cell0 (phase 0 side), cell1 (phase 1 side)."""
    f, cell, _ = facets(tets)
    same = np.all(f[1:] == f[:-1], axis=1)
    i0 = np.nonzero(same)[0]
    ca, cb = cell[i0], cell[i0 + 1]
    keep = phase[ca] != phase[cb]
    ca, cb, fv = ca[keep], cb[keep], f[i0][keep]
    swap = phase[ca] == 1
    c0 = np.where(swap, cb, ca)
    c1 = np.where(swap, ca, cb)
    return fv, c0, c1


def boundary_facets(tets):
    """Exterior facets: verts (nb,3), owning cell (nb,)."""
    f, cell, _ = facets(tets)
    n = len(f)
    same_next = np.zeros(n, dtype=bool)
    same_next[:-1] = np.all(f[1:] == f[:-1], axis=1)
    same_prev = np.zeros(n, dtype=bool)
    same_prev[1:] = same_next[:-1]
    ext = ~(same_next | same_prev)
    return f[ext], cell[ext]


def tri_area(xyz, fv):
    """Measure of facets: triangle area (fv (n,3)) or edge length (fv (n,2))."""
    if fv.shape[1] == 2:
        return np.linalg.norm(xyz[fv[:, 1]] - xyz[fv[:, 0]], axis=1)
    a = xyz[fv[:, 1]] - xyz[fv[:, 0]]
    b = xyz[fv[:, 2]] - xyz[fv[:, 0]]
    return 0.5 * np.linalg.norm(np.cross(a, b), axis=1)


def scalar_pattern(nv, tets):
    """The reference's scalar P1 sparsity in mesh-vertex numbering: row v couples to every
    w sharing a cell with v, diagonal included, columns sorted (UFC dofmap
    comri/one-comp/hpc-fenics-cpp/ufc/Bloch_Torrey3D.cpp:3869-3884 + PETSc sorted AIJ)."""
    n = np.asarray(tets).shape[1]
    r = np.repeat(tets, n, axis=1).ravel()
    c = np.tile(tets, (1, n)).ravel()
    A = sp.coo_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(nv, nv)).tocsr()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32)


def _assemble(n, rows, cols, vals):
    A = sp.coo_matrix((vals, (rows, cols)), shape=(n, n)).tocsr()
    A.sort_indices()
    return A


class Operators:
    """All real matrices of the path on the active-dof numbering, sharing one pattern."""


def assemble(xyz, tets, phase=None, D=1.0, invT2=0.0, kappa=0.0, kappa_facet=None,
             bnd_kappa_vertex=None, vmaster=None):
    """Assemble M,S,R,Jx,Jy,Jz,(I),(B) + lumped mass on the active-dof numbering.

    kappa: scalar membrane permeability (`-p`); kappa_facet: optional callable
    (fv, c0, c1)->(ni,) or array aligned with interface_facets() for variable permeability.
    bnd_kappa_vertex: optional (nv,) vertex values of the P1-interpolated artificial
    permeability marker kappa_e^h (DmriFemLib.py:601-610) -> boundary matrix B.
    """
    xyz = as_xyz3(xyz)
    tets = np.asarray(tets)
    nv = len(xyz)
    nvc = tets.shape[1]                 # 4: tetrahedra, 3: triangles, 2: segments
    nvf = nvc - 1                       # vertices per facet
    if nvc == 2 and (phase is not None or bnd_kappa_vertex is not None):
        raise ValueError("segment meshes: one compartment, Neumann ends")
    cell_dofs, ndof, dv, dc, vc2dof = dof_map(nv, tets, phase, vmaster)
    em = element_matrices(xyz, tets, D, invT2)
    rows = np.repeat(cell_dofs, nvc, axis=1).ravel()
    cols = np.tile(cell_dofs, (1, nvc)).ravel()
    ops = Operators()
    ops.ndof, ops.nv, ops.cell_dofs, ops.dof_vertex, ops.dof_comp, ops.vc2dof = ndof, nv, cell_dofs, dv, dc, vc2dof
    ops.phase = None if phase is None else np.asarray(phase).astype(np.int32)
    ops.xyz, ops.cells, ops.Dcoef = xyz, tets, D
    trip = {k: (rows, cols, em[k].ravel()) for k in ("M", "S", "R", "Jx", "Jy", "Jz")}
    # interface (DmriFemLib.py:47-50, 112, 139): kappa*(u0-u1)(v0-v1) on facets with |jump(phase)|=1
    if phase is not None:
        fv, c0, c1 = interface_facets(tets, ops.phase)
        if len(fv):
            if kappa_facet is None:
                kf = np.full(len(fv), float(kappa))
            elif callable(kappa_facet):
                kf = np.asarray(kappa_facet(fv, c0, c1), dtype=float)
            else:
                kf = np.asarray(kappa_facet, dtype=float)
            area = tri_area(xyz, fv)
            fm = (kf * area)[:, None, None] * (1.0 + np.eye(nvf))[None] / float(nvf * (nvf + 1))   # facet mass
            d0 = vc2dof[fv, 0]
            d1 = vc2dof[fv, 1]
            assert (d0 >= 0).all() and (d1 >= 0).all()
            rr, cc, vv = [], [], []
            for (ra, ca, sgn) in ((d0, d0, 1.0), (d1, d1, 1.0), (d0, d1, -1.0), (d1, d0, -1.0)):
                rr.append(np.repeat(ra, nvf, axis=1).ravel())
                cc.append(np.tile(ca, (1, nvf)).ravel())
                vv.append(sgn * fm.ravel())
            trip["I"] = (np.concatenate(rr), np.concatenate(cc), np.concatenate(vv))
        ops.iface = (fv, c0, c1)
    # boundary mass with P1-interpolated kappa_e (DmriFemLib.py:92, 143, 601-610; Appendix A.6)
    if bnd_kappa_vertex is not None:
        bf, bcell = boundary_facets(tets)
        kv = np.asarray(bnd_kappa_vertex, dtype=float)[bf]                 # (nb,3)
        keep = np.any(kv != 0.0, axis=1)
        bf, bcell, kv = bf[keep], bcell[keep], kv[keep]
        area = tri_area(xyz, bf)
        # int phi_i phi_j phi_k = |F| * {1/10, 1/30, 1/60} (triangle), |F| * {1/4, 1/12} (edge)
        sk = kv.sum(axis=1)
        Bm = (sk[:, None, None] + kv[:, :, None] + kv[:, None, :])          # i != j
        Bd = 2.0 * sk[:, None] + 4.0 * kv                                   # i == j
        I3 = np.eye(nvf)
        Bm = (Bm * (1 - I3)[None] + Bd[:, :, None] * I3[None]) * area[:, None, None] / float(nvf * (nvf + 1) * (nvf + 2))
        ph = np.zeros(len(tets), dtype=np.int32) if phase is None else ops.phase
        bd = vc2dof[bf, ph[bcell][:, None]]
        trip["B"] = (np.repeat(bd, nvf, axis=1).ravel(), np.tile(bd, (1, nvf)).ravel(), Bm.ravel())
        ops.bnd = (bf, bcell)
    # shared pattern = union of every structural entry (cell couplings + interface couplings)
    rr = np.concatenate([trip[k][0] for k in ("M", "I") if k in trip]).astype(np.int64)
    cc = np.concatenate([trip[k][1] for k in ("M", "I") if k in trip]).astype(np.int64)
    keys = np.unique(rr * ndof + cc)
    prow = (keys // ndof).astype(np.int64)
    ops.colidx = (keys % ndof).astype(np.int32)
    ops.rowptr = np.concatenate([[0], np.cumsum(np.bincount(prow, minlength=ndof))]).astype(np.int32)
    ops.nnz = len(keys)
    for name in ("M", "S", "R", "Jx", "Jy", "Jz", "I", "B"):
        data = np.zeros(ops.nnz)
        if name in trip:
            r_, c_, v_ = trip[name]
            idx = np.searchsorted(keys, r_.astype(np.int64) * ndof + c_)
            assert np.array_equal(keys[idx], r_.astype(np.int64) * ndof + c_)
            data = np.bincount(idx, weights=v_, minlength=ops.nnz)
        setattr(ops, name, sp.csr_matrix((data, ops.colidx, ops.rowptr), shape=(ndof, ndof)))
    ops.lumped = np.asarray(ops.M.sum(axis=1)).ravel()       # 1^T M  -> signal weights
    return ops


def expand_to_reference_layout(ops, u):
    """Active-dof complex vector -> the reference's blocked real layout
    (u0r,u0i[,u1r,u1i]) with dof = comp*N_vert + vertex
    (comri/two-comp/hpc-fenics-cpp/ufc/Bloch_Torrey_NoTime3D.cpp:6372-6400); inactive = 0."""
    ncomp = 1 if ops.phase is None else 2
    out = np.zeros((2 * ncomp, ops.nv))
    out[2 * ops.dof_comp, ops.dof_vertex] = u.real
    out[2 * ops.dof_comp + 1, ops.dof_vertex] = u.imag
    return out.ravel()


# --------------------------------------------------------------------------- sequences


class Sequence:
    """f(s), F(s)=int_0^s f, int_0^T F^2 evaluated exactly for piecewise profiles.

    Mirrors MRI_parameters (DmriFemLib.py:800-858): `fs_sym` is a sympy Piecewise in `s`,
    itime_profile_sym integrates it symbolically, integral_term_for_gb integrates F^2 over
    [0,T], convert_b2q gives q = sqrt(b)/sqrt(int F^2).  sympy is what the reference uses,
    so the oracle uses it too (the product driver has its own copy of this logic)."""

    def __init__(self, fs_sym, T, s=None):
        import sympy
        self.sympy = sympy
        self.s = s if s is not None else sympy.Symbol("s")
        self.fs_sym = fs_sym
        self.T = T
        u = sympy.Symbol("u")
        self.ifs_sym = sympy.integrate(fs_sym.subs(self.s, u), (u, 0, self.s))
        self.int4gb = float(sympy.integrate(self.ifs_sym * self.ifs_sym, (self.s, 0, T)))

    def f(self, t):
        return float(self.fs_sym.subs(self.s, t))

    def F(self, t):
        return float(self.ifs_sym.subs(self.s, t))

    def q_from_b(self, b):
        return np.sqrt(b) / np.sqrt(self.int4gb)


def pgse(delta, Delta):
    """GCloudDmriSolver.py:187-193 (strict `<`: f(delta)=0, f(Delta)=-1, f(T)=0)."""
    import sympy
    s = sympy.Symbol("s")
    T = Delta + delta
    fs = sympy.Piecewise((1.0, s < delta), (0.0, s < Delta), (-1.0, s < T), (0.0, True))
    return Sequence(fs, T, s)


def q2g(q):
    return q / 2.675e8 * 1e12          # DmriFemLib.py:684-686


def time_grid(T, k, closed=True):
    """t_n of the loop `while t < T + k` (DmriFemLib.py:897-910), t accumulated in floating
    point exactly like the reference.  closed=False gives the `t < T` loop of
    ConvergenceTest.ipynb cell 10 / comri multilayer main.cpp:301."""
    ts = []
    t = 0.0
    lim = T + k if closed else T
    while t < lim:
        ts.append(t)
        t += k
    return np.array(ts)


# --------------------------------------------------------------------------- Krylov (PETSc restated)


class _Pc:
    """Left preconditioner K^-1 of the Krylov restatements, used as `Kinv * vector`: None (identity), the real Jacobi
    diagonal (array), or any callable vector -> vector (ILU(0): ilu0_factor)."""

    def __init__(self, spec):
        self.spec = spec

    def __mul__(self, v):
        if self.spec is None:
            return v
        if callable(self.spec):
            return self.spec(v)
        return v / self.spec


def ilu0_factor(A):
    """PETSc PCILU with its defaults restated (third party, PETSc 3.7: MatILUFactorSymbolic/Numeric, levels = 0, natural
    ordering, no shift): incomplete LU on the pattern of A, row by row (IKJ), no pivoting; L has a unit diagonal and is
    stored with U in one array like PETSc's factored AIJ matrix.  A: complex CSR with sorted columns, i.e. the
    reference's real (re,im)-split matrix with every 2x2 block [[a,-b],[b,a]] written as the complex number a + ib --
    scalar ILU(0) on the split matrix drops no fill inside the (full) blocks, so both factorisations apply the same
    operator when unknowns are ordered vertex by vertex.  Returns K^-1: v -> U^-1 L^-1 v."""
    A = sp.csr_matrix(A).astype(complex)
    A.sort_indices()
    n = A.shape[0]
    rp, ci, lu = A.indptr, A.indices, A.data.copy()
    diag = np.full(n, -1, dtype=np.int64)
    for i in range(n):
        pos = {int(ci[k]): k for k in range(rp[i], rp[i + 1])}
        for k in range(rp[i], rp[i + 1]):
            c = int(ci[k])
            if c >= i:
                break
            lu[k] = lu[k] / lu[diag[c]]                    # l_ic
            lic = lu[k]
            for kk in range(diag[c] + 1, rp[c + 1]):       # row i -= l_ic * (U part of row c), inside the pattern only
                j = pos.get(int(ci[kk]))
                if j is not None:
                    lu[j] -= lic * lu[kk]
        diag[i] = pos[i]
    L = sp.csr_matrix((lu, ci, rp), shape=(n, n))
    Lo = (sp.tril(L, -1) + sp.identity(n, dtype=complex)).tocsr()
    Up = sp.triu(L, 0).tocsr()

    def apply(v):
        y = spla.spsolve_triangular(Lo, np.asarray(v, dtype=complex), lower=True, unit_diagonal=True)
        return spla.spsolve_triangular(Up, y, lower=False)

    apply.lu = lu
    return apply


def bicgstab_petsc(matvec, b, diag, rtol=1e-9, atol=1e-10, maxit=100000, x0=None, dtol=1e4):
    """PETSc 3.7 KSPSolve_BCGS with left PCJACOBI and the default convergence test
    (preconditioned residual 2-norm <= max(rtol*||K^-1 b||, atol)), on COMPLEX storage but
    with REAL inner products Re(a^H b): identical to running it on the reference's
    real (re,im)-split system.  `diag` is the real Jacobi diagonal P_ii.
    Returns x, iterations, final preconditioned residual norm, reason (>0 converged)."""
    rdot = lambda a, c: float(np.dot(a.real, c.real) + np.dot(a.imag, c.imag))
    Kinv = _Pc(diag)
    n = len(b)
    if x0 is None:
        x = np.zeros(n, dtype=complex)
        r = Kinv * b
        bnorm = np.sqrt(rdot(r, r))
    else:
        x = x0.astype(complex).copy()
        r = Kinv * (b - matvec(x))
        kb = Kinv * b
        bnorm = np.sqrt(rdot(kb, kb))
    dp = np.sqrt(rdot(r, r))
    ttol = max(rtol * bnorm, atol)
    if dp <= ttol:
        return x, 0, dp, 2 if dp <= atol and not dp <= rtol * bnorm else 2
    rp = r.copy()
    rhoold = alpha = omegaold = 1.0
    p = np.zeros(n, dtype=complex)
    v = np.zeros(n, dtype=complex)
    i = 0
    while i < maxit:
        rho = rdot(r, rp)
        beta = (rho / rhoold) * (alpha / omegaold)
        p = r - (omegaold * beta) * v + beta * p
        v = Kinv * matvec(p)
        d1 = rdot(v, rp)
        if d1 == 0.0:
            return x, i, dp, -5
        alpha = rho / d1
        s = r - alpha * v
        t = Kinv * matvec(s)
        d1 = rdot(s, t)
        d2 = rdot(t, t)
        if d2 == 0.0:
            if rdot(s, s) != 0.0:
                return x, i, dp, -5
            x = x + alpha * p
            return x, i + 1, 0.0, 2
        omega = d1 / d2
        x = x + alpha * p + omega * s
        r = s - omega * t
        dp = np.sqrt(rdot(r, r))
        rhoold, omegaold = rho, omega
        i += 1
        if not np.isfinite(dp):
            return x, i, dp, -9
        if dp <= ttol:
            return x, i, dp, 2
        if dp >= dtol * bnorm:
            return x, i, dp, -4
        if rho == 0.0:
            return x, i, dp, -5
    return x, i, dp, -3


def gmres_petsc(matvec, b, diag, rtol=1e-6, atol=1e-15, maxit=10000, restart=30, x0=None, dtol=1e4):
    """PETSc 3.7 KSPGMRES restated: restart 30, classical Gram-Schmidt without refinement, left PCJACOBI
    (diag=None: no preconditioner), Givens rotations, default convergence test on the preconditioned
    residual estimate.  REAL inner products on complex storage == GMRES on the reference's real
    (re,im)-split system.  Returns x, iterations, residual norm, reason."""
    rdot = lambda a, c: float(np.dot(a.real, c.real) + np.dot(a.imag, c.imag))
    Kinv = _Pc(diag)
    n = len(b)
    kb = Kinv * b
    bnorm = np.sqrt(rdot(kb, kb))
    if x0 is None:
        x = np.zeros(n, dtype=complex)
        r = kb.copy()
    else:
        x = x0.astype(complex).copy()
        r = Kinv * (b - matvec(x))
    res = np.sqrt(rdot(r, r))
    ttol = max(rtol * bnorm, atol)
    if res <= ttol:
        return x, 0, res, 2
    its = 0
    m = restart
    while True:
        V = [r / res]
        H = np.zeros((m + 1, m))
        cs, sn, rs = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        rs[0] = res
        reason = 0
        k = 0
        for j in range(m):
            w = Kinv * matvec(V[j])
            hcol = np.array([rdot(w, V[i]) for i in range(j + 1)])       # classical GS: all dots first
            for i in range(j + 1):
                w = w - hcol[i] * V[i]
            tt = np.sqrt(rdot(w, w))
            H[:j + 1, j] = hcol
            H[j + 1, j] = tt
            for i in range(j):
                t0 = H[i, j]
                H[i, j] = cs[i] * t0 + sn[i] * H[i + 1, j]
                H[i + 1, j] = -sn[i] * t0 + cs[i] * H[i + 1, j]
            den = np.hypot(H[j, j], H[j + 1, j])
            if den == 0.0:
                reason = -5
                break
            cs[j], sn[j] = H[j, j] / den, H[j + 1, j] / den
            rs[j + 1] = -sn[j] * rs[j]
            rs[j] = cs[j] * rs[j]
            H[j, j] = cs[j] * H[j, j] + sn[j] * H[j + 1, j]
            res = abs(rs[j + 1])
            its += 1
            k = j + 1
            if not np.isfinite(res):
                reason = -9
            elif res <= ttol:
                reason = 2
            elif res >= dtol * bnorm:
                reason = -4
            elif its >= maxit:
                reason = -3
            elif tt == 0.0:
                reason = -5
            if reason:
                break
            V.append(w / tt)
        y = np.zeros(k)
        for i in range(k - 1, -1, -1):
            y[i] = (rs[i] - H[i, i + 1:k] @ y[i + 1:k]) / H[i, i]
        for i in range(k):
            x = x + y[i] * V[i]
        if reason:
            return x, its, res, reason
        r = kb - Kinv * matvec(x)
        res = np.sqrt(rdot(r, r))
        if res <= ttol:
            return x, its, res, 2


# --------------------------------------------------------------------------- weak pseudo-periodic


def domain_sizes(xyz, tets):
    """bbox, hmin, hmax.  h of a cell = longest edge (DOLFIN >= 2017 `Cell::h`, third party;
    SURVEY Appendix C.17)."""
    x = xyz[tets]
    n = np.asarray(tets).shape[1]
    e = [np.linalg.norm(x[:, i] - x[:, j], axis=1) for i in range(n) for j in range(i + 1, n)]
    h = np.max(np.stack(e, axis=1), axis=1)
    return xyz.min(axis=0), xyz.max(axis=0), float(h.min()), float(h.max())


def periodic_marker(xyz, pdir, lo, hi, hmin):
    """Vertex values of kappa_e^h = (3e-3/hmin) * 1[vertex within 1e-2*hmin of a periodic
    face] (DmriFemLib.py:590, 599-610; the C `||` of products makes it 0/1)."""
    eps = 1e-2 * hmin
    m = np.zeros(len(xyz), dtype=bool)
    for d in range(3):
        if pdir[d]:
            m |= (xyz[:, d] < lo[d] + eps) | (xyz[:, d] > hi[d] - eps)
    return (3e-3 / hmin) * m.astype(float)


def periodic_term(xyz, tets, ops, pdir, lo, hi, q, gdir, theta):
    """Returns the callable (u, F_prev) -> (1-theta) * B * u_bc of ThetaMethodL_wBC*c
    (DmriFemLib.py:71-76, 115-121) with u_bc from WeakPseudoPeriodic_*.eval (:270-321, 401-448):
    P1 evaluation of the previous solution at the mirrored point (brute-force search over the
    boundary triangles of the opposite face; closest triangle if none contains it, as
    allow_extrapolation does), rotated by exp(i q (g.(x'-x)) F(t_p)); last matching direction wins."""
    g = np.asarray(gdir, dtype=float)
    g = g / np.linalg.norm(g)
    bf, _ = boundary_facets(np.asarray(tets))
    nv = len(xyz)
    planar = bf.shape[1] == 2          # triangle mesh in the x-y plane: boundary facets are edges
    rows = []          # (dof, [(src dof or -1, weight)]*3, g.dx)
    for v in range(nv):
        hit = None
        for d in range(2 if planar else 3):
            if pdir[d]:
                if abs(xyz[v, d] - lo[d]) <= 1e-7:
                    hit = (d, hi[d])
                if abs(xyz[v, d] - hi[d]) <= 1e-7:
                    hit = (d, lo[d])
        if hit is None:
            continue
        d, target = hit
        other = [a for a in range(2 if planar else 3) if a != d]
        tri = bf[np.all(np.abs(xyz[bf][:, :, d] - target) <= 1e-7, axis=1)]
        p = xyz[v, other]
        best, bw = None, None
        for t in (tri if planar else ()):          # P1 along the opposite boundary edge
            a, b = xyz[t][:, other[0]]
            w0 = (p[0] - b) / (a - b)
            w = np.array([w0, 1.0 - w0])
            if best is None or w.min() > bw.min():
                best, bw = t, w
        for t in (() if planar else tri):
            a, b, c = xyz[t][:, other]
            den = (b[1] - c[1]) * (a[0] - c[0]) + (c[0] - b[0]) * (a[1] - c[1])
            w0 = ((b[1] - c[1]) * (p[0] - c[0]) + (c[0] - b[0]) * (p[1] - c[1])) / den
            w1 = ((c[1] - a[1]) * (p[0] - c[0]) + (a[0] - c[0]) * (p[1] - c[1])) / den
            w = np.array([w0, w1, 1 - w0 - w1])
            if best is None or w.min() > bw.min():
                best, bw = t, w
        gdx = g[d] * (target - xyz[v, d])
        for comp in (0, 1):
            dof = ops.vc2dof[v, comp]
            if dof >= 0:
                rows.append((dof, [(ops.vc2dof[t, comp], wt) for t, wt in zip(best, bw)], gdx))

    def term(u, F_prev):
        ubc = np.zeros(ops.ndof, dtype=complex)
        for dof, srcs, gdx in rows:
            val = sum(wt * u[s] for s, wt in srcs if s >= 0)
            ubc[dof] = val * np.exp(1j * q * gdx * F_prev)
        return (1.0 - theta) * (ops.B @ ubc)

    return term


# --------------------------------------------------------------------------- strongly imposed periodicity


def periodic_vertex_map(xyz, pdir, lo, hi, tol):
    """Master vertex of every vertex for `constrained_domain=PeriodicBD` (DmriFemLib.py:327-375): a vertex within
    `tol` of the max face of a periodic direction is identified with the vertex at the same place on the min face
    (PeriodicBD.map shifts by the box length; `inside` = the min faces are the masters).  Edges and corners of the
    box wrap in every periodic direction they touch.  The mesh must be periodic: a slave without a partner raises.
    (How DOLFIN resolves vertices that are slave in one direction and master in another is third party; wrapping
    all directions is the periodic lattice.)"""
    xyz = as_xyz3(xyz)
    w = xyz.copy()
    slave = np.zeros(len(xyz), dtype=bool)
    for d in range(3):
        if pdir[d]:
            on = np.abs(xyz[:, d] - hi[d]) < tol
            w[on, d] = lo[d]
            slave |= on
    key = np.round(w / tol).astype(np.int64)
    order = np.lexsort((key[:, 2], key[:, 1], key[:, 0]))
    ks = key[order]
    first = np.ones(len(ks), dtype=bool)
    first[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    group = np.cumsum(first) - 1
    vm = np.arange(len(xyz))
    # master of a group: its (unique) non-slave member
    gid = np.empty(len(xyz), dtype=np.int64)
    gid[order] = group
    master_of_group = -np.ones(group[-1] + 1, dtype=np.int64)
    ns = np.nonzero(~slave)[0]
    if len(np.unique(gid[ns])) != len(ns):
        raise ValueError("two distinct non-slave vertices coincide after wrapping: tol too large")
    master_of_group[gid[ns]] = ns
    vm = master_of_group[gid]
    if (vm < 0).any():
        raise ValueError("the mesh is not periodic: %d vertices on a max face have no partner on the min face"
                         % int((vm < 0).sum()))
    return vm


def strong_operators(ops, gdir):
    """The matrices the transformed (strongly periodic) equation adds (FuncF_sBC, outer_interface,
    inner_interface, DmriFemLib.py:147-165), on the pattern of `ops` (built with the same vmaster):

        W[i,j] = int (g.Dg) phi_i phi_j                       (from -q^2 F^2 (g.Dg) u v)
        C[i,j] = int (g.D grad phi_j + grad phi_j . D g) phi_i (from -i q F (g.D grad u + grad u.Dg) v)
        N[i,j] = sum over the facets bounding the compartment of the cell (exterior facets and interface facets,
                 seen from either side) of (D g . n_out) int_facet phi_i phi_j
                 (outer_interface: -i (qF + 1e-16)(Dg.n) u v ds; inner_interface reduces to the same expression
                 per side once D0 = D1 = D is inserted, as every call site does; the 1e-16 guard is dropped)

    so that with Phi = F(t):  F_s(Phi; u, v) = -[S + R + q^2 Phi^2 W] u v - i q Phi [C - N] u v (+ kappa jump terms).
    Returns (W, G) with G = C - N, csr on ops' pattern."""
    g = np.asarray(gdir, dtype=float)
    g = g / np.linalg.norm(g)
    xyz, cells = ops.xyz, np.asarray(ops.cells)
    nc, nvc = cells.shape
    d = nvc - 1
    if nvc == 4:
        _, vol, grad = tet_geometry(xyz, cells)
    elif nvc == 3:
        vol, grad = tri_geometry(xyz, cells)
    else:
        raise ValueError("strong periodic BC: tetrahedra or triangles")
    D = np.asarray(ops.Dcoef, dtype=float)
    if D.ndim == 0:
        Dc = np.broadcast_to(D * np.eye(3), (nc, 3, 3))
    elif D.ndim == 1:
        Dc = D[:, None, None] * np.eye(3)[None]
    elif D.ndim == 2:
        Dc = np.broadcast_to(D, (nc, 3, 3))
    else:
        Dc = D
    Dg = np.einsum("cab,b->ca", Dc, g)                    # D g
    DTg = np.einsum("cba,b->ca", Dc, g)                   # D^T g
    gDg = np.einsum("ca,a->c", Dg, g)
    I = np.eye(nvc)
    Wm = (gDg * vol)[:, None, None] * (1.0 + I)[None] / float((d + 1) * (d + 2))
    # C[i,j] = |T|/(d+1) * ((D + D^T) g) . grad phi_j      (int phi_i = |T|/(d+1))
    cj = np.einsum("ca,cja->cj", Dg + DTg, grad)
    Cm = (vol / (d + 1.0))[:, None, None] * np.broadcast_to(cj[:, None, :], (nc, nvc, nvc))
    rows = np.repeat(ops.cell_dofs, nvc, axis=1).ravel()
    cols = np.tile(ops.cell_dofs, (1, nvc)).ravel()
    n = ops.ndof
    W = sp.coo_matrix((Wm.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    C = sp.coo_matrix((Cm.ravel(), (rows, cols)), shape=(n, n)).tocsr()
    # facets bounding a compartment: exterior facets, and interface facets from both sides.  For the facet of
    # cell T opposite its local vertex o:  n_out |F| = -d |T| grad(lambda_o), so
    # (Dg.n_out) int_F phi_i phi_j = -d |T| (Dg.grad lambda_o) (1 + d_ij) / (d (d+1)) = -|T| (Dg.grad lambda_o)(1+d_ij)/(d+1)
    f, cell, lf = facets(cells)
    nf = len(f)
    same_next = np.zeros(nf, dtype=bool)
    same_next[:-1] = np.all(f[1:] == f[:-1], axis=1)
    same_prev = np.zeros(nf, dtype=bool)
    same_prev[1:] = same_next[:-1]
    ph = np.zeros(nc, dtype=np.int32) if ops.phase is None else ops.phase
    partner = np.where(same_next, np.roll(cell, -1), np.where(same_prev, np.roll(cell, 1), -1))
    bounding = (partner < 0) | (ph[np.maximum(partner, 0)] != ph[cell])
    fc, fl = cell[bounding], lf[bounding]
    loc = (_FACES if nvc == 4 else _EDGES)[fl]                                   # local vertices of the facet
    coef = -vol[fc] * np.einsum("fa,fa->f", Dg[fc], grad[fc, fl]) / (d + 1.0)
    nvf = nvc - 1
    Nm = coef[:, None, None] * (1.0 + np.eye(nvf))[None]
    fd = ops.cell_dofs[fc[:, None], loc]                                         # dofs of the cell's own compartment
    N = sp.coo_matrix((Nm.ravel(), (np.repeat(fd, nvf, axis=1).ravel(), np.tile(fd, (1, nvf)).ravel())),
                      shape=(n, n)).tocsr()
    return W, (C - N).tocsr()


def theta_solve_strong(ops, seq, q, gdir, k, theta=0.5, closed=True, ic=None):
    """MRI_simulation.solve with ThetaMethodF/L_sBC1c/2c (DmriFemLib.py:166-238, 878-915): the transformed equation on
    a periodic function space.  The matrix uses F(t_n), the right-hand side F(t_{n-1}) (ift_f / ift_p_f, :901-902),
    and BOTH carry theta (the linear form is written with `+theta*FuncF_sBC`, :183, :232-233 -- equal to (1-theta)
    for the theta = 0.5 the class fixes).  Exact (sparse LU) stepping.
    Returns dict(u, signal, voi, n_steps)."""
    W, G = strong_operators(ops, gdir)
    K0 = ops.S + ops.R + ops.I
    ic = np.ones(ops.ndof) if ic is None else np.asarray(ic, dtype=float)
    u = ic.astype(complex)
    ts = time_grid(seq.T, k, closed)
    tp = 0.0
    lus = {}
    for t in ts:
        Fn, Fp = seq.F(t), seq.F(tp)
        b = (ops.M / k - theta * (K0 + (q * Fp) ** 2 * W)) @ u - 1j * theta * q * Fp * (G @ u)
        key = round(Fn, 300)
        if key not in lus:
            lus[key] = spla.splu((ops.M / k + theta * (K0 + (q * Fn) ** 2 * W) + 1j * theta * q * Fn * G).tocsc())
        u = lus[key].solve(b)
        tp = t
    return dict(u=u, signal=float(ops.lumped @ u.real), voi=float(ops.lumped @ ic), n_steps=len(ts))


# --------------------------------------------------------------------------- the theta loop


def theta_solve(ops, seq, q, gdir, k, theta=0.5, solver="lu", rtol=1e-9, atol=1e-10, maxit=100000,
                closed=True, ic=None, nonzero_guess=False, rhs_uses_current_f=False,
                periodic=None, return_history=False, restart=30):
    """MRI_simulation.solve (DmriFemLib.py:878-915) on the pre-assembled operators.

    A_n uses f(t_n); b_n uses f(t_{n-1}), t_{-1}=0 (DmriFemLib.py:901-902, 909); the comri
    C++ solvers use f(t_n) on both sides (one-comp/hpc-fenics-cpp/main.cpp:300-301)
    -> rhs_uses_current_f=True.  solver: "lu" (exact discrete solve; stands for MUMPS) or
    "bicgstab" (PETSc restatement with Jacobi).  periodic: None or a callable
    (u, F_prev) -> (1-theta)*B*u_bc contribution vector.
    Returns dict(u, signal, voi, iters, n_steps)."""
    g = np.asarray(gdir, dtype=float)
    g = g / np.linalg.norm(g)                                   # DmriFemLib.py:819-821
    Jg = (g[0] * ops.Jx + g[1] * ops.Jy + g[2] * ops.Jz).tocsr()
    K0 = ops.S + ops.R + ops.I
    P = (ops.M / k + theta * (K0 + ops.B)).tocsr()
    Q = (ops.M / k - (1.0 - theta) * K0).tocsr()
    diag = P.diagonal()
    ic = np.ones(ops.ndof) if ic is None else np.asarray(ic, dtype=float)
    u = ic.astype(complex)
    ts = time_grid(seq.T, k, closed)
    lus = {}
    iters = []
    hist = []
    tp = 0.0
    for t in ts:
        cA = q * seq.f(t)
        cb = cA if rhs_uses_current_f else q * seq.f(tp)
        b = Q @ u - 1j * (1.0 - theta) * cb * (Jg @ u)
        if periodic is not None:
            b = b + periodic(u, seq.F(tp))
        if solver == "lu":
            key = round(cA, 300)
            if key not in lus:
                lus[key] = spla.splu((P + 1j * theta * cA * Jg).tocsc())
            u = lus[key].solve(b)
            iters.append(0)
        else:
            mv = lambda x, cA=cA: P @ x + 1j * theta * cA * (Jg @ x)
            pc = diag
            if solver.endswith("_none"):
                pc = None
            elif solver.endswith("_ilu"):                   # KrylovSolver("gmres", "ilu"), fenics-cpp/main.cpp:180-183
                key = round(cA, 300)
                if key not in lus:
                    lus[key] = ilu0_factor(P + 1j * theta * cA * Jg)
                pc = lus[key]
            if solver.startswith("gmres"):
                u, it, dp, reason = gmres_petsc(mv, b, pc, rtol, atol, maxit, restart, x0=u if nonzero_guess else None)
            else:
                u, it, dp, reason = bicgstab_petsc(mv, b, pc, rtol, atol, maxit, x0=u if nonzero_guess else None)
            if reason < 0:
                raise RuntimeError("Krylov solver did not converge: reason %d" % reason)
            iters.append(it)
        if return_history:
            hist.append(float(ops.lumped @ u.real))
        tp = t
    out = dict(u=u, signal=float(ops.lumped @ u.real), voi=float(ops.lumped @ ic),
               iters=np.array(iters), n_steps=len(ts))
    if return_history:
        out["history"] = np.array(hist)
    return out
