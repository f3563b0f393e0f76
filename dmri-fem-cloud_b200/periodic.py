"""Weak pseudo-periodic boundary condition: the gather operator behind `WeakPseudoPeriodic_1c/_2c`
(DmriFemLib.py:256-324, 386-450).

For a vertex x on the min/max face of a periodic direction the reference evaluates the previous
solution at the mirrored point x' (that coordinate replaced by the opposite face's value) by P1 point
evaluation with extrapolation allowed, and rotates it by exp(i*q*(g.(x'-x))*F(t_p)).  If several
directions match, the LAST one in x, y, z order wins (the assignments overwrite, :276-313).  The
face test is |x - face| <= 1e-7 (:271).

This is mesh bookkeeping done once per mesh on the host: it produces, per boundary dof, up to three
source dofs with barycentric weights and the displacement x'-x.  The per-step arithmetic
(u_bc = phase * sum w u, then (1-theta) * B * u_bc) runs on the GPU (csrc/solve.cu)."""
import numpy as np

FACE_TOL = 1e-7


def _boundary_triangles(tets):
    faces = np.array([[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]])
    f = np.sort(tets[:, faces], axis=2).reshape(-1, 3)
    cell = np.repeat(np.arange(len(tets)), 4)
    order = np.lexsort((f[:, 2], f[:, 1], f[:, 0]))
    f, cell = f[order], cell[order]
    same_next = np.zeros(len(f), dtype=bool)
    same_next[:-1] = np.all(f[1:] == f[:-1], axis=1)
    same_prev = np.zeros(len(f), dtype=bool)
    same_prev[1:] = same_next[:-1]
    ext = ~(same_next | same_prev)
    return f[ext], cell[ext]


def _plane_triangles(xyz, tets, d, target):
    """Facets with all three vertices on the plane x_d = target.  The plane is a face of the bounding box, so
    these facets are exterior: no facet sort over the whole mesh is needed (row-partitioned set-up, where no
    single GPU sees the whole mesh)."""
    on = np.abs(xyz[:, d] - target) <= FACE_TOL
    ont = on[tets]
    cand = np.nonzero(ont.sum(axis=1) >= 3)[0]
    out = []
    for lf in range(4):
        keep = [k for k in range(4) if k != lf]
        m = ont[cand][:, keep].all(axis=1)
        out.append(np.sort(tets[cand[m]][:, keep], axis=1))
    f = np.concatenate(out) if out else np.zeros((0, 3), dtype=np.int64)
    return np.unique(f, axis=0) if len(f) else f.reshape(0, 3)


def _locate(points2, tri_xy, ncand=16):
    """For each 2-D point: containing triangle (or the least-outside one among the `ncand` triangles with the
    nearest centroids: extrapolation, like allow_extrapolation) and its barycentric weights.  Vectorised."""
    from scipy.spatial import cKDTree
    a, b, c = tri_xy[:, 0], tri_xy[:, 1], tri_xy[:, 2]
    d = (b[:, 1] - c[:, 1]) * (a[:, 0] - c[:, 0]) + (c[:, 0] - b[:, 0]) * (a[:, 1] - c[:, 1])
    cen = tri_xy.mean(axis=1)
    k = min(ncand, len(cen))
    _, cand = cKDTree(cen).query(points2, k=k)
    cand = cand.reshape(len(points2), k)
    px, py = points2[:, 0:1], points2[:, 1:2]
    w0 = ((b[cand, 1] - c[cand, 1]) * (px - c[cand, 0]) + (c[cand, 0] - b[cand, 0]) * (py - c[cand, 1])) / d[cand]
    w1 = ((c[cand, 1] - a[cand, 1]) * (px - c[cand, 0]) + (a[cand, 0] - c[cand, 0]) * (py - c[cand, 1])) / d[cand]
    w2 = 1.0 - w0 - w1
    W = np.stack([w0, w1, w2], axis=2)                    # (np, k, 3)
    best = np.argmax(W.min(axis=2), axis=1)               # inside: min weight >= 0; else least outside
    rows = np.arange(len(points2))
    tri, w = cand[rows, best], W[rows, best]
    # On anisotropic / strongly graded faces the containing triangle need not be among the `ncand` nearest centroids:
    # points that ended up outside every candidate are searched again over ALL triangles (exhaustive, few points).
    out = np.nonzero(w.min(axis=1) < -1e-9)[0]
    for i0 in range(0, len(out), 256):
        sel = out[i0:i0 + 256]
        qx, qy = points2[sel, 0:1], points2[sel, 1:2]
        v0 = ((b[:, 1] - c[:, 1]) * (qx - c[:, 0]) + (c[:, 0] - b[:, 0]) * (qy - c[:, 1])) / d
        v1 = ((c[:, 1] - a[:, 1]) * (qx - c[:, 0]) + (a[:, 0] - c[:, 0]) * (qy - c[:, 1])) / d
        V = np.stack([v0, v1, 1.0 - v0 - v1], axis=2)     # (points, triangles, 3)
        bt = np.argmax(V.min(axis=2), axis=1)
        r = np.arange(len(sel))
        better = V[r, bt].min(axis=1) > w[sel].min(axis=1)
        tri[sel[better]] = bt[better]
        w[sel[better]] = V[r, bt][better]
    return tri, w


def _line_segments(xyz, tris, d, target):
    """Triangle mesh in the x-y plane: edges with both vertices on the line x_d = target."""
    on = np.abs(xyz[:, d] - target) <= FACE_TOL
    out = []
    for a, b in ((1, 2), (0, 2), (0, 1)):
        m = on[tris[:, a]] & on[tris[:, b]]
        out.append(np.sort(tris[m][:, [a, b]], axis=1))
    f = np.concatenate(out)
    return np.unique(f, axis=0) if len(f) else f.reshape(0, 2)


def _locate_1d(p, seg_x):
    """For each coordinate p along the opposite boundary line: the containing edge (or the least-outside one:
    extrapolation) and its two P1 weights."""
    a, b = seg_x[:, 0][None, :], seg_x[:, 1][None, :]
    w0 = (p[:, None] - b) / (a - b)
    W = np.stack([w0, 1.0 - w0], axis=2)                  # (np, ns, 2)
    best = np.argmax(W.min(axis=2), axis=1)
    rows = np.arange(len(p))
    return best, W[rows, best]


def vertex_map(xyz, pdir, lo, hi, tol):
    """Strongly imposed periodicity (`constrained_domain = PeriodicBD`, DmriFemLib.py:327-375): the master vertex of
    every vertex.  A vertex within `tol` of the max face of a periodic direction (PeriodicBD.map) is identified with
    the vertex at the same place on the min face (PeriodicBD.inside: the min faces are the masters); vertices on
    edges / corners of the box wrap in every periodic direction they touch.  The mesh has to be periodic: a max-face
    vertex without a partner raises.  (PeriodicBD uses tol = 1e-2*hmin, :334.)"""
    xyz = np.asarray(xyz, dtype=float)
    if xyz.shape[1] == 2:
        xyz = np.hstack([xyz, np.zeros((len(xyz), 1))])
    nv = len(xyz)
    wrapped = xyz.copy()
    slave = np.zeros(nv, dtype=bool)
    for d in range(3):
        if pdir[d]:
            on = np.abs(xyz[:, d] - hi[d]) < tol
            wrapped[on, d] = lo[d]
            slave |= on
    # every wrapped position must coincide (within tol) with a vertex that is on no max face: nearest-neighbour search
    from scipy.spatial import cKDTree
    vm = np.arange(nv, dtype=np.int64)
    masters = np.nonzero(~slave)[0]
    sl = np.nonzero(slave)[0]
    if len(sl):
        if len(masters) == 0:
            raise RuntimeError("periodic map: every vertex lies on a max face")
        dist, idx = cKDTree(xyz[masters]).query(wrapped[sl], k=1)
        if (dist >= tol).any():
            raise RuntimeError("the mesh is not periodic: %d vertices on a max face have no partner on the min face"
                               % int((dist >= tol).sum()))
        vm[sl] = masters[idx]
    return vm.astype(np.int32)


def build_gather(xyz, tets, phase, pdir, lo, hi, dof_vertex, dof_comp, bfacets=None):
    """Returns dof (nb,), src (nb,3) dof ids or -1, w (nb,3), dx (nb,3).

    bfacets: exterior facets touching the periodic faces (btfem_get_boundary_facets) -- found on the GPU during
    assembly; None: search them here.
    dof_vertex/dof_comp: the library's dof map (btfem_get_dofmap).  Field `comp` evaluated at a vertex
    where that compartment is inactive is 0 (the reference's pinned dofs), i.e. src = -1."""
    xyz = np.asarray(xyz, dtype=float)
    tets = np.asarray(tets)
    planar = tets.shape[1] == 3          # triangle mesh in the x-y plane: boundary facets are edges (third vertex -1)
    if xyz.shape[1] == 2:
        xyz = np.hstack([xyz, np.zeros((len(xyz), 1))])
    if planar:
        pdir = [pdir[0], pdir[1], 0]
    nv = len(xyz)
    vc2dof = -np.ones((nv, 2), dtype=np.int64)
    vc2dof[dof_vertex, dof_comp] = np.arange(len(dof_vertex))
    tets = np.asarray(tets)
    bf = np.asarray(bfacets, dtype=np.int64) if bfacets is not None else None
    # per vertex: (direction, side) of the LAST matching periodic direction
    vdir = -np.ones(nv, dtype=np.int64)
    vside = np.zeros(nv, dtype=np.int64)
    for d in range(3):
        if pdir[d]:
            on_lo = np.abs(xyz[:, d] - lo[d]) <= FACE_TOL
            on_hi = np.abs(xyz[:, d] - hi[d]) <= FACE_TOL
            vdir[on_lo] = d
            vside[on_lo] = 0
            vdir[on_hi] = d                                # max face is tested second (:284, :299, :311)
            vside[on_hi] = 1
    src_v = -np.ones((nv, 3), dtype=np.int64)
    wts = np.zeros((nv, 3))
    dxs = np.zeros((nv, 3))
    for d in range(3):
        if not pdir[d]:
            continue
        other = [a for a in range(3) if a != d]
        for side in (0, 1):
            vs = np.nonzero((vdir == d) & (vside == side))[0]
            if len(vs) == 0:
                continue
            target = hi[d] if side == 0 else lo[d]         # mirrored onto the opposite face
            if planar:
                o = 1 - d
                seg = bf[:, :2] if bf is not None else _line_segments(xyz, tets, d, target).astype(np.int64)
                seg = seg[np.all(np.abs(xyz[seg][:, :, d] - target) <= FACE_TOL, axis=1)]
                if len(seg) == 0:
                    continue
                s_idx, W = _locate_1d(xyz[vs, o], xyz[seg][:, :, o])
                src_v[vs, :2] = seg[s_idx]
                wts[vs, :2] = W
                dxs[vs, d] = target - xyz[vs, d]
                continue
            if bf is not None:
                tri = bf[np.all(np.abs(xyz[bf][:, :, d] - target) <= FACE_TOL, axis=1)]
            else:
                tri = _plane_triangles(xyz, tets, d, target).astype(np.int64)
            if len(tri) == 0:
                continue
            t_idx, W = _locate(xyz[vs][:, other], xyz[tri][:, :, other])
            src_v[vs] = tri[t_idx]
            wts[vs] = W
            dxs[vs, d] = target - xyz[vs, d]
    rows = np.nonzero(vdir >= 0)[0]
    dof, src, w, dx = [], [], [], []
    for comp in (0, 1):
        act = rows[vc2dof[rows, comp] >= 0]
        if len(act) == 0:
            continue
        dof.append(vc2dof[act, comp])
        s = vc2dof[src_v[act], comp]
        s[src_v[act] < 0] = -1
        src.append(s)
        w.append(wts[act])
        dx.append(dxs[act])
    if not dof:
        return (np.zeros(0, np.int32), np.zeros((0, 3), np.int32), np.zeros((0, 3)), np.zeros((0, 3)))
    dof = np.concatenate(dof)
    order = np.argsort(dof, kind="stable")
    return (dof[order].astype(np.int32), np.concatenate(src)[order].astype(np.int32),
            np.concatenate(w)[order], np.concatenate(dx)[order])
