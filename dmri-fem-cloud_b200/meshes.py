"""Mesh inputs of the Bloch-Torrey path: readers for the reference's fixture formats and
deterministic synthetic generators for the benchmark shapes.

Readers follow the numbering rules of the reference's tooling so that "mesh numbering"
means the same thing on both sides:

* gmsh v2 ASCII ``.msh``: vertices = the nodes used by tetrahedra, numbered in the order
  they appear in ``$Nodes``; cells in file order
  (comri/meshes/neuron_download/dolfin-convert.py:357-597); cell marker = first gmsh tag
  (DmriFemLib.py:703-742, token ``x[3]``).
* DOLFIN XML ``<mesh>``: vertex ``index`` / tetrahedron ``index`` attributes.

Meshes are plain numpy: xyz (nv,3) float64, tets (nc,4) int32, optional marker (nc,) int32.
"""
import io
import re
import zipfile

import numpy as np


# ----------------------------------------------------------------------------- readers


def _open_text(path):
    if str(path).endswith(".zip"):
        with zipfile.ZipFile(path) as z:
            names = [n for n in z.namelist() if not n.startswith("__MACOSX") and not n.endswith("/")]
            return io.StringIO(z.read(names[0]).decode("utf-8", "replace"))
    return open(path, "r")


def read_gmsh2(path):
    """Read a gmsh v2 ASCII mesh (optionally zipped): the tetrahedra (element type 4), or, when the file has
    none, the triangles (type 2) -- a 2-D mesh, returned with two coordinate columns when every z is 0
    (dolfin-convert writes gdim 2 then), else a surface in 3-D."""
    with _open_text(path) as f:
        lines = f.read().split("\n")
    i = lines.index("$Nodes")
    nn = int(lines[i + 1])
    node_ids = np.empty(nn, dtype=np.int64)
    coords = np.empty((nn, 3))
    for k in range(nn):
        p = lines[i + 2 + k].split()
        node_ids[k] = int(p[0])
        coords[k] = (float(p[1]), float(p[2]), float(p[3]))
    j = lines.index("$Elements")
    ne = int(lines[j + 1])
    cells = {4: ([], []), 2: ([], [])}
    nvert = {4: 4, 2: 3}
    for k in range(ne):
        p = lines[j + 2 + k].split()
        et = int(p[1])
        if et in cells:
            nt = int(p[2])
            cells[et][0].append([int(a) for a in p[3 + nt:3 + nt + nvert[et]]])
            cells[et][1].append(int(p[3]) if nt > 0 else 0)
    et = 4 if cells[4][0] else 2
    tets = np.array(cells[et][0], dtype=np.int64).reshape(-1, nvert[et])
    marks = cells[et][1]
    used = np.zeros(node_ids.max() + 1, dtype=bool)
    used[tets.ravel()] = True
    keep = used[node_ids]                       # file order of $Nodes, unused nodes dropped
    new_id = -np.ones(node_ids.max() + 1, dtype=np.int64)
    new_id[node_ids[keep]] = np.arange(int(keep.sum()))
    xyz = coords[keep].copy()
    if et == 2 and not np.any(xyz[:, 2]):
        xyz = xyz[:, :2].copy()
    return xyz, new_id[tets].astype(np.int32), np.array(marks, dtype=np.int32)


def write_gmsh2(path, xyz, cells, tags=None):
    """Write a gmsh v2 ASCII mesh (`$MeshFormat 2.2 0 8`): tetrahedra (element type 4) or triangles (type 2), one
    physical/elementary tag pair per element -- the format the reference's `.msh` fixtures come in
    (comri/meshes/*.msh.zip) and read_gmsh2 / GetPartitionMarkers (DmriFemLib.py:703-742) parse."""
    xyz = np.asarray(xyz, dtype=float)
    cells = np.asarray(cells)
    x3 = np.zeros((len(xyz), 3))
    x3[:, :xyz.shape[1]] = xyz
    et = {4: 4, 3: 2}[cells.shape[1]]
    tags = np.zeros(len(cells), dtype=np.int64) if tags is None else np.asarray(tags)
    with open(path, "w") as f:
        f.write("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n%d\n" % len(x3))
        for i, p in enumerate(x3):
            f.write("%d %.17g %.17g %.17g\n" % (i + 1, p[0], p[1], p[2]))
        f.write("$EndNodes\n$Elements\n%d\n" % len(cells))
        for i, c in enumerate(cells):
            f.write("%d %d 2 %d %d %s\n" % (i + 1, et, tags[i], tags[i], " ".join(str(int(v) + 1) for v in c)))
        f.write("$EndElements\n")


def read_dolfin_xml(path):
    """Read a DOLFIN XML mesh (optionally zipped): celltype tetrahedron, triangle (dim 2 or 3) or interval."""
    with _open_text(path) as f:
        txt = f.read()
    nv = int(re.search(r'<vertices size="(\d+)"', txt).group(1))
    nc = int(re.search(r'<cells size="(\d+)"', txt).group(1))
    head = re.search(r'<mesh celltype="(\w+)" dim="(\d+)"', txt)
    celltype, gdim = (head.group(1), int(head.group(2))) if head else ("tetrahedron", 3)
    xyz = np.zeros((nv, 3))
    for m in re.finditer(r'<vertex index="(\d+)" x="([^"]+)" y="([^"]+)"(?: z="([^"]+)")?', txt):
        xyz[int(m.group(1))] = (float(m.group(2)), float(m.group(3)), float(m.group(4) or 0.0))
    if celltype == "interval":      # curve in 3-D (neuron skeleton, Manifolds.ipynb)
        segs = np.zeros((nc, 2), dtype=np.int32)
        for m in re.finditer(r'<interval index="(\d+)" v0="(\d+)" v1="(\d+)"', txt):
            segs[int(m.group(1))] = [int(m.group(2)), int(m.group(3))]
        return (xyz[:, :gdim].copy() if gdim < 3 else xyz), segs
    if celltype == "triangle":
        tris = np.zeros((nc, 3), dtype=np.int32)
        for m in re.finditer(r'<triangle index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)"', txt):
            tris[int(m.group(1))] = [int(m.group(k)) for k in (2, 3, 4)]
        return (xyz[:, :2].copy() if gdim == 2 else xyz), tris
    tets = np.zeros((nc, 4), dtype=np.int32)
    for m in re.finditer(r'<tetrahedron index="(\d+)" v0="(\d+)" v1="(\d+)" v2="(\d+)" v3="(\d+)"', txt):
        tets[int(m.group(1))] = [int(m.group(k)) for k in (2, 3, 4, 5)]
    return xyz, tets


def read_dolfin_markers(path):
    """Cell markers from a DOLFIN XML <mesh_value_collection> / <mesh_function> file, as msh2xml writes them
    (DmriFemLib.py:725-746: one <value cell_index=.. local_entity="0" value=..> per cell); phase = marker % 2."""
    with _open_text(path) as f:
        txt = f.read()
    m = re.search(r'<mesh_value_collection[^>]*size="(\d+)"', txt)
    vals = [(int(a), int(b)) for a, b in re.findall(r'<value cell_index="(\d+)"[^>]*value="(\d+)"', txt)]
    if not vals:       # plain <mesh_function>: <entity index=.. value=..>
        vals = [(int(a), int(b)) for a, b in re.findall(r'<entity index="(\d+)" value="(\d+)"', txt)]
        m = re.search(r'<mesh_function[^>]*size="(\d+)"', txt)
    n = int(m.group(1)) if m else (max(v[0] for v in vals) + 1)
    out = np.zeros(n, dtype=np.int32)
    for c, v in vals:
        out[c] = v
    return out


def phase_from_submesh(xyz, tets, sub_xyz, sub_tets, decimals=9):
    """Cells of (xyz,tets) that are also cells of the sub-mesh get phase 1, others 0.

    Restates CreatePhaseFunc(mymesh, [], [cmpt_mesh], None) (DmriFemLib.py:766-793;
    PreprocessingMultiCompt.py:107-114): a cell is in the odd group when its midpoint lies
    inside the compartment mesh.  For a sub-mesh that is a cell subset of the parent this
    is the same as matching rounded cell midpoints."""
    mid = np.round(xyz[tets].mean(axis=1), decimals)
    smid = np.round(sub_xyz[sub_tets].mean(axis=1), decimals)
    keys = {tuple(r) for r in smid}
    return np.array([1 if tuple(r) in keys else 0 for r in mid], dtype=np.int32)


# ----------------------------------------------------------------------------- generators

# Kuhn/Freudenthal split of the unit cube around the (0,0,0)-(1,1,1) diagonal, cube corner
# numbering v = i + 2j + 4k.  Same triangulation DOLFIN's BoxMesh produces (SURVEY App. B note ii).
_KUHN = np.array([[0, 1, 3, 7], [0, 1, 7, 5], [0, 5, 7, 4], [0, 3, 2, 7], [0, 6, 4, 7], [0, 2, 6, 7]])


def box_mesh(p0, p1, nx, ny, nz):
    """Structured box, vertices numbered x-fastest, 6 Kuhn tets per cube."""
    xs = np.linspace(p0[0], p1[0], nx + 1)
    ys = np.linspace(p0[1], p1[1], ny + 1)
    zs = np.linspace(p0[2], p1[2], nz + 1)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    xyz = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    base = (i + (nx + 1) * (j + (ny + 1) * k)).ravel()
    off = np.array([di + (nx + 1) * (dj + (ny + 1) * dk) for dk in (0, 1) for dj in (0, 1) for di in (0, 1)])
    corners = base[:, None] + off[None, :]                       # (ncube, 8)
    tets = corners[:, _KUHN].reshape(-1, 4).astype(np.int32)
    return xyz, tets


def extrude_triangulation(xy, tris, zs):
    """Extrude a 2-D triangulation into prisms and split each prism into 3 tets with the
    smallest-global-index rule, which makes neighbouring prisms agree on every shared quad."""
    nv2 = len(xy)
    nl = len(zs)
    xyz = np.concatenate([np.column_stack([xy, np.full(nv2, z)]) for z in zs], axis=0)
    tets = []
    tris = np.sort(np.asarray(tris), axis=1)                    # a<b<c within each triangle
    for l in range(nl - 1):
        a, b, c = (tris[:, m] + l * nv2 for m in range(3))
        A, B, C = a + nv2, b + nv2, c + nv2
        # bottom (a,b,c), top (A,B,C); a<b<c and x<X, so the diagonals chosen are a-B, a-C, b-C
        tets.append(np.stack([a, b, c, C], axis=1))
        tets.append(np.stack([a, b, C, B], axis=1))
        tets.append(np.stack([a, B, C, A], axis=1))
    return xyz, np.concatenate(tets, axis=0).astype(np.int32)


def disk_triangulation(radii, nr_per_layer, nsec):
    """Polar triangulation of concentric rings.  radii: increasing layer radii
    [R1, R2, ...]; nr_per_layer: radial subdivisions per layer; nsec: sectors.
    Returns xy, tris, layer (per triangle, 0 = innermost)."""
    rs, lay = [], []
    r_prev = 0.0
    for li, (R, n) in enumerate(zip(radii, nr_per_layer)):
        for m in range(1, n + 1):
            rs.append(r_prev + (R - r_prev) * m / n)
            lay.append(li)
        r_prev = R
    th = 2 * np.pi * np.arange(nsec) / nsec
    xy = [np.zeros((1, 2))]
    for r in rs:
        xy.append(np.column_stack([r * np.cos(th), r * np.sin(th)]))
    xy = np.concatenate(xy, axis=0)
    tris, tl = [], []
    ring = lambda m, s: 1 + m * nsec + (s % nsec)
    for s in range(nsec):
        tris.append([0, ring(0, s), ring(0, s + 1)])
        tl.append(lay[0])
    for m in range(1, len(rs)):
        for s in range(nsec):
            a, b = ring(m - 1, s), ring(m - 1, s + 1)
            c, d = ring(m, s), ring(m, s + 1)
            tris.append([a, c, d])
            tris.append([a, d, b])
            tl += [lay[m], lay[m]]
    return xy, np.array(tris, dtype=np.int64), np.array(tl, dtype=np.int32)


def layered_cylinder(radii=(5.0, 7.5, 10.0), height=5.0, nr_per_layer=(6, 3, 3), nsec=48, nz=4):
    """Concentric-layer cylinder (axis z) with conforming interfaces; marker = layer index
    (phase = marker % 2, DmriFemLib.py:764)."""
    xy, tris, tl = disk_triangulation(radii, nr_per_layer, nsec)
    zs = np.linspace(-height / 2, height / 2, nz + 1)
    xyz, tets = extrude_triangulation(xy, tris, zs)
    marker = np.concatenate([np.concatenate([tl, tl, tl]) for _ in range(nz)]).astype(np.int32)
    return xyz, tets, marker


def icosphere(level=2):
    """Unit-sphere triangulation by `level` midpoint subdivisions of the icosahedron (20 * 4^level triangles)."""
    t = (1.0 + np.sqrt(5.0)) / 2.0
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], dtype=float)
    v /= np.linalg.norm(v, axis=1)[:, None]
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2],
                  [10, 7, 6], [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11],
                  [6, 2, 10], [8, 6, 7], [9, 8, 1]])
    verts = [tuple(p) for p in v]
    for _ in range(level):
        cache, nf = {}, []

        def mid(a, b):
            key = (min(a, b), max(a, b))
            if key not in cache:
                m = np.asarray(verts[a]) + np.asarray(verts[b])
                verts.append(tuple(m / np.linalg.norm(m)))
                cache[key] = len(verts) - 1
            return cache[key]

        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [[a, ab, ca], [b, bc, ab], [c, ca, bc], [ab, bc, ca]]
        f = np.array(nf)
    return np.array(verts), f


def layered_sphere(radii=(5.0, 7.5, 10.0), nr_per_layer=(4, 2, 2), level=2):
    """Concentric-layer ball with conforming interfaces (the multilayered sphere of T2_Relaxation.ipynb /
    MultilayeredStructures.ipynb cell 12): shells of an icosphere at increasing radii, the prisms between two
    shells split into 3 tets by the smallest-index rule (like extrude_triangulation), the innermost shell coned to
    the centre.  marker = layer index (phase = marker % 2)."""
    sv, sf = icosphere(level)
    rs, lay = [], []
    r_prev = 0.0
    for li, (R, n) in enumerate(zip(radii, nr_per_layer)):
        for m in range(1, n + 1):
            rs.append(r_prev + (R - r_prev) * m / n)
            lay.append(li)
        r_prev = R
    ns = len(sv)
    xyz = np.concatenate([np.zeros((1, 3))] + [r * sv for r in rs], axis=0)
    shell = lambda m, idx: 1 + m * ns + idx
    tets = [np.column_stack([np.zeros(len(sf), dtype=np.int64), shell(0, sf[:, 0]), shell(0, sf[:, 1]), shell(0, sf[:, 2])])]
    marker = [np.full(len(sf), lay[0])]
    tri = np.sort(sf, axis=1)
    for m in range(1, len(rs)):
        a, b, c = (shell(m - 1, tri[:, q]) for q in range(3))
        A, B, C = (shell(m, tri[:, q]) for q in range(3))
        tets += [np.stack([a, b, c, C], axis=1), np.stack([a, b, C, B], axis=1), np.stack([a, B, C, A], axis=1)]
        marker += [np.full(len(sf), lay[m])] * 3
    return xyz, np.concatenate(tets, axis=0).astype(np.int32), np.concatenate(marker).astype(np.int32)


def cylinder(radius=3.0, length=25.0, nr=4, nsec=16, nz=20):
    """Single-compartment cylinder of the `cyl*_r_3E_6` shape (axis z)."""
    xyz, tets, _ = layered_cylinder((radius,), length, (nr,), nsec, nz)
    return xyz, tets


def box_with_sphere(half=10.0, n=20, radius=5.0):
    """Cell-in-box: Kuhn box [-half,half]^3 with phase 1 where the cell centroid lies in the
    sphere (synthetic)."""
    xyz, tets = box_mesh((-half,) * 3, (half,) * 3, n, n, n)
    cen = xyz[tets].mean(axis=1)
    phase = (np.linalg.norm(cen, axis=1) < radius).astype(np.int32)
    return xyz, tets, phase


def ecs_slab(nx, ny, nz=2, lx=77.09, ly=75.89, lz=1.0, ncyl=226, rmin=2.0, rmax=5.0, seed=226):
    """Slab with `ncyl` seeded non-overlapping circular cylinders (axis z); phase 1 inside
    the cylinders by cell-centroid test (ECS_226Cylinders.ipynb bbox)."""
    rng = np.random.default_rng(seed)
    cs, rs = [], []
    tries = 0
    while len(cs) < ncyl and tries < 200000:
        tries += 1
        r = rng.uniform(rmin, rmax)
        c = np.array([rng.uniform(-lx + r, lx - r), rng.uniform(-ly + r, ly - r)])
        if all(np.linalg.norm(c - c2) > r + r2 + 0.3 for c2, r2 in zip(cs, rs)):
            cs.append(c)
            rs.append(r)
    cs, rs = np.array(cs), np.array(rs)
    xyz, tets = box_mesh((-lx, -ly, -lz), (lx, ly, lz), nx, ny, nz)
    cen = xyz[tets].mean(axis=1)[:, :2]
    phase = np.zeros(len(tets), dtype=np.int32)
    for c, r in zip(cs, rs):
        phase |= (np.linalg.norm(cen - c[None], axis=1) < r).astype(np.int32)
    return xyz, tets, phase


def neuron_like(n_dend=8, soma_r=10.0, dend_r=1.0, dend_len=200.0, h=1.0, seed=5):
    """Neuron-like shape: a soma ball with straight dendrite tubes, carved out of a Kuhn
    grid of spacing h by a cell-centroid test (synthetic)."""
    rng = np.random.default_rng(seed)
    dirs = rng.normal(size=(n_dend, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    L = soma_r + dend_len
    n = int(np.ceil(2 * L / h))
    # build only the occupied cubes: march along each dendrite
    cubes = set()
    rr = int(np.ceil(soma_r / h)) + 1
    for i in range(-rr, rr + 1):
        for j in range(-rr, rr + 1):
            for k in range(-rr, rr + 1):
                cubes.add((i, j, k))
    dr = int(np.ceil(dend_r / h)) + 1
    for d in dirs:
        for s in np.arange(0.0, L, h / 2):
            c = np.floor(d * s / h).astype(int)
            for i in range(-dr, dr + 1):
                for j in range(-dr, dr + 1):
                    for k in range(-dr, dr + 1):
                        cubes.add((c[0] + i, c[1] + j, c[2] + k))
    cubes = np.array(sorted(cubes), dtype=np.int64)
    offs = np.array([[di, dj, dk] for dk in (0, 1) for dj in (0, 1) for di in (0, 1)])
    corner_ijk = cubes[:, None, :] + offs[None, :, :]                      # (ncube,8,3)
    big = 1 << 20
    key = (corner_ijk[..., 0] + big) + ((corner_ijk[..., 1] + big) << 21) + ((corner_ijk[..., 2] + big) << 42)
    uk, inv = np.unique(key.ravel(), return_inverse=True)
    corners = inv.reshape(-1, 8)
    vi = (uk & ((1 << 21) - 1)) - big
    vj = ((uk >> 21) & ((1 << 21) - 1)) - big
    vk = (uk >> 42) - big
    xyz = np.column_stack([vi, vj, vk]).astype(float) * h
    tets = corners[:, _KUHN].reshape(-1, 4)
    cen = xyz[tets].mean(axis=1)
    inside = np.linalg.norm(cen, axis=1) < soma_r
    for d in dirs:
        s = cen @ d
        perp = np.linalg.norm(cen - s[:, None] * d[None], axis=1)
        inside |= (s > 0) & (s < L) & (perp < dend_r)
    tets = tets[inside]
    used = np.unique(tets)
    remap = -np.ones(len(xyz), dtype=np.int64)
    remap[used] = np.arange(len(used))
    return xyz[used].copy(), remap[tets].astype(np.int32)


def shuffle_vertices(xyz, tets, seed=0):
    """Random vertex renumbering (what an unstructured mesher hands you)."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(xyz))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    return xyz[perm].copy(), inv[tets].astype(np.int32)


def coordinate_order(xyz, tets, axes=(1, 2, 0), decimals=9):
    """Renumber vertices lexicographically by coordinates, axes[0] slowest: contiguous vertex blocks are then slabs
    normal to axes[0] (what the row partition over GPUs wants for slab-like domains such as the ECS)."""
    key = np.round(xyz, decimals)
    perm = np.lexsort(tuple(key[:, a] for a in reversed(axes)))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    return xyz[perm].copy(), inv[tets].astype(np.int32)


def rcm_order(xyz, tets):
    """Reverse Cuthill-McKee vertex renumbering for gather locality in the SpMV."""
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    nv = len(xyz)
    r = np.repeat(tets, 4, axis=1).ravel()
    c = np.tile(tets, (1, 4)).ravel()
    A = sp.coo_matrix((np.ones(len(r), dtype=np.int8), (r, c)), shape=(nv, nv)).tocsr()
    perm = reverse_cuthill_mckee(A, symmetric_mode=True)
    inv = np.empty_like(perm)
    inv[perm] = np.arange(nv)
    return xyz[perm].copy(), inv[tets].astype(np.int32)


def fibonacci_hemisphere(n, seed=64):
    """n gradient directions spread over the half-sphere (HARDI sweep)."""
    i = np.arange(n) + 0.5
    z = i / n
    phi = np.pi * (1 + 5 ** 0.5) * i
    r = np.sqrt(1 - z * z)
    return np.column_stack([r * np.cos(phi), r * np.sin(phi), z])
