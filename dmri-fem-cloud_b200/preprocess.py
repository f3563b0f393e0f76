"""Pre-processing of the solver input, the part of PreprocessingOneCompt.py / PreprocessingMultiCompt.py that is
not mesh generation: phase function and partition markers from compartment sub-meshes or from a marker file
(CreatePhaseFunc, DmriFemLib.py:748-799), per-compartment T2 / IC / diffusion tensor as cell (DG0) arrays
(PreprocessingMultiCompt.py:118-146), written -- like the reference -- to a DOLFIN HDF5 container `<name>.h5`
(mesh, T2, ic, phase, d00..d22, PreprocessingMultiCompt.py:148-152; hdf5io.write_dolfin_h5), or to the `.npz` side
format of cli.py when `-o` names a `.npz`.

  python -m ... preprocess  -m mesh.xml [-odd cmpt1.xml ...] [-even cmptA.xml ...] [-pmk pmk_mesh.xml]
                            [-D0 3e-3 3e-3] [-T2 1e6 1e6] [-IC 1 1] -o files.h5
"""
import sys

import numpy as np

from . import hdf5io
from . import meshes


def create_phase_func(xyz, cells, evengroup=(), oddgroup=(), partition_marker=None):
    """CreatePhaseFunc (DmriFemLib.py:748-799).  With `partition_marker` (one int per cell): phase = marker % 2 and
    the list of markers in order of first appearance.  Otherwise a cell whose midpoint lies in even sub-mesh k gets
    marker 2k+2 / phase 0, in odd sub-mesh k marker 2k+1 / phase 1 (odd groups are tested last and win), other cells
    keep marker 0; the partition list holds the markers of the given groups, then any further marker in front.
    Sub-meshes are (xyz, cells) pairs that are cell subsets of the mesh (meshes.phase_from_submesh)."""
    cells = np.asarray(cells)
    if partition_marker is not None:
        pm = np.asarray(partition_marker).astype(np.int64)
        plist = []
        for v in pm:
            if v not in plist:
                plist.append(int(v))
        return (pm % 2).astype(np.int32), plist
    pm = np.zeros(len(cells), dtype=np.int64)
    phase = np.zeros(len(cells), dtype=np.int32)
    if len(evengroup) and not len(oddgroup):
        phase[:] = 0                                   # DmriFemLib.py:767-770 (initial fill)
    elif len(oddgroup) and not len(evengroup):
        phase[:] = 1
    plist = [-1] * (len(evengroup) + len(oddgroup))
    for k, sub in enumerate(evengroup):
        inside = meshes.phase_from_submesh(xyz, cells, sub[0], sub[1]).astype(bool)
        pm[inside], phase[inside] = 2 * k + 2, 0
        if inside.any():
            plist[k] = 2 * k + 2
    for k, sub in enumerate(oddgroup):
        inside = meshes.phase_from_submesh(xyz, cells, sub[0], sub[1]).astype(bool)
        pm[inside], phase[inside] = 2 * k + 1, 1
        if inside.any():
            plist[k + len(evengroup)] = 2 * k + 1
    for v in pm:
        if int(v) not in plist:
            plist.insert(0, int(v))
    return phase, plist, pm.astype(np.int32)


def cell_fields(partition_marker, D0_array, T2_array, IC_array):
    """T2, ic and the diagonal diffusion tensor per cell, indexed by the partition marker
    (PreprocessingMultiCompt.py:131-146)."""
    pm = np.asarray(partition_marker)
    D = np.asarray(D0_array, dtype=float)[pm]
    z = np.zeros(len(pm))
    out = {"T2": np.asarray(T2_array, dtype=float)[pm], "ic": np.asarray(IC_array, dtype=float)[pm]}
    for a in range(3):
        for b in range(3):
            out["d%d%d" % (a, b)] = D.copy() if a == b else z.copy()
    return out


def _load(path):
    if ".msh" in path:
        xyz, cells, _ = meshes.read_gmsh2(path)
        return xyz, cells
    return meshes.read_dolfin_xml(path)


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    mesh, odd, even, pmk, ofile = None, [], [], None, "files.h5"
    D0, T2, IC = None, None, None

    def floats(i):
        vals = []
        while i < len(argv) and not (argv[i].startswith("-") and not argv[i][1:2].isdigit() and argv[i][1:2] != "."):
            vals.append(float(argv[i]))
            i += 1
        return vals

    for i, a in enumerate(argv):
        if a == "-m":
            mesh = argv[i + 1]
        elif a == "-odd":
            odd.append(argv[i + 1])
        elif a == "-even":
            even.append(argv[i + 1])
        elif a == "-pmk":
            pmk = argv[i + 1]
        elif a == "-o":
            ofile = argv[i + 1]
        elif a == "-D0":
            D0 = floats(i + 1)
        elif a == "-T2":
            T2 = floats(i + 1)
        elif a == "-IC":
            IC = floats(i + 1)
    if mesh is None:
        print(__doc__)
        return 2
    xyz, cells = _load(mesh)
    if pmk is not None:
        marker = meshes.read_dolfin_markers(pmk)
        phase, plist = create_phase_func(xyz, cells, partition_marker=marker)
    else:
        phase, plist, marker = create_phase_func(xyz, cells, [_load(p) for p in even], [_load(p) for p in odd])
    print("Partition markers:", plist)
    n = int(np.max(marker)) + 1
    fields = cell_fields(marker, D0 or [3e-3] * n, T2 or [1e6] * n, IC or [1.0] * n)     # PreprocessingMultiCompt.py:122-125
    if ofile.endswith(".npz"):
        np.savez(ofile, xyz=xyz, tets=cells, phase=phase, marker=marker, **fields)
    else:                                          # filename + '.h5' whatever the extension (PreprocessingMultiCompt.py:144-146)
        ofile = (ofile.rsplit(".", 1)[0] if "." in ofile.rsplit("/", 1)[-1] else ofile) + ".h5"
        hdf5io.write_dolfin_h5(ofile, xyz, cells, dict(phase=phase, **fields))
    print("Write to ", ofile)
    return 0


if __name__ == "__main__":
    sys.exit(main())
