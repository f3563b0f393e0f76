"""GCloudDmriSolver.py-compatible command line on top of libbtfem.

Same flags, defaults and quirks as /root/reference/GCloudDmriSolver.py:52-133:
  -f file  -M 0|1  -isperiodic|-IsPeriodic 0|1  -N n (parsed, unused)  -b bvalue  -q qvalue (parsed,
  ignored: bvalue is always used, GCloudDmriSolver.py:184)  -p kappa  -D Delta  -d delta  -K D0
  -k dt  -T2 T2  -gdir gx gy gz  -pdir px py pz
A parse error prints 'Something goes wrong with the inputs!' and continues with what was parsed.

T2 quirk, kept: the reference parses `-T2` and reads the `T2` dataset (GCloudDmriSolver.py:115-119, 167-171) but
never assigns either to `mri_para.T2`, so its forms always run with T2 = 1e16 (DmriFemLib.py:807).  Identical
flags give identical signals here: T2 is parsed, read and -- by default -- not applied.  `-applyT2 1` (an
extension, not a reference flag) applies it: the per-cell `T2` dataset, or the `-T2` scalar.

Input file: the DOLFIN HDF5 container the pre-processing scripts write (datasets mesh, T2, ic, phase, d00..d22;
PreprocessingMultiCompt.py:148-152), read by hdf5io.py (pure Python: no libhdf5 in this image); the same datasets
from a NumPy `.npz` with keys
  xyz (nv,3), tets (nc,4) [, phase (nc,), T2 (nc,), ic (nc,), d00 .. d22 (nc,)]
or, mesh only, gmsh v2 `.msh[.zip]` / DOLFIN `.xml[.zip]` (then -K applies, like the reference's
is_kcoeff_from_file = 0 path).
"""
import sys

import numpy as np
import sympy as sp

from . import dmrifemlib as dl
from . import hdf5io
from . import meshes


def load_input(path):
    data = {}
    if path.endswith(".npz"):
        z = np.load(path)
        data = {k: z[k] for k in z.files}
    elif path.endswith(".h5") or path.endswith(".hdf5"):
        data = hdf5io.read_dolfin_h5(path)
    elif ".msh" in path:
        xyz, tets, marker = meshes.read_gmsh2(path)
        data = {"xyz": xyz, "tets": tets, "marker": marker}
    elif ".xml" in path:
        xyz, tets = meshes.read_dolfin_xml(path)
        data = {"xyz": xyz, "tets": tets}
    else:
        raise RuntimeError("unsupported input file: " + path)
    return data


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    # default parameters (GCloudDmriSolver.py:52-55)
    g0, g1, g2 = 0, 1, 0
    kcoeff = 3e-3
    Nsteps = 100
    bvalue = 1000
    kappa = 1e-5
    delta, Delta = 10600, 43100
    T2 = 1e16
    k = 200
    is_kcoeff_from_file = 1
    is_T2_from_file = 1
    IsDomainPeriodic = False
    IsDomainMultiple = False
    PeriodicDir = [0, 0, 0]
    ffile = None
    apply_T2 = 0
    try:
        for i in range(0, len(argv)):
            arg = argv[i]
            if arg == '-f':
                ffile = argv[i + 1]
                print('input file:', ffile)
            if arg == '-M':
                IsDomainMultiple = int(argv[i + 1])
                print('IsDomainMultiple:', IsDomainMultiple)
            if arg == '-IsPeriodic' or arg == '-isperiodic':
                IsDomainPeriodic = int(argv[i + 1])
                print('IsDomainPeriodic:', IsDomainPeriodic)
            if arg == '-N':
                Nsteps = int(argv[i + 1])
                print('Nsteps:', Nsteps)
            if arg == '-b':
                bvalue = float(argv[i + 1])
                print('bvalue:', bvalue)
            if arg == '-q':
                qvalue = float(argv[i + 1])
                print('qvalue:', bvalue)        # sic (GCloudDmriSolver.py:90)
            if arg == '-p':
                kappa = float(argv[i + 1])
                print('permeability:', kappa)
            if arg == '-D':
                Delta = float(argv[i + 1])
                print('Delta:', Delta)
            if arg == '-d':
                delta = float(argv[i + 1])
                print('delta:', delta)
            if arg == '-K':
                is_kcoeff_from_file = 0
                print("Reading diffusion coefficient from command line")
                kcoeff = float(argv[i + 1])
                print('diffusion coefficient:', kcoeff)
            if arg == '-k':
                k = float(argv[i + 1])
                print('time step size:', k)
            if arg == '-T2':
                is_T2_from_file = 0
                T2 = float(argv[i + 1])
                print('T2: ', T2)
            if arg == '-applyT2':
                apply_T2 = int(argv[i + 1])
                print('apply T2 (extension; the reference never does):', apply_T2)
            if arg == '-gdir':
                g0 = float(argv[i + 1])
                g1 = float(argv[i + 2])
                g2 = float(argv[i + 3])
                print('(g0, g1, g2):', g0, g1, g2)
            if arg == '-pdir':
                PeriodicDir[0] = int(argv[i + 1])
                PeriodicDir[1] = int(argv[i + 2])
                PeriodicDir[2] = int(argv[i + 3])
                print('PeriodicDir=[', PeriodicDir[0], PeriodicDir[1], PeriodicDir[2], ']')
    except Exception:
        print('Something goes wrong with the inputs!')

    data = load_input(ffile)
    mymesh = dl.Mesh(data["xyz"], data["tets"])
    have_tensor = all(("d%d%d" % (a, b)) in data for a in range(3) for b in range(3))
    if is_kcoeff_from_file == 1 and not have_tensor:
        is_kcoeff_from_file = 0                 # mesh-only inputs carry no tensor: fall back to -K
    if is_T2_from_file == 1 and "T2" not in data:
        is_T2_from_file = 0

    mri_simu = dl.MRI_simulation()
    mri_para = dl.MRI_parameters()
    mri_para.bvalue = bvalue
    mri_para.delta, mri_para.Delta = delta, Delta
    mri_para.set_gradient_dir(mymesh, g0, g1, g2)
    mri_para.T = mri_para.Delta + mri_para.delta
    mri_para.fs_sym = sp.Piecewise((1., mri_para.s < mri_para.delta), (0., mri_para.s < mri_para.Delta),
                                   (-1., mri_para.s < mri_para.T), (0., True))
    if is_T2_from_file == 1:
        print("Reading T2 from file: ", ffile)
    if apply_T2 and is_T2_from_file == 0:
        mri_para.T2 = T2
    mri_para.Apply()
    mri_simu.k = k
    mri_simu.nskip = 5

    mydomain = dl.MyDomain(mymesh, mri_para)
    if IsDomainMultiple == 1:
        if "phase" in data:
            mydomain.phase = np.asarray(data["phase"]).astype(np.int32)
        elif "marker" in data:
            mydomain.phase = (np.asarray(data["marker"]) % 2).astype(np.int32)   # DmriFemLib.py:764
        print("Reading phase function from file: ", ffile)
    mydomain.PeriodicDir = PeriodicDir
    mydomain.IsDomainPeriodic = IsDomainPeriodic
    mydomain.IsDomainMultiple = IsDomainMultiple
    mydomain.kappa = kappa
    if apply_T2 and is_T2_from_file == 1:
        mydomain.T2_cell = np.asarray(data["T2"], dtype=float)
    mydomain.Apply()
    if is_kcoeff_from_file == 1:
        print("Impose diffusion tensor from file: ", ffile)
        mydomain.ImposeDiffusionTensor(*[data["d%d%d" % (a, b)] for a in range(3) for b in range(3)])
    else:
        print("Impose diffusion coefficient from command line, D0=", kcoeff)
        mydomain.D0 = kcoeff
        mydomain.D = mydomain.D0

    linsolver = dl.KrylovSolver("bicgstab", "jacobi")
    linsolver.parameters["relative_tolerance"] = 1e-9
    linsolver.parameters["absolute_tolerance"] = 1e-10
    linsolver.parameters["maximum_iterations"] = 100000

    mri_simu.solve(mydomain, mri_para, linsolver)
    return dl.PostProcessing(mydomain, mri_para, mri_simu, None, '')


if __name__ == "__main__":
    main()
