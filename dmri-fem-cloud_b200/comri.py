"""The comri FEniCS-HPC demos on libbtfem: same command-line flags, same scheme, same result line.

The reference ships three C++ drivers built on DOLFIN-HPC with pre-assembled matrices
(`A = MSI + f(t)*gnorm*J` by dup / scale / add every step, RHS re-assembled every step):

  one-comp    comri/one-comp/hpc-fenics-cpp/main.cpp     -N -b -q -d -D -j -m -f -r -k -s -v gx gy gz -K
  two-comp    comri/two-comp/hpc-fenics-cpp/main.cpp     the same + -c cell.xml (compartment sub-mesh) -p kappa
  multilayer  comri/multilayer/hpc-fenics-cpp/main.cpp   -N -b -d -D -j -m -f -r

What differs from DmriFemLib.solve and is kept here:
  * f(t_n) on BOTH sides of the theta step (the linear form reads the same `ft_f` as the matrix,
    one-comp main.cpp:300-301, 315) -- DmriFemLib lags the right-hand side by one step;
  * FT(t) = 1 on [0,delta), -1 on [Delta, Delta+delta] (closed at the end, main.cpp:86-91);
  * loop `while (t < T + dt)` (one-comp :296, two-comp :884), `while (t < T)` in multilayer (:301);
  * dt = T/Nsteps unless -k is given (one-comp :214-215); multilayer has no -k;
  * "gnorm" is DmriFemLib's q-value: sqrt(b)/sqrt(delta^2 (Delta - delta/3)); "qvalue" = gnorm/2.675e8*1e12;
  * phase: two-comp marks cells whose midpoint lies in the -c sub-mesh (MarkPhase :288-335), without -c the
    spherical shells r = R0*[1, 1.5, 2], R0 = 1 (:242-285); multilayer marks torus shells, R = 20,
    r = 5*[1, 1.5, 2] (:113-135), value = shell index % 2;
  * Krylov: bicgstab + jacobi (one-comp :231, two-comp :744), bicgstab + none in multilayer (:237); the
    tolerances are DOLFIN-HPC parameter defaults (third party, not in the tree): KRYLOV below, overridable;
  * initial condition: L2 projection of 1 (InitialCondition3D.ufl) = 1 in the P1 space; result s/s0 with
    s = Comp_Sig3D functional (int u_r, phase-weighted in two compartments).

Not carried over: -r (uniform refinement needs a mesh refiner), -s / -f (DOLFIN .bin output files), the
artificial-permeability boundary term of two-comp, which its own main.cpp comments out (:896-900).
"""
import sys
import time

import numpy as np

from . import btfem as _bt
from . import meshes

KRYLOV = {"rtol": 1e-12, "atol": 1e-15, "maxit": 100000}
G_RATIO = 2.675e8


def FT(t, delta, Delta):
    """comri/one-comp/hpc-fenics-cpp/main.cpp:86-91."""
    return 1.0 * (0 <= t < delta) - 1.0 * (Delta <= t <= Delta + delta)


def time_grid(T, dt, variant):
    """t accumulated like the C++ loop (`t += dt`)."""
    ts, t = [], 0.0
    end = T if variant == "multilayer" else T + dt
    while t < end:
        ts.append(t)
        t += dt
    return np.array(ts)


def shell_phase(mid, variant):
    """MarkPhase without a region file: value = (index of the first shell containing the midpoint) % 2."""
    if variant == "multilayer":
        R, r = 20.0, 5.0 * np.array([1.0, 1.5, 2.0])
        d = (R - np.hypot(mid[:, 0], mid[:, 1])) ** 2 + mid[:, 2] ** 2
        inside = d[:, None] < (r * r)[None, :]
    else:
        r = np.array([1.0, 1.5])                    # `i < ncomps-1` (two-comp main.cpp:272)
        inside = np.linalg.norm(mid, axis=1)[:, None] < r[None, :]
    first = np.where(inside.any(axis=1), inside.argmax(axis=1), 0)
    return (first % 2).astype(np.int32)


def load_mesh(path):
    if ".msh" in path:
        xyz, cells, _ = meshes.read_gmsh2(path)
    elif path.endswith(".npz"):
        z = np.load(path)
        xyz, cells = z["xyz"], z["tets"]
    else:
        xyz, cells = meshes.read_dolfin_xml(path)
    return xyz, cells


def parse(variant, argv):
    """The `switch (argv[optind][1])` of the three mains; defaults as declared there."""
    p = {"one-comp": dict(N=100, b=1000.0, delta=40000.0, Delta=40000.0, dt=100.0, K=2.4e-3, g=(0.0, 1.0, 0.0)),
         "two-comp": dict(N=100, b=4000.0, delta=40000.0, Delta=40000.0, dt=100.0, K=3e-3, g=(0.0, 1.0, 0.0),
                          kappa=5e-5),
         "multilayer": dict(N=100, b=1000.0, delta=1000.0, Delta=1000.0, dt=None, K=3e-3, g=(0.0, 0.0, 1.0),
                            kappa=5e-5)}[variant]
    p.update(q=None, nskip=5, mesh="mesh.xml", cell=None, dir="results", nrefine=0, is_dt=False, is_b=False,
             is_q=False)
    full = variant != "multilayer"
    i = 1
    while i < len(argv):
        a = argv[i]
        if a.startswith("-") and len(a) > 1:
            c = a[1]
            nxt = argv[i + 1] if i + 1 < len(argv) else None
            if c == "N":
                p["N"] = int(nxt)
            elif c == "b":
                p["b"], p["is_b"] = float(nxt), True
            elif c == "q" and full:
                p["q"], p["is_q"] = float(nxt), True
            elif c == "d":
                p["delta"] = float(nxt)
            elif c == "D":
                p["Delta"] = float(nxt)
            elif c == "j":
                p["nskip"] = int(nxt)
            elif c == "m":
                p["mesh"] = nxt
            elif c == "f":
                p["dir"] = nxt
            elif c == "r":
                p["nrefine"] = int(nxt)
            elif c == "k" and full:
                p["dt"], p["is_dt"] = float(nxt), True
            elif c == "v" and full:
                g = np.array([float(argv[i + 1]), float(argv[i + 2]), float(argv[i + 3])])
                p["g"] = tuple(g / np.linalg.norm(g))
            elif c == "K" and full:
                p["K"] = float(nxt)
            elif c == "c" and variant == "two-comp":
                p["cell"] = nxt
            elif c == "p" and variant == "two-comp":
                p["kappa"] = float(nxt)
        i += 1
    return p


def run(variant, p, device=0, out=sys.stdout):
    """Returns dict(s=s/s0, s0, signal, stats, ts).  `p`: parse() result (or the same keys)."""
    start = time.time()
    delta, Delta = p["delta"], p["Delta"]
    den = np.sqrt(delta * delta * (Delta - delta / 3.0))
    if variant == "multilayer" or p.get("is_b") or not p.get("is_q"):
        bvalue = p["b"]
        gnorm = np.sqrt(bvalue) / den
        qvalue = gnorm / G_RATIO * 1e12
    else:
        qvalue = p["q"]
        gnorm = qvalue * G_RATIO * 1e-12
        bvalue = gnorm * gnorm * delta * delta * (Delta - delta / 3.0)
    if p.get("nrefine"):
        raise RuntimeError("-r (uniform mesh refinement) is not available: refine the mesh before calling")
    print("\nReading mesh...", file=out)
    xyz, cells = load_mesh(p["mesh"]) if isinstance(p["mesh"], str) else p["mesh"]
    print("done\n", file=out)
    phase = None
    if variant != "one-comp":
        print("Generating phase function...", file=out)
        mid = np.asarray(xyz)[np.asarray(cells)].mean(axis=1)
        if variant == "two-comp" and p.get("cell") is not None:
            print("Reading a given submesh", file=out)
            sub = load_mesh(p["cell"]) if isinstance(p["cell"], str) else p["cell"]
            phase = meshes.phase_from_submesh(xyz, cells, sub[0], sub[1])
        else:
            if variant == "two-comp":
                print("Submesh is not given", file=out)
            phase = shell_phase(mid, variant)
        print("done", file=out)
    T = Delta + delta
    dt = p["dt"] if (p.get("is_dt") and variant != "multilayer") else T / p["N"]
    theta = 0.5
    g = np.asarray(p["g"], dtype=float)
    if variant == "one-comp":
        print("kcoeff: %e, Nsteps: %d, dt: %f, delta: %f, Delta: %f, gnorm: %f\n" % (p["K"], p["N"], dt, delta, Delta,
                                                                                      gnorm), file=out)
    elif variant == "two-comp":
        print("kcoeff: %e, perm: %e, Nsteps: %d, dt: %f, delta: %f, Delta: %f, gnorm: %e\n" % (
            p["K"], p["kappa"], p["N"], dt, delta, Delta, gnorm), file=out)
    print("Gradient direction: %f %f %f" % tuple(g), file=out)
    ts = time_grid(T, dt, variant)
    c = gnorm * np.array([FT(t, delta, Delta) for t in ts])
    with _bt.BTFem(device) as fem:
        fem.set_mesh(xyz, cells, phase)
        fem.set_diffusion(p["K"])
        if phase is not None:
            fem.set_permeability(p["kappa"])
        print("Preparing no-time matrices ...", file=out)
        fem.assemble()
        for n in range(0, len(ts), 1 if variant != "one-comp" else max(1, p["nskip"])):
            print("t=%f, dt=%f, gnorm=%f, step_counter=%d, Completed %.1f%%\n" % (ts[n], dt, gnorm, n, ts[n] / T * 100),
                  file=out)
        stats = fem.solve(dt, theta, c, c, g, pc="none" if variant == "multilayer" else "jacobi", **KRYLOV)
    s0, s = stats["voi"], stats["signal"]
    print("s0=%f\n" % s0, file=out)
    if variant == "one-comp":
        print("b: %f, gnorm: %f, q: %f, gdir: (%f, %f, %f), s: %f\n" % (bvalue, gnorm, qvalue, g[0], g[1], g[2], s / s0),
              file=out)
    elif variant == "two-comp":
        print("b: %f, gnorm: %e, q: %e, perm: %e, gdir: (%f, %f, %f), s: %f\n" % (bvalue, gnorm, qvalue, p["kappa"],
                                                                                   g[0], g[1], g[2], s / s0), file=out)
    else:
        print("s=%f\n" % (s / s0), file=out)
    print("Runtime = %f\n" % (time.time() - start), file=out)
    return dict(s=s / s0, s0=s0, signal=s, stats=stats, ts=ts, gnorm=gnorm, qvalue=qvalue, bvalue=bvalue, dt=dt,
                phase=phase)


def main(argv=None):
    argv = list(sys.argv if argv is None else argv)
    if len(argv) < 2 or argv[1] not in ("one-comp", "two-comp", "multilayer"):
        print("usage: comri.py one-comp|two-comp|multilayer [flags of the corresponding comri main.cpp]")
        return 2
    variant = argv[1]
    run(variant, parse(variant, argv[1:]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
