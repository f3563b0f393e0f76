#!/usr/bin/env python3
"""TEST INFRASTRUCTURE ONLY (oracle/): build the reference's own element kernels.

The only arithmetic of the hot path that is present as source under /root/reference
is the FFC-generated UFC code of the comri FEniCS-HPC solvers.
The full path (DOLFIN assembler, PETSc KSP) is third-party and absent, so it cannot
be built here; the element kernels can.  This recipe

  1. reads the generated ``.cpp`` files where they lie under /root/reference,
  2. lifts each ``<integral class>::tabulate_tensor(double* A, const double* const* w,
     const ufc::cell& ...)`` definition (first overload; located by signature, not
     by line number) into a scratch translation unit together with ``ufc_shim/ufc.h``
     and flat C wrappers,
  3. compiles it with g++ into ``oracle/_ref/libufcref.so`` and deletes the scratch
     source again, so that no reference source text stays in the tree.

``oracle/_ref/`` is git-ignored (it is a build product) but travels to the GPU box.
Nothing outside tests/, __graft_entry__ and bench.py's cpu arm may load it.

Kernels lifted (reference file : class):
  one-comp/hpc-fenics-cpp/ufc/Bloch_Torrey_NoTime3D.cpp : cell_integral_0_0     (8x8,  m+j+s)
  one-comp/hpc-fenics-cpp/ufc/Bloch_Torrey3D.cpp        : cell_integral_0_0     (8x8,  theta-scheme a)
                                                          cell_integral_1_0     (8,    theta-scheme L)
  one-comp/hpc-fenics-cpp/ufc/Comp_Sig3D.cpp            : cell_integral_0_0     (signal functional)
  two-comp/hpc-fenics-cpp/ufc/Bloch_Torrey_NoTime3D.cpp : cell_integral_0_0     (16x16)
                                                          exterior_facet_integral_0_0 (16x16)
                                                          interior_facet_integral_0_0 (32x32)
  two-comp/hpc-fenics-cpp/ufc/Comp_Sig3D.cpp            : cell_integral_0_0
"""
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("BTFEM_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
OUT_SO = os.path.join(OUT_DIR, "libufcref.so")

# (namespace tag, reference-relative path, [(class suffix, kind)])
UNITS = [
    ("oc_notime", "comri/one-comp/hpc-fenics-cpp/ufc/Bloch_Torrey_NoTime3D.cpp",
     [("bloch_torrey_notime3d_cell_integral_0_0", "cell")]),
    ("oc_bt", "comri/one-comp/hpc-fenics-cpp/ufc/Bloch_Torrey3D.cpp",
     [("bloch_torrey3d_cell_integral_0_0", "cell"),
      ("bloch_torrey3d_cell_integral_1_0", "cell")]),
    ("oc_sig", "comri/one-comp/hpc-fenics-cpp/ufc/Comp_Sig3D.cpp",
     [("comp_sig3d_cell_integral_0_0", "cell")]),
    ("tc_notime", "comri/two-comp/hpc-fenics-cpp/ufc/Bloch_Torrey_NoTime3D.cpp",
     [("bloch_torrey_notime3d_cell_integral_0_0", "cell"),
      ("bloch_torrey_notime3d_exterior_facet_integral_0_0", "ext"),
      ("bloch_torrey_notime3d_interior_facet_integral_0_0", "int")]),
    ("tc_sig", "comri/two-comp/hpc-fenics-cpp/ufc/Comp_Sig3D.cpp",
     [("comp_sig3d_cell_integral_0_0", "cell")]),
]

WSTRIDE = 32  # doubles reserved per coefficient in the flat `w` passed by the wrappers

BASE = {"cell": "ufc::cell_integral", "ext": "ufc::exterior_facet_integral",
        "int": "ufc::interior_facet_integral"}
DECL = {
    "cell": "void tabulate_tensor(double* A, const double * const * w, const ufc::cell& c) const;",
    "ext": "void tabulate_tensor(double* A, const double * const * w, const ufc::cell& c, unsigned int facet) const;",
    "int": "void tabulate_tensor(double* A, const double * const * w, const ufc::cell& c0, const ufc::cell& c1, unsigned int facet0, unsigned int facet1) const;",
}


def lift_definition(text, cls):
    """Return the text of the first `void cls::tabulate_tensor(...) const {...}`."""
    m = re.search(r"void\s+%s::tabulate_tensor\s*\(" % re.escape(cls), text)
    if not m:
        raise RuntimeError("tabulate_tensor of %s not found" % cls)
    start = m.start()
    i = text.index("{", m.end())
    depth = 0
    while True:
        ch = text[i]
        if ch == "{":
            depth += 1
        elif ch == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    return text[start:i + 1]


def wrapper(tag, cls, kind):
    name = "ufcref_%s_%s" % (tag, cls)
    pre = ("  const double* wp[16]; for (int i = 0; i < nw; ++i) wp[i] = w + i * %d;\n" % WSTRIDE)
    cellsetup = ("  double* %(c)sx[4]; for (int i = 0; i < 4; ++i) %(c)sx[i] = const_cast<double*>(%(x)s) + 3 * i;\n"
                 "  ufc::cell %(c)s; %(c)s.coordinates = %(c)sx;\n")
    if kind == "cell":
        return ('extern "C" void %s(double* A, const double* w, int nw, const double* x) {\n' % name
                + pre + cellsetup % {"c": "c", "x": "x"}
                + "  %s::%s k; k.tabulate_tensor(A, wp, c);\n}\n" % (tag, cls))
    if kind == "ext":
        return ('extern "C" void %s(double* A, const double* w, int nw, const double* x, int facet) {\n' % name
                + pre + cellsetup % {"c": "c", "x": "x"}
                + "  %s::%s k; k.tabulate_tensor(A, wp, c, (unsigned)facet);\n}\n" % (tag, cls))
    return ('extern "C" void %s(double* A, const double* w, int nw, const double* x0, const double* x1, int f0, int f1) {\n' % name
            + pre + cellsetup % {"c": "c0", "x": "x0"} + cellsetup % {"c": "c1", "x": "x1"}
            + "  %s::%s k; k.tabulate_tensor(A, wp, c0, c1, (unsigned)f0, (unsigned)f1);\n}\n" % (tag, cls))


def build(force=False, verbose=False):
    """Build oracle/_ref/libufcref.so if the reference is present.  Returns its path or None."""
    if os.path.exists(OUT_SO) and not force:
        return OUT_SO
    if not os.path.isdir(REF_ROOT):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    parts = ["#include <cmath>\n#include <stdexcept>\n#include <ufc.h>\n"]
    wraps = []
    for tag, rel, classes in UNITS:
        with open(os.path.join(REF_ROOT, rel), "r", errors="replace") as f:
            text = f.read()
        parts.append("namespace %s {\n" % tag)
        for cls, kind in classes:
            parts.append("struct %s : public %s { %s };\n" % (cls, BASE[kind], DECL[kind]))
            parts.append(lift_definition(text, cls) + "\n")
            wraps.append(wrapper(tag, cls, kind))
        parts.append("}\n")
    parts.extend(wraps)
    parts.append('extern "C" int ufcref_wstride() { return %d; }\n' % WSTRIDE)
    fd, scratch = tempfile.mkstemp(suffix=".cpp", dir=OUT_DIR)
    try:
        with os.fdopen(fd, "w") as f:
            f.write("".join(parts))
        cmd = ["g++", "-O2", "-fPIC", "-shared", "-w", "-I", os.path.join(HERE, "ufc_shim"),
               scratch, "-o", OUT_SO]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    finally:
        os.unlink(scratch)
    return OUT_SO


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose=True)
    print("built" if p else "reference not present; nothing built", p or "")
