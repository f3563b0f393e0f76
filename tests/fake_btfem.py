"""TEST DOUBLE -- never imported by the package.  A stand-in for dmri_fem_cloud_b200.btfem.BTFem backed by the
oracle, so that the `-m "not gpu"` suite can drive the HOST layers (dmrifemlib, cli, comri, sweep: flag parsing,
sequence scalars, call order, result lines) end to end on a box without a GPU.  It proves nothing about the CUDA
path: the parity tests proper are the `-m gpu` ones, which call libbtfem through the C-ABI."""
import numpy as np

import bt_oracle as orc


class FakeBTFem:
    def __init__(self, device=0, lib=None):
        self.device = device
        self.D, self.invT2, self.kappa, self.kmarker = 1.0, 0.0, 0.0, None
        self.periodic = None
        self.vmaster = None
        self.ic = None
        self.h2d_bytes = 0
        self.calls = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self.calls.append("close")

    def set_mesh(self, xyz, tets, phase=None):
        self.xyz = np.asarray(xyz, dtype=float)
        self.tets = np.asarray(tets)
        self.phase = None if phase is None else np.asarray(phase).astype(np.int32)
        self.nv, self.nc = len(self.xyz), len(self.tets)
        self.two_comp = phase is not None
        self.calls.append("set_mesh")

    def set_phase(self, phase=None):
        self.phase = None if phase is None else np.asarray(phase).astype(np.int32)
        self.two_comp = phase is not None

    def mesh_stats(self):
        _, _, hmin, hmax = orc.domain_sizes(orc.as_xyz3(self.xyz), self.tets)
        return hmin, hmax

    def bbox(self):
        x3 = orc.as_xyz3(self.xyz)
        return x3.min(axis=0), x3.max(axis=0)

    def set_diffusion(self, D):
        self.D = D

    def set_relaxation(self, inv_t2):
        self.invT2 = inv_t2

    def set_permeability(self, kappa, marker=None):
        self.kappa, self.kmarker = kappa, marker

    def set_periodic(self, pdir, kappa_e, tol, lo, hi):
        self.periodic = (list(pdir), np.asarray(lo, float), np.asarray(hi, float)) if sum(pdir) > 0 else None

    def set_periodic_map(self, vmaster=None):
        self.vmaster = None if vmaster is None else np.asarray(vmaster)

    def boundary_facets(self):
        return None

    def set_periodic_gather(self, *a):
        pass

    def set_initial(self, ic=None):
        self.ic = None if ic is None else np.asarray(ic, dtype=float)

    def assemble(self):
        kw = {}
        if self.two_comp:
            if self.kmarker is not None:
                kt, mk = np.asarray(self.kappa, float), np.asarray(self.kmarker)
                kw["kappa_facet"] = lambda fv, c0, c1: kt[mk[c0], mk[c1]]
            else:
                kw["kappa"] = float(self.kappa)
        if self.periodic is not None:
            pdir, lo, hi = self.periodic
            hmin, _ = self.mesh_stats()
            kw["bnd_kappa_vertex"] = orc.periodic_marker(orc.as_xyz3(self.xyz), pdir, lo, hi, hmin)
        if self.vmaster is not None:
            kw["vmaster"] = self.vmaster
        self.ops = orc.assemble(self.xyz, self.tets, self.phase, D=self.D, invT2=self.invT2, **kw)
        self.ndof, self.nnz = self.ops.ndof, self.ops.nnz
        self.calls.append("assemble")

    def dofmap(self):
        return self.ops.dof_vertex, self.ops.dof_comp

    def _loop(self, dt, theta, cA, cb, g, q, Fb):
        ops = self.ops
        g = np.asarray(g, dtype=float)
        Jg = (g[0] * ops.Jx + g[1] * ops.Jy + g[2] * ops.Jz).tocsr()
        K0 = ops.S + ops.R + ops.I
        P = (ops.M / dt + theta * (K0 + ops.B)).tocsc()
        Q = (ops.M / dt - (1.0 - theta) * K0).tocsr()
        ic = np.ones(ops.ndof) if self.ic is None else self.ic[ops.dof_vertex]
        per = None
        if self.periodic is not None:
            pdir, lo, hi = self.periodic
            per = orc.periodic_term(orc.as_xyz3(self.xyz), self.tets, ops, pdir, lo, hi, q, g, theta)
        import scipy.sparse.linalg as spla
        u = ic.astype(complex)
        lus = {}
        if self.vmaster is not None:          # transformed equation: cA = q F(t_n), cb = q F(t_{n-1}), theta on both sides
            W, G = orc.strong_operators(ops, g)
            for n in range(len(cA)):
                b = (ops.M / dt - theta * (K0 + cb[n] ** 2 * W)) @ u - 1j * theta * cb[n] * (G @ u)
                if cA[n] not in lus:
                    lus[cA[n]] = spla.splu((ops.M / dt + theta * (K0 + cA[n] ** 2 * W) + 1j * theta * cA[n] * G).tocsc())
                u = lus[cA[n]].solve(b)
            cA = []
        for n in range(len(cA)):
            b = Q @ u - 1j * (1.0 - theta) * cb[n] * (Jg @ u)
            if per is not None:
                b = b + per(u, Fb[n])
            if cA[n] not in lus:
                lus[cA[n]] = spla.splu((P + 1j * theta * cA[n] * Jg).tocsc())
            u = lus[cA[n]].solve(b)
        self.u = u
        comp = ops.dof_comp
        sc = tuple(float(ops.lumped[comp == c] @ u.real[comp == c]) for c in (0, 1))
        vc = tuple(float(ops.lumped[comp == c] @ ic[comp == c]) for c in (0, 1))
        return dict(signal=float(ops.lumped @ u.real), signal_comp=sc, voi=float(ops.lumped @ ic), voi_comp=vc,
                    whole_vol=float(ops.lumped.sum()), loop_ms=0.0, setup_ms=0.0, total_iters=0, max_iters=0,
                    n_spmv=0, n_kernels=0, last_reason=2, n_steps=len(cA))

    def solve(self, dt, theta, cA, cb, gdir, q=0.0, Fb=None, **kw):
        self.calls.append(("solve", kw.get("ksp", "bicgstab"), kw.get("pc", "jacobi")))
        return self._loop(dt, theta, np.asarray(cA, float), np.asarray(cb, float), gdir, q, Fb)

    def solve_batch(self, dt, theta, members, **kw):
        self.calls.append(("solve_batch", len(members)))
        return [self._loop(dt, theta, np.asarray(cA, float), np.asarray(cb, float), g, 0.0, None)
                for cA, cb, g in members]

    def solution(self):
        return self.u
