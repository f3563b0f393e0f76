"""Pins the oracle's element-level arithmetic against the reference's OWN generated kernels:
(a) the committed golden tensors (tests/golden/ufc_element_tensors.npz, produced by
tests/golden/make_golden.py from oracle/_ref) -- always; (b) oracle/_ref live when it is built."""
import os

import numpy as np
import pytest

import bt_oracle as orc
import ufc_ref
from conftest import GOLDEN

G = np.load(os.path.join(GOLDEN, "ufc_element_tensors.npz"))
TET = np.array([[0, 1, 2, 3]])
RTOL = 1e-12


def close(a, b, tol=RTOL):
    return np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-300)


def blk(A, n, fi, fj):
    """(field fi, field fj) 4x4 block of a blocked mixed-element tensor with n fields."""
    A = A.reshape(4 * n, 4 * n)
    return A[4 * fi:4 * fi + 4, 4 * fj:4 * fj + 4]


def test_one_comp_mass_stiffness_j():
    for c in range(len(G["x"])):
        x = G["x"][c]
        em = orc.element_matrices(x, TET, D=G["K"][c])
        gx = sum(G["g"][c][d] * em[n][0] for d, n in enumerate(("Jx", "Jy", "Jz")))
        M, S, J = G["oc_mass"][c], G["oc_stiff"][c], G["oc_j"][c]
        assert close(blk(M, 2, 0, 0), em["M"][0]) and close(blk(M, 2, 1, 1), em["M"][0])
        assert np.all(blk(M, 2, 0, 1) == 0) and np.all(blk(M, 2, 1, 0) == 0)
        assert close(blk(S, 2, 0, 0), em["S"][0]) and close(blk(S, 2, 1, 1), em["S"][0])
        # j = -GX*(ui*vr - ur*vi): (vr,ui) block = -J, (vi,ur) block = +J  <=>  + i*J in complex form
        assert close(blk(J, 2, 0, 1), -gx) and close(blk(J, 2, 1, 0), gx)
        assert np.max(np.abs(blk(J, 2, 0, 0))) <= 1e-14 * np.max(np.abs(gx))


def test_one_comp_theta_forms_and_signal():
    """The reference's per-step bilinear and linear forms equal P + i*theta*c*Jg and
    (Q - i(1-theta)c Jg) u of the oracle's operator split."""
    th, dt = float(G["theta"]), float(G["dt"])
    for c in range(len(G["x"])):
        x = G["x"][c]
        em = orc.element_matrices(x, TET, D=G["K"][c])
        Jg = sum(G["g"][c][d] * em[n][0] for d, n in enumerate(("Jx", "Jy", "Jz")))
        cc = G["ft"][c] * G["gnorm"][c]
        P = em["M"][0] / dt + th * em["S"][0]
        Q = em["M"][0] / dt - (1 - th) * em["S"][0]
        A = G["oc_a"][c]
        assert close(blk(A, 2, 0, 0), P) and close(blk(A, 2, 1, 1), P)
        assert close(blk(A, 2, 0, 1), -th * cc * Jg) and close(blk(A, 2, 1, 0), th * cc * Jg)
        u = G["u"][c][:4] + 1j * G["u"][c][4:]
        b = Q @ u - 1j * (1 - th) * cc * (Jg @ u)
        assert close(G["oc_L"][c][:4], b.real, 1e-11) and close(G["oc_L"][c][4:], b.imag, 1e-11)
        assert abs(G["oc_sig"][c] - em["M"][0].sum(axis=1) @ u.real) <= 1e-12 * abs(G["oc_sig"][c]) + 1e-15


def test_two_comp_cell_phase_weighting():
    for c in range(len(G["x"])):
        x = G["x"][c]
        em = orc.element_matrices(x, TET, D=G["K"][c])
        Jg = sum(G["g"][c][d] * em[n][0] for d, n in enumerate(("Jx", "Jy", "Jz")))
        full = em["M"][0] + em["S"][0]
        for ph, key in ((0, "tc_cell_ph0"), (1, "tc_cell_ph1")):
            A = G[key][c]
            on, off = (0, 2) if ph == 0 else (2, 0)       # fields (u0r,u0i,u1r,u1i)
            assert close(blk(A, 4, on, on), full) and close(blk(A, 4, on + 1, on + 1), full)
            assert close(blk(A, 4, on, on + 1), -Jg) and close(blk(A, 4, on + 1, on), Jg)
            assert np.all(np.abs(A.reshape(16, 16)[4 * off:4 * off + 8, :]) == 0)
            assert np.all(np.abs(A.reshape(16, 16)[:, 4 * off:4 * off + 8]) == 0)


def _scatter_macro(A, cellA, cellB, nv=5):
    """32x32 macro tensor -> (4*nv)^2 global matrix in the reference layout dof = field*nv + vertex."""
    loc = []
    for cell in (cellA, cellB):
        for f in range(4):
            loc += [f * nv + v for v in cell]
    Gm = np.zeros((4 * nv, 4 * nv))
    A = A.reshape(32, 32)
    for i, gi in enumerate(loc):
        for j, gj in enumerate(loc):
            Gm[gi, gj] += A[i, j]
    return Gm


def test_two_comp_interface_facet():
    """Reference interior-facet integral (kappa*(u0-u1)(v0-v1)*|jump(phase)| on dS) vs the oracle's I."""
    for c in range(len(G["if_points"])):
        pts = G["if_points"][c]
        rec = G["if_record"][c]
        cellA = rec[3:7].astype(int)
        cellB = rec[7:11].astype(int)
        A = rec[11:]
        ref = _scatter_macro(A, cellA, cellB)
        tets = np.array([cellA, cellB])
        ops = orc.assemble(pts, tets, phase=np.array([0, 1]), D=1.0, kappa=G["if_kappa"][c])
        I = ops.I.toarray()
        mine = np.zeros((20, 20))
        for i in range(ops.ndof):
            for j in range(ops.ndof):
                for ri in (0, 1):                          # re and im rows carry the same real operator
                    gi = (2 * ops.dof_comp[i] + ri) * 5 + ops.dof_vertex[i]
                    gj = (2 * ops.dof_comp[j] + ri) * 5 + ops.dof_vertex[j]
                    mine[gi, gj] = I[i, j]
        assert close(mine, ref), c


def test_two_comp_exterior_facet():
    """Reference exterior-facet integral kappa_e/h * u v * phase on ds with P1 kappa_e vs the
    oracle's boundary-mass closed form (Appendix A.6)."""
    faces = orc._FACES
    for c in range(len(G["x"])):
        x = G["x"][c]
        ke = G["ext_kappa_e"][c]
        for ph in (0, 1):
            for facet in range(4):
                A = G["tc_ext"][c][ph * 4 + facet].reshape(16, 16)
                fv = faces[facet]
                area = orc.tri_area(x, fv[None])[0]
                kv = ke[fv]
                sk = kv.sum()
                B = np.zeros((4, 4))
                for a in range(3):
                    for b in range(3):
                        w = (2 * sk + 4 * kv[a]) if a == b else (sk + kv[a] + kv[b])
                        B[fv[a], fv[b]] = area * w / 60.0
                on = 2 * ph
                assert close(A[4 * on:4 * on + 4, 4 * on:4 * on + 4], B, 1e-11)
                assert close(A[4 * on + 4:4 * on + 8, 4 * on + 4:4 * on + 8], B, 1e-11)


@pytest.mark.skipif(not ufc_ref.available(), reason="oracle/_ref not built (reference absent)")
def test_live_reference_kernels_match_golden():
    """The committed golden tensors are what the reference kernels produce today."""
    live = ufc_ref.golden_element_tensors(seed=2024, ncell=24)
    for k in G.files:
        assert np.array_equal(np.asarray(live[k]), G[k]), k


def test_unit_tet_closed_forms():
    x = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    em = orc.element_matrices(x, TET)
    assert abs(em["M"][0][0, 0] - 1 / 60) < 1e-16 and abs(em["M"][0][0, 1] - 1 / 120) < 1e-16
    assert np.allclose(em["S"][0][0], [0.5, -1 / 6, -1 / 6, -1 / 6])
    assert abs(em["Jx"][0][0, 0] - 1 / 360) < 1e-16
