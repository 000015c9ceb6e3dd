"""CPU model of the blocked in-place Gauss-Jordan inverse of csrc/wifi_inverse_tc.cu (cinverse_warp_kernel): NB scalar steps with
implicit partial pivoting composed into one rank-NB update  A <- A - C' R~  (C' = the multiplier vectors + e_(r_s), R~ = the
pivot rows after a unit-lower-triangular transform), panel columns replaced by the factored panel, result un-permuted at the
end.  Checks the algebra the kernel relies on against numpy.linalg.inv (replaces inverse(), utils.c:141-170)."""
import numpy as np
import pytest


def blocked_gj_inverse(A, NB=4, fused_pivot_row=False):
    n0 = A.shape[0]
    N = -(-n0 // NB) * NB
    a = np.zeros((N, N), complex); a[:n0, :n0] = A
    for i in range(n0, N):
        a[i, i] = 1.0                                   # padding rows / columns pivot on themselves
    used = np.zeros(N, bool); rowof = np.zeros(N, int); kof = np.zeros(N, int)
    for K in range(0, N, NB):
        P = a[:, K:K + NB].copy()                       # the panel, factored alone (warp 0 of the kernel)
        Cp = np.zeros((N, NB), complex); rs = []
        for s in range(NB):
            cand = np.where(~used)[0]
            r = cand[np.argmax(np.abs(P[cand, s]))]
            inv = 1.0 / P[r, s]
            if fused_pivot_row:
                # cinverse_warp_kernel (csrc/wifi_inverse_tc.cu): C'_(r_s) = (p - 1) / p makes the pivot row an ordinary row --
                # rho_u - ((p - 1) / p) rho_u = rho_u / p -- and only the pivot column itself is special
                t = P[:, s].copy(); t[r] -= 1.0
                c = t * inv
                Cp[:, s] = c
                rho = P[r, :].copy()
                for u in range(NB):
                    if u != s:
                        P[:, u] -= c * rho[u]
                P[:, s] = -c; P[r, s] = inv
            else:
                c = P[:, s] * inv; c[r] = -inv
                Cp[:, s] = c; Cp[r, s] += 1.0               # C' = c + e_(r_s)
                rho = P[r, :].copy(); rho[s] = 1.0
                P[r, :] = 0; P[:, s] = 0
                P -= np.outer(c, rho)
            used[r] = True; rowof[K + s] = r; kof[r] = K + s; rs.append(r)
        rho = np.zeros((NB, N), complex)
        for s in range(NB):                             # rho^(s) = a[r_s] - sum_{t<s} c^(t)[r_s] rho^(t); c^(t)[r_s] = C'[r_s][t] for t < s
            rho[s] = a[rs[s], :] - sum(Cp[rs[s], t] * rho[t] for t in range(s))
        a -= Cp @ rho
        a[:, K:K + NB] = P
    Y = np.zeros((N, N), complex)
    Y[np.ix_(kof, rowof)] = a
    return Y[:n0, :n0]


@pytest.mark.parametrize("n", [33, 47, 53, 56, 64])
@pytest.mark.parametrize("NB", [2, 4])
@pytest.mark.parametrize("fused", [False, True])
def test_blocked_gauss_jordan_model(n, NB, fused):
    rng = np.random.default_rng(n * 10 + NB)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    Y = blocked_gj_inverse(A, NB, fused)
    ref = np.linalg.inv(A)
    assert np.abs(Y - ref).max() / np.abs(ref).max() < 1e-12
    B = A.copy(); B[0, 0] = 0                          # a zero leading element: un-pivoted elimination (the reference's determinant) fails here
    assert np.abs(blocked_gj_inverse(B, NB, fused) @ B - np.eye(n)).max() < 1e-11


def lookahead_gj_inverse(A, NB, OB):
    """The outer-block form of cinverse_warp_kernel (FP64: NB = 2, OB = 4): the inner panels of an outer block are factored one
    after the other, each applying its rank-NB update to the 8-wide TILE COLUMN that holds the block only; every other column
    takes the rank-OB update of the whole outer block in one pass.  Pivot row s of the block is transformed with all earlier
    pivots of the block outside the tile column, and with the pivots of its own inner panel inside it (the earlier panels'
    updates are already in the matrix there)."""
    n0 = A.shape[0]
    N = -(-n0 // 8) * 8
    a = np.zeros((N, N), complex); a[:n0, :n0] = A
    for i in range(n0, N):
        a[i, i] = 1.0
    used = np.zeros(N, bool); rowof = np.zeros(N, int); kof = np.zeros(N, int)
    for K0 in range(0, N, OB):
        Cp = np.zeros((N, OB), complex); rho = np.zeros((OB, N), complex); rs = []
        in0 = np.zeros(N, bool); in0[8 * (K0 // 8):8 * (K0 // 8) + 8] = True
        for p in range(OB // NB):
            Kp = K0 + p * NB
            P = a[:, Kp:Kp + NB].copy()
            for s in range(NB):
                cand = np.where(~used)[0]
                r = cand[np.argmax(np.abs(P[cand, s]))]
                inv = 1.0 / P[r, s]
                t = P[:, s].copy(); t[r] -= 1.0
                c = t * inv
                Cp[:, p * NB + s] = c
                prow = P[r, :].copy()
                for u in range(NB):
                    if u != s:
                        P[:, u] -= c * prow[u]
                P[:, s] = -c; P[r, s] = inv
                used[r] = True; rowof[Kp + s] = r; kof[r] = Kp + s; rs.append(r)
            for s in range(NB):
                S = p * NB + s
                v = a[rs[S], :].copy()
                for t in range(S):
                    m = np.ones(N, bool)
                    if t < p * NB:
                        m[in0] = False                  # already applied to the tile column by the narrow updates
                    v[m] -= Cp[rs[S], t] * rho[t][m]
                rho[S] = v
            a[:, in0] -= Cp[:, p * NB:(p + 1) * NB] @ rho[p * NB:(p + 1) * NB][:, in0]
            a[:, Kp:Kp + NB] = P
        a[:, ~in0] -= Cp @ rho[:, ~in0]
    Y = np.zeros((N, N), complex)
    Y[np.ix_(kof, rowof)] = a
    return Y[:n0, :n0]


@pytest.mark.parametrize("n", [33, 40, 53, 64])
@pytest.mark.parametrize("NB,OB", [(2, 4), (4, 8), (2, 8)])
def test_lookahead_gauss_jordan_model(n, NB, OB):
    rng = np.random.default_rng(n * 10 + NB + OB)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    ref = np.linalg.inv(A)
    assert np.abs(lookahead_gj_inverse(A, NB, OB) - ref).max() / np.abs(ref).max() < 1e-12
