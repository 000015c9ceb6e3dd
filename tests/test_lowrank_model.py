"""CPU model of the low-rank per-frame PS_MMSE of csrc/wifi_lowrank.cu: the push-through form
    H = U (sigma2 L^-1 + U^H diag(|x|^2) U)^-1 U^H (conj(x) (.) rx),      R = U L U^H (eigenvalues below 64 eps l_max dropped)
against the oracle's long-double 53 x 53 solve of  R (R + sigma2 diag(1/|x|^2))^-1 (rx/tx)  (WiFi_channel_estimation_PS_MMSE.m:16-33).
Pins the algebra, the rank gate and the table layout the kernel relies on; the kernel itself is checked in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import synth
from synth import rel_err

NSC = synth.NSC


def lowrank_tables(R):
    """What wifi_mmse_lowrank_prepare builds: eigen pairs of R above the rounding level, largest first."""
    lam, V = np.linalg.eigh(R)
    keep = lam > 64 * np.finfo(float).eps * lam.max()
    order = np.argsort(-lam[keep])
    return V[:, keep][:, order], lam[keep][order]


def lowrank_mmse(U, lam, tx, rx, s2):
    a = np.conj(tx) * rx                                         # |x|^2 rx/x
    m = np.abs(tx) ** 2
    t = a @ U.conj()                                             # [n][r] = U^H a
    P = U.conj()[:, :, None] * U[:, None, :]                     # P_k,ij = conj(U_ki) U_kj
    G = np.einsum("nk,kij->nij", m, P)
    S = G + s2[:, None, None] * np.diag(1.0 / lam)[None]
    w = np.linalg.solve(S, t[:, :, None])[:, :, 0]
    return w @ U.T                                               # H = U w


@pytest.fixture(scope="module")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.mark.parametrize("taps", [1, 4, 7, 8])
@pytest.mark.parametrize("qam", [False, True])
def test_push_through_form_equals_the_53x53_solve(oracle, taps, qam):
    n = 24
    fr = synth.make_frames(n, seed=900 + taps, sigma2="perframe")
    tx, rx, s2 = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy(), fr["sigma2"]
    if qam:                                                      # per-frame, per-bin moduli (x9 span)
        rng = np.random.default_rng(taps)
        lv = np.array([-3, -1, 1, 3]) / np.sqrt(10.0)
        q = lv[rng.integers(0, 4, tx.shape)] + 1j * lv[rng.integers(0, 4, tx.shape)]
        tx, rx = tx * q, rx * q
    R = synth.channel_covariance(taps)
    U, lam = lowrank_tables(R)
    assert U.shape[1] == taps                                    # the rank gate sees exactly the channel taps
    got = lowrank_mmse(U, lam, tx, rx, s2)
    assert rel_err(got, oracle.mmse_perframe(R, tx, rx, s2)) < 1e-10
    # the r x r system is well conditioned (the 53 x 53 one: ~1e7): FP32 arithmetic is enough for the 1e-4 bound
    m = np.abs(tx) ** 2
    S = np.einsum("nk,ki,kj->nij", m, U.conj(), U) + s2[:, None, None] * np.diag(1.0 / lam)[None]
    assert np.linalg.cond(S).max() < 50


def test_full_rank_covariance_is_not_low_rank():
    U, _ = lowrank_tables(synth.random_hpd(np.random.default_rng(1)))
    assert U.shape[1] == NSC
    # rounded to complex64 a rank-4 covariance is no longer rank-deficient: the operands must be formed from the FP64 matrix
    U32, _ = lowrank_tables(synth.channel_covariance().astype(np.complex64).astype(np.complex128))
    assert U32.shape[1] > 8
