"""Pin the CPU oracle (oracle/wifi_oracle.c) against the reference's own outputs:
tests/golden/*.npz were produced by the reference's sequential C code compiled in place
(tests/golden/make_golden.py) and from its matlab.mat.  CPU only."""
import numpy as np
import pytest

import synth
from synth import rel_err

NSC, NBLK = 53, 15


def test_inputs_h_shape(gold):
    g = gold["inputs_h"]
    assert g["tx_symb"].shape == (795,) and g["rx_preamble_fft"].shape == (53,)
    assert float(g["ow2"]) == 9.6172e-08                                  # inputs.h:18
    assert np.allclose(np.sign(g["tx_preamble_fft"].real[np.arange(53) != 26]), synth.LTS[np.arange(53) != 26])


def test_estimators_match_reference_on_inputs_h(oracle, gold):
    g, r = gold["inputs_h"], gold["ref_c_outputs"]
    assert rel_err(oracle.lt_ls(g["tx_preamble_fft"], g["rx_preamble_fft"]), r["lt_ls"]) < 1e-15
    tx = g["tx_symb"].reshape(NBLK, NSC); rx = g["rx_symb"].reshape(NBLK, NSC)
    for name in ("ps_linear", "ps_cubic", "ps_sinc"):
        got = getattr(oracle, name)(tx, rx)
        assert rel_err(got, r[name + "_blocks"]) < 1e-15, name
    # whole-frame stride, block 0 only (main.c:30-33)
    got = oracle.ps_linear(g["tx_symb"], g["rx_symb"], frame_stride=795)
    assert rel_err(got[0], r["ps_linear_blocks"][0]) < 1e-15


def test_survey_appendix_c_spot_values(oracle, gold):
    """SURVEY App. C (sequential C on inputs.h block 0, printed to 17 digits)."""
    r = gold["ref_c_outputs"]
    assert abs(r["lt_ls"][0] - (0.0090364160807643769 + 0.00092392284117541026j)) < 1e-17
    assert r["lt_ls"][26] == 0
    assert abs(r["ps_linear_blocks"][0][52] - (0.00051507313233677404 - 0.012534787556476966j)) < 1e-17
    assert abs(r["ps_cubic_blocks"][0][47] - (-0.0082556464525039664 - 0.0075616473135333958j)) < 1e-17
    assert abs(r["ps_sinc_blocks"][0][26] - (0.0073525116590952601 - 0.0050235452592959683j)) < 1e-17


def test_estimators_match_reference_on_synthetic(oracle, gold):
    r = gold["ref_c_outputs"]
    assert rel_err(oracle.lt_ls(r["syn_tx_pre"], r["syn_rx_pre"]), r["syn_lt_ls"]) < 1e-15
    for name in ("ps_linear", "ps_cubic", "ps_sinc"):
        assert rel_err(getattr(oracle, name)(r["syn_tx_blk0"], r["syn_rx_blk0"]), r["syn_" + name]) < 1e-15, name


def test_lt_ls_nan_when_re_equals_im(oracle, gold):
    r, g = gold["ref_c_outputs"], gold["inputs_h"]
    got = oracle.lt_ls(r["nan_tx_pre"], g["rx_preamble_fft"])
    assert np.isnan(r["nan_lt_ls"][3]) and np.isnan(got[3])
    assert rel_err(got, r["nan_lt_ls"]) < 1e-15


def test_utils_match_reference(oracle, gold):
    u = gold["ref_utils"]
    assert rel_err(oracle.multiply(u["A53"], u["B53"]), u["mul_53x53"]) < 1e-15
    assert rel_err(oracle.multiply(u["A53"], u["v53"]), u["mul_53x1"]) < 1e-15
    assert rel_err(oracle.multiply(u["A7x5"], u["B5x3"]), u["mul_7x5x3"]) < 1e-15
    assert np.array_equal(oracle.hermitian_as_written(u["A53"]), u["herm_53"])
    assert np.array_equal(oracle.hermitian_as_written(u["A7x5"]), u["herm_7x5"])
    assert np.all(u["herm_53"].imag == 0)                                 # utils.c:6 is real-valued (sic)
    assert rel_err(oracle.outer(u["A53"], u["B53"]), u["outer_53"]) < 1e-15
    assert np.array_equal(oracle.identity(53, 9.6172e-08), u["ident_53"])
    assert np.array_equal(oracle.addition_as_written(u["A53"], u["B53"]), u["add_53"])
    assert np.array_equal(u["add_53"], 2 * u["A53"])                      # utils.c:117 ignores M2 (sic)
    assert np.all(np.isnan(u["mul_mismatch"]))                            # utils.c:18-19: prints, writes nothing
    with pytest.raises(ValueError):
        oracle.multiply(u["A7x5"], u["A7x5"])


def test_inverse_cofactor_matches_reference(oracle, gold):
    u = gold["ref_utils"]
    for n in (2, 3, 6, 10):
        got = oracle.inverse_cofactor(u[f"inv_in_{n}"])
        assert rel_err(got, u[f"inv_out_{n}"]) < 1e-15, n
        assert rel_err(oracle.inverse_gj(u[f"inv_in_{n}"]), u[f"inv_out_{n}"]) < 1e-12, n


def test_inverse_53_pd_vs_reference(oracle, gold):
    u = gold["ref_utils"]
    a = u["inv_in_53pd"]
    y = oracle.inverse_gj(a)
    # the reference's un-pivoted cofactor inverse is itself only ~1e-9 accurate at this
    # conditioning (cond ~ 1e6); both must invert A
    assert np.abs(y @ a - np.eye(53)).max() < 1e-9
    assert np.abs(u["inv_out_53pd"] @ a - np.eye(53)).max() < 1e-6
    assert rel_err(y, u["inv_out_53pd"], floor=1e-2) < 1e-6


def test_matlab_goldens(oracle, gold):
    m = gold["matlab_mat"]
    tx = m["tx_symb"].T.reshape(1, NBLK, NSC); rx = m["rx_symb"].T.reshape(1, NBLK, NSC)   # MATLAB 53x15 -> [15][53]
    lt = oracle.lt_ls(m["tx_preamble_fft"].ravel(), m["rx_preamble_fft"].ravel())
    assert rel_err(lt, m["H_EST_LT_LS"].ravel()) < 1e-14
    for which, key in (("linear", "H_EST_PS_Linear"), ("cubic", "H_EST_PS_Cubic"), ("sinc", "H_EST_PS_Sinc")):
        assert rel_err(oracle.ps_matlab(which, tx, rx)[0], m[key].ravel()) < 1e-13, which
    eq = oracle.equalize(rx, m["H_EST_LT_LS"].ravel(), m["H_EST_PS_Linear"].ravel())
    assert rel_err(eq[0], m["eq_symbols"].T, floor=1e-6) < 1e-13
    assert np.all(eq[0][:, 26] == 0)


def test_frontend_matlab_goldens(oracle, gold):
    """CP strip + 64-point FFT + circshift 26 + keep 53 (WiFi_blocks_extraction.m), LTS averaging and the noise
    estimate (WiFi_RX.m:19-31) against the workspace the reference saved in matlab.mat."""
    m, g = gold["matlab_mat"], gold["inputs_h"]
    for side in ("tx", "rx"):
        symb, pre, ow2 = oracle.frontend(m[side + "_packet"].ravel(), m[side + "_lptot"].ravel())
        # the DC bin of tx (-1e-4 against 8.875) is a cancellation residue of the FFT: floor 1e-3 like the estimators
        assert rel_err(symb[0], m[side + "_symb"].T) < 1e-12, side
        assert rel_err(pre[0], m[side + "_preamble_fft"].ravel()) < 1e-12, side
    assert abs(ow2[0] / float(g["ow2"]) - 1) < 1e-4          # inputs.h:18 prints ow2 to 5 significant digits


def test_inputs_h_is_rounded_matlab_frame(gold):
    """layout check: tx_symb[53*b + k] (inputs.h) == MATLAB tx_symb(k+1, b+1) to 4 decimals"""
    m, g = gold["matlab_mat"], gold["inputs_h"]
    assert np.abs(g["tx_symb"].reshape(NBLK, NSC) - m["tx_symb"].T).max() < 1e-4
    assert np.abs(g["rx_symb"].reshape(NBLK, NSC) - m["rx_symb"].T).max() < 1e-4


def test_mmse_known_answers(oracle, gold):
    g, r, k = gold["inputs_h"], gold["ref_c_outputs"], gold["mmse_kat"]
    tx0, rx0 = g["tx_symb"][:53], g["rx_symb"][:53]
    full = oracle.mmse_cconv(tx0, rx0, float(g["ow2"]), r["lt_ls"])
    assert rel_err(full, k["inputs_h_full"]) < 1e-13            # 40-digit mpmath, full 53x53 solve
    assert rel_err(oracle.mmse_rank1(tx0, rx0, float(g["ow2"]), r["lt_ls"]), k["inputs_h_rank1"]) < 1e-14
    assert rel_err(k["inputs_h_full"], k["inputs_h_rank1"]) < 1e-14       # the two mpmath routes agree
    # g = 1.0058770039341261847 - 0.00013887971098788543i, H[0], H[26]
    assert abs(complex(k["inputs_h_g"]) - (1.0058770039341261847 - 0.00013887971098788543j)) < 1e-15
    assert abs(k["inputs_h_full"][0] - (0.0090896514477585877939 + 0.00092809776449416434917j)) < 1e-16
    assert full[26] == 0
    # the MATLAB text (explicit F, unconjugated X) agrees with the north-star form on this frame
    assert rel_err(oracle.mmse_matlab_block(tx0, rx0, float(g["ow2"]), r["lt_ls"]), full) < 1e-9
    # general R, per-frame sigma
    got = oracle.mmse_perframe(k["gen_R"], k["gen_tx"], k["gen_rx"], k["gen_sigma2"])
    assert rel_err(got, k["gen_H"]) < 1e-13


def test_mmse_pinned_by_the_references_own_routines(oracle, gold):
    """PS_MMSE composed from the reference's own COMPILED multiply / multiplyVxVeqM / identity / inverse (utils.c:16-31,55-65,
    84-93,141-170) as WiFi_channel_estimation_PS_MMSE.m:25-32 prescribes -- only the conjugate transposes and the M1 + M2 are the
    generating script's, because utils.c:6,117 are defective as written (tests/golden/make_golden.py mmse_composed).
    (i) a full-rank covariance at 37 dB: the oracle's per-frame solve and its shared filter equal the reference-routine result to
    1e-12 -- this pins the FORMULA; (ii) the inputs.h frame in main.c:148's convention (rank-one R, cond(Ryy) 4e6): both the .m
    text and the north-star form, to the 5e-9 the reference's un-pivoted cofactor inverse itself reaches there
    (|Ryy^-1 Ryy - I| = 1.7e-9 is stored next to the vectors)."""
    c, g = gold["mmse_ref_composed"], gold["inputs_h"]
    s2 = float(c["gen_sigma2"])
    tx, rx = c["gen_tx"][None], c["gen_rx"][None]
    assert float(c["gen_residual"]) < 1e-12
    assert rel_err(oracle.mmse_perframe(c["gen_R"], tx, rx, np.array([s2]))[0], c["gen_H"]) < 1e-12
    W = oracle.mmse_filter(c["gen_R"], s2 / np.abs(c["gen_tx"]) ** 2)
    assert rel_err(oracle.mmse_apply(W, rx / tx)[0], c["gen_H"]) < 1e-12
    tx0, rx0, ow2 = g["tx_symb"][:53], g["rx_symb"][:53], float(g["ow2"])
    assert rel_err(oracle.mmse_cconv(tx0, rx0, ow2, c["H_ls"]), c["north_star_form"]) < 2e-8
    assert rel_err(oracle.mmse_matlab_block(tx0, rx0, ow2, c["H_ls"]), c["matlab_form"]) < 2e-8
    assert rel_err(c["matlab_form"], c["north_star_form"]) < 2e-8          # the reference's routines agree with each other
    assert rel_err(oracle.mmse_perframe(c["R"], tx0[None], rx0[None], np.array([ow2]))[0], c["north_star_form"]) < 2e-8


def test_mmse_shared_filter_equals_perframe(oracle):
    fr = synth.make_frames(6, seed=3)
    R = synth.channel_covariance()
    tx, rx = fr["tx_symb"][:, 0, :], fr["rx_symb"][:, 0, :]
    # shared |x|^2 (BPSK) and sigma -> one filter W for all frames
    d = synth.OW2 / np.abs(tx[0]) ** 2
    assert np.allclose(np.abs(tx) ** 2, np.abs(tx[0]) ** 2)
    W = oracle.mmse_filter(R, d)
    assert rel_err(oracle.mmse_apply(W, rx / tx), oracle.mmse_perframe(R, tx, rx, synth.OW2)) < 1e-10


def test_reference_mmse_as_written_is_nan_documented():
    """main.c:148-212 as written returns NaN (278 s/frame): not run here; the claim is pinned by
    the chain utils.c:117 (Ryy = 2 ow2 I) -> un-pivoted minors of a diagonal matrix -> 0/0."""
    from oracle.pyoracle import Oracle
    o = Oracle()
    ryy = o.addition_as_written(o.identity(6, 9.6172e-08), np.ones((6, 6), complex))
    assert np.array_equal(ryy, 2 * o.identity(6, 9.6172e-08))
    inv = o.inverse_cofactor(ryy)
    assert np.isnan(inv).any()


def _inplace_gauss_jordan(A):
    """numpy restatement of cinverse_reg_kernel (csrc/wifi_solve.cu): in-place Gauss-Jordan with IMPLICIT partial pivoting --
    no row is ever moved; step k picks its pivot row r among the rows not used yet with the kernel's 32-bit key (|a|^2 as
    float32 bits with the low 6 bits dropped, then 63 - row so that ties go to the lowest row), scales the column by 1/pivot
    (c_r = -1/pivot), publishes row r with entry k := 1, clears both and updates every entry; the result is un-permuted on
    the way out, Y[kof[i]][rowof[j]] = a[i][j]."""
    a = np.array(A, np.complex128)
    n = a.shape[0]
    used = np.zeros(n, bool)
    rowof, kof = np.zeros(n, int), np.zeros(n, int)
    for k in range(n):
        v = (np.abs(a[:, k]) ** 2).astype(np.float32).view(np.uint32).astype(np.int64)
        key = np.where(used, -1, ((v & 0x7FFFFFC0) | (63 - np.arange(n))))
        r = int(np.argmax(key))
        inv = 1.0 / a[r, k]
        c = a[:, k] * inv
        c[r] = -inv
        row = a[r, :].copy()
        row[k] = 1.0
        a[:, k] = 0
        a[r, :] = 0
        a -= np.outer(c, row)
        used[r] = True
        rowof[k], kof[r] = r, k
    Y = np.empty_like(a)
    Y[np.ix_(kof, rowof)] = a
    return Y, rowof


@pytest.mark.parametrize("n", [33, 53, 64])
def test_inplace_gauss_jordan_restatement(oracle, n):
    """The algorithm of the register-resident batched inverse (the replacement of utils.c:141-170 for orders 33..64), restated
    on the CPU, against the oracle's long-double pivoted inverse: pins the implicit-pivoting bookkeeping (rowof / kof) and the
    quantised pivot key on matrices whose pivot order is a non-trivial permutation."""
    rng = np.random.default_rng(n)
    A = 3 * np.roll(np.eye(n), 1 + n // 3, axis=1) + (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n)
    Y, rowof = _inplace_gauss_jordan(A)
    assert sorted(rowof) == list(range(n)) and not np.array_equal(rowof, np.arange(n))
    assert rel_err(Y, oracle.inverse_gj(A), floor=1e-2) < 1e-12
    assert np.abs(Y @ A - np.eye(n)).max() < 1e-13
