#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE ITSELF.  Run in the build container only
(`python tests/golden/make_golden.py`); needs /root/reference and oracle/_ref (built by
oracle/Makefile from the reference's own main.c/utils.c, compiled in place).  The GPU box
and the test-suite only ever read the committed .npz files.

Files written:
  inputs_h.npz        the inputs.h frame (inputs.h:18-1724) as parsed by the C compiler
  ref_c_outputs.npz   reference sequential C outputs: 4 estimators on the inputs.h frame
                      (preamble / every OFDM block) and on 64 seeded synthetic frames
  ref_utils.npz       reference multiply / hermitian / multiplyVxVeqM / identity / addition /
                      inverse on seeded matrices (incl. one 53x53 Hermitian PD inverse)
  matlab_mat.npz      the arrays of the reference's matlab.mat used as MATLAB-mode goldens
  mmse_kat.npz        intended-MMSE known answers: 40-digit mpmath evaluation (full 53x53
                      solve and rank-1 closed form) on the inputs.h frame + a general-R,
                      per-frame-sigma case.  (The reference has no usable MMSE output.)
  mmse_ref_composed.npz   PS_MMSE of OFDM block 0 of the inputs.h frame computed with the REFERENCE'S OWN COMPILED ROUTINES
                      (multiply, multiplyVxVeqM, identity, inverse: utils.c:16-31,55-65,84-93,141-170), composed as its MATLAB twin
                      prescribes (WiFi_channel_estimation_PS_MMSE.m:25-32); only the conjugate transposes and the M1 + M2 are
                      done by this script, because hermitian() and addition() are defective as written (utils.c:6,117).
                      `python tests/golden/make_golden.py composed` writes this file alone (two cofactor inverses: ~15 s).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.pyoracle import Reference, NSC, NBLK  # noqa: E402
import synth  # noqa: E402


def mmse_composed(ref, inp, H_ls):
    """WiFi_channel_estimation_PS_MMSE.m:16-32 for OFDM block 0 (main.c:30-33), every product and the inverse by the reference."""
    tx0, rx0, ow2 = inp["tx_symb"][:NSC], inp["rx_symb"][:NSC], float(inp["ow2"])
    t = np.arange(NSC)
    F = np.exp(-2j * np.pi * np.outer(t, t) / NSC)                       # .m:17-23 == main.c:22-26
    Fh = F.conj().T                                                      # F' (by the script: utils.c:6 is not a conjugate transpose)
    ht = ref.multiply(Fh / NSC, H_ls.reshape(NSC, 1))                    # ifft(H_EST, 53)                           .m:27
    Rhh = ref.outer(ht, ht.conj().T)                                     # ifft(H_EST) * ifft(H_EST)'  (multiplyVxVeqM, main.c:189)
    X4 = np.diag(tx0)                                                    # .m:29
    FR = ref.multiply(F, Rhh)
    FRFh = ref.multiply(FR, Fh)                                          # F Rhh F'
    Rhy = ref.multiply(ref.multiply(Rhh, Fh), X4)                        # .m:30   (X4, not X4': the .m text as written)
    Ryy = ref.multiply(ref.multiply(X4, FRFh), X4.conj().T) + ref.identity(NSC, ow2)      # .m:31 (M1 + M2 by the script: utils.c:117)
    Ryy_inv = ref.inverse(Ryy)                                           # pinv of a Hermitian PD matrix = its inverse      .m:32
    H_m = ref.multiply(ref.multiply(ref.multiply(F, Rhy), Ryy_inv), rx0.reshape(NSC, 1)).ravel()
    # the north-star form R (R + ow2 (X X^H)^-1)^-1 (rx/tx), R = F Rhh F', with the same routines
    A = FRFh + np.diag(ow2 / np.abs(tx0) ** 2)
    H_ns = ref.multiply(ref.multiply(FRFh, ref.inverse(A)), (rx0 / tx0).reshape(NSC, 1)).ravel()
    # a better-conditioned system (shared full-rank covariance, 37 dB SNR: cond ~ 1e4) pins the formula itself to ~1e-12: on the
    # inputs.h frame the reference's own un-pivoted cofactor inverse of Ryy (cond 4e6, rank-one R) is only good to ~2e-9
    fr = synth.make_frames(1, seed=53, sigma2=2e-4)
    Rg = synth.channel_covariance()
    gtx, grx = fr["tx_symb"][0, 0, :], fr["rx_symb"][0, 0, :]
    Ag = Rg + np.diag(2e-4 / np.abs(gtx) ** 2)
    Ag_inv = ref.inverse(Ag)
    gen_H = ref.multiply(ref.multiply(Rg, Ag_inv), (grx / gtx).reshape(NSC, 1)).ravel()
    return {"H_ls": H_ls, "matlab_form": H_m, "north_star_form": H_ns, "R": FRFh, "Ryy_residual": np.abs(Ryy_inv @ Ryy - np.eye(NSC)).max(),
            "gen_R": Rg, "gen_tx": gtx, "gen_rx": grx, "gen_sigma2": np.array(2e-4), "gen_H": gen_H,
            "gen_residual": np.abs(Ag_inv @ Ag - np.eye(NSC)).max()}


def main():
    ref = Reference()
    if len(sys.argv) > 1 and sys.argv[1] == "composed":
        inp = ref.inputs()
        H_ls = ref.estimate("lt_ls", inp["tx_preamble_fft"], inp["rx_preamble_fft"])
        c = mmse_composed(ref, inp, H_ls)
        np.savez(os.path.join(HERE, "mmse_ref_composed.npz"), **c)
        print("mmse_ref_composed.npz written; |Ryy^-1 Ryy - I| =", c["Ryy_residual"])
        return
    rng = np.random.default_rng(0x80211)

    # ---- 1. inputs.h ----
    inp = ref.inputs()
    np.savez(os.path.join(HERE, "inputs_h.npz"), **inp)
    with open(os.path.join(HERE, "inputs_h_frame.f64"), "wb") as f:      # raw doubles for the C host driver (host/)
        np.array([float(inp["ow2"])]).tofile(f)
        for key in ("tx_preamble_fft", "rx_preamble_fft", "tx_symb", "rx_symb"):
            np.ascontiguousarray(inp[key]).view(np.float64).tofile(f)
    txs = inp["tx_symb"].reshape(NBLK, NSC)
    rxs = inp["rx_symb"].reshape(NBLK, NSC)

    # ---- 2. reference estimators ----
    out = {}
    out["lt_ls"] = ref.estimate("lt_ls", inp["tx_preamble_fft"], inp["rx_preamble_fft"])
    for name in ("ps_linear", "ps_cubic", "ps_sinc"):
        out[name + "_blocks"] = ref.estimate(name, txs, rxs)          # [15][53], row 0 = main.c's block 0
    fr = synth.make_frames(64, seed=0x80211, sigma2=None)
    out["syn_tx_pre"], out["syn_rx_pre"] = fr["tx_pre"], fr["rx_pre"]
    out["syn_tx_blk0"], out["syn_rx_blk0"] = fr["tx_symb"][:, 0, :], fr["rx_symb"][:, 0, :]
    out["syn_lt_ls"] = ref.estimate("lt_ls", fr["tx_pre"], fr["rx_pre"])
    for name in ("ps_linear", "ps_cubic", "ps_sinc"):
        out["syn_" + name] = ref.estimate(name, fr["tx_symb"][:, 0, :], fr["rx_symb"][:, 0, :])
    # LT_LS bug-compat: Re(tx) == Im(tx) -> NaN (main.c:69-72)
    tq = inp["tx_preamble_fft"].copy(); tq[3] = 2.5 + 2.5j
    out["nan_tx_pre"] = tq
    out["nan_lt_ls"] = ref.estimate("lt_ls", tq, inp["rx_preamble_fft"])
    np.savez(os.path.join(HERE, "ref_c_outputs.npz"), **out)

    # ---- 3. reference utils ----
    u = {}

    def crand(*s):
        return rng.standard_normal(s) + 1j * rng.standard_normal(s)

    u["A53"], u["B53"], u["v53"] = crand(53, 53), crand(53, 53), crand(53, 1)
    u["A7x5"], u["B5x3"] = crand(7, 5), crand(5, 3)
    u["mul_53x53"] = ref.multiply(u["A53"], u["B53"])
    u["mul_53x1"] = ref.multiply(u["A53"], u["v53"])
    u["mul_7x5x3"] = ref.multiply(u["A7x5"], u["B5x3"])
    u["herm_53"] = ref.hermitian(u["A53"])
    u["herm_7x5"] = ref.hermitian(u["A7x5"])
    u["outer_53"] = ref.outer(u["A53"], u["B53"])          # multiplyVxVeqM: col 0 of M1 x row 0 of M2 (main.c:189 passes 53x53 dims)
    u["mul_mismatch"] = ref.multiply(u["A7x5"], u["A7x5"])  # col1 != row2: prints, writes nothing (utils.c:18-19) -> stays NaN-filled
    u["ident_53"] = ref.identity(53, 9.6172e-08)
    u["add_53"] = ref.addition(u["A53"], u["B53"])
    for n in (2, 3, 6, 10):
        a = crand(n, n) + n * np.eye(n)
        u[f"inv_in_{n}"], u[f"inv_out_{n}"] = a, ref.inverse(a)
    R = synth.channel_covariance()
    pd = R + np.diag(np.full(53, 9.6172e-08 / 8.875 ** 2))
    u["inv_in_53pd"], u["inv_out_53pd"] = pd, ref.inverse(pd)          # ~7 s
    np.savez(os.path.join(HERE, "ref_utils.npz"), **u)

    # ---- 4. matlab.mat ----
    import scipy.io
    m = scipy.io.loadmat("/root/reference/matlab.mat")
    keep = ["H_EST_LT_LS", "H_EST_PS_Linear", "H_EST_PS_Cubic", "H_EST_PS_Sinc", "eq_symbols",
            "tx_preamble_fft", "rx_preamble_fft", "tx_symb", "rx_symb", "rx_preamble1", "rx_preamble2",
            "tx_packet", "rx_packet", "tx_lptot", "rx_lptot"]
    np.savez(os.path.join(HERE, "matlab_mat.npz"), **{k: np.asarray(m[k]) for k in keep})

    # ---- 5. MMSE known answers (mpmath, 40 digits) ----
    import mpmath as mp
    mp.mp.dps = 40

    def mpc(z):
        return mp.mpc(mp.mpf(float(z.real)), mp.mpf(float(z.imag)))

    def mmse_mp(R, tx, rx, s2):
        n = len(tx)
        A = mp.matrix(n, n)
        y = mp.matrix(n, 1)
        for i in range(n):
            for j in range(n):
                A[i, j] = R[i, j]
            x = mpc(tx[i])
            A[i, i] += mp.mpf(float(s2)) / (x.real ** 2 + x.imag ** 2)
            y[i] = mpc(rx[i]) / x
        z = mp.lu_solve(A, y)
        Hm = R * z
        return np.array([complex(Hm[i]) for i in range(n)])

    k = {}
    hls = out["lt_ls"]
    tx0, rx0 = txs[0], rxs[0]
    Rm = mp.matrix(53, 53)
    for i in range(53):
        for j in range(53):
            Rm[i, j] = mpc(hls[i]) * mp.conj(mpc(hls[j]))
    k["inputs_h_full"] = mmse_mp(Rm, tx0, rx0, inp["ow2"])
    v = [mpc(tx0[i]) * mpc(hls[i]) for i in range(53)]
    vy = sum(mp.conj(v[i]) * mpc(rx0[i]) for i in range(53))
    vv = sum(abs(v[i]) ** 2 for i in range(53))
    g = vy / (mp.mpf(float(inp["ow2"])) + vv)
    k["inputs_h_g"] = np.array(complex(g))
    k["inputs_h_rank1"] = np.array([complex(g * mpc(hls[i])) for i in range(53)])
    # general full-rank R, per-frame sigma
    fr8 = synth.make_frames(8, seed=7, sigma2="perframe")
    Rg = synth.channel_covariance()
    Rgm = mp.matrix(53, 53)
    for i in range(53):
        for j in range(53):
            Rgm[i, j] = mpc(Rg[i, j])
    k["gen_R"] = Rg
    k["gen_tx"], k["gen_rx"], k["gen_sigma2"] = fr8["tx_symb"][:, 0, :], fr8["rx_symb"][:, 0, :], fr8["sigma2"]
    k["gen_H"] = np.stack([mmse_mp(Rgm, k["gen_tx"][f], k["gen_rx"][f], k["gen_sigma2"][f]) for f in range(8)])
    np.savez(os.path.join(HERE, "mmse_kat.npz"), **k)
    np.savez(os.path.join(HERE, "mmse_ref_composed.npz"), **mmse_composed(ref, inp, out["lt_ls"]))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
