"""Host-side synthetic 802.11a frames of the inputs.h shape for the tests.

Test infrastructure: numpy only.  The product's own on-device generator
(wifi_synth_frames in the CUDA library) follows the same recipe but is not required to be
bit-identical; full-size GPU tests compare a sampled subset of device-generated frames
against the oracle instead.
"""
import numpy as np

NSC, NBLK, DC = 53, 15, 26
PILOTS = (5, 19, 33, 47)
OW2 = 9.6172e-08            # inputs.h:18
AMP = 8.875                 # |tx| of the inputs.h frame (inputs.h:20-74)
TAPS = 4
TAP_POWER = 8e-5

# 802.11a long-training sequence L_-26..26 (the sign pattern of inputs.h:20-74)
LTS = np.array([1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 1, 1, -1, -1, 1, 1, -1, 1, -1, 1, 1, 1, 1, 0,
                1, -1, -1, 1, 1, -1, 1, -1, 1, -1, -1, -1, -1, -1, 1, 1, -1, -1, 1, -1, 1, -1, 1, 1, 1, 1], np.float64)


def tap_powers(taps=TAPS):
    return TAP_POWER * 2.0 ** (-np.arange(taps))


def channel_covariance(taps=TAPS, ridge=0.0):
    """Theoretical E[H H^H] of the generator's channel: Hermitian PSD 53x53 (rank `taps`)."""
    k = np.arange(NSC) - DC
    l = np.arange(taps)
    E = np.exp(-2j * np.pi * np.outer(k, l) / 64.0)            # [53][taps]
    R = (E * tap_powers(taps)) @ E.conj().T
    R = 0.5 * (R + R.conj().T)
    return R + ridge * np.eye(NSC)


def random_hpd(rng, scale=1e-4, n=NSC):
    """A full-rank random Hermitian positive-definite matrix (general-R tests)."""
    a = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return scale * (a @ a.conj().T) / n


def make_frames(n, seed=0x80211, sigma2=None, dtype=np.complex128):
    """sigma2: None -> shared OW2; 'perframe' -> log-uniform [1e-8, 1e-5]; float -> shared."""
    rng = np.random.default_rng(seed)
    if sigma2 is None:
        s2 = np.full(n, OW2)
    elif isinstance(sigma2, str):
        s2 = 10.0 ** rng.uniform(-8, -5, n)
    else:
        s2 = np.full(n, float(sigma2))
    k = np.arange(NSC) - DC
    l = np.arange(TAPS)
    E = np.exp(-2j * np.pi * np.outer(l, k) / 64.0)            # [taps][53]
    a = (rng.standard_normal((n, TAPS)) + 1j * rng.standard_normal((n, TAPS))) * np.sqrt(tap_powers() / 2)
    H = a @ E                                                   # [n][53]

    def cn(shape):
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * np.sqrt(s2 / 2).reshape((n,) + (1,) * (len(shape) - 1))

    tx_pre = np.broadcast_to(AMP * LTS, (n, NSC)).astype(np.complex128).copy()
    tx_pre[:, DC] = -2e-4
    rx_pre = H * tx_pre + cn((n, NSC))
    bits = rng.integers(0, 2, (n, NBLK, NSC)) * 2 - 1
    tx_symb = (AMP * bits).astype(np.complex128)
    tx_symb[:, :, DC] = -1e-4
    rx_symb = H[:, None, :] * tx_symb + cn((n, NBLK, NSC))
    out = dict(tx_pre=tx_pre, rx_pre=rx_pre, tx_symb=tx_symb, rx_symb=rx_symb, H_true=H, sigma2=s2)
    if dtype != np.complex128:
        for key in ("tx_pre", "rx_pre", "tx_symb", "rx_symb", "H_true"):
            out[key] = out[key].astype(dtype)
    return out


def rel_err(got, ref, floor=1e-3):
    """Per-sub-carrier relative error with the survey's floor:
    |d| / max(|ref_k|, floor * max_k |ref|), max over everything.  NaNs must coincide."""
    got = np.asarray(got); ref = np.asarray(ref)
    nan_g, nan_r = np.isnan(got), np.isnan(ref)
    if not np.array_equal(nan_g, nan_r):
        return np.inf
    g = np.where(nan_r, 0, got); r = np.where(nan_r, 0, ref)
    scale = np.abs(r).max(axis=-1, keepdims=True) if r.ndim else np.abs(r)
    den = np.maximum(np.abs(r), floor * scale)
    den = np.where(den == 0, 1.0, den)
    return float((np.abs(g - r) / den).max()) if g.size else 0.0


def to_time_domain(fr, seed=1, lts_noise=2e-4):
    """Time samples whose front-end output (WiFi_blocks_extraction.m, WiFi_RX.m:19-31) is `fr` again:
    symb[i] = X[(i - 26) mod 64] is inverted, a 16-sample cyclic prefix is prepended to every OFDM block, and the long
    training field is [32 guard samples, p2, p1] with p1,2 = t -+ w (so (p1 + p2)/2 = t and ow2 = sum |2w|^2 / 128)."""
    rng = np.random.default_rng(seed)

    def td(spec53):                                     # [..., 53] -> [..., 64] time samples
        X = np.zeros(spec53.shape[:-1] + (64,), np.complex128)
        X[..., (np.arange(NSC) - 26) % 64] = spec53
        return np.fft.ifft(X, axis=-1)

    out = {}
    n = fr["tx_pre"].shape[0]
    for side in ("tx", "rx"):
        x = td(fr[side + "_symb"].astype(np.complex128))                    # [n][15][64]
        out[side + "_packet"] = np.concatenate([x[..., 48:], x], axis=-1).reshape(n, NBLK * 80)
        t = td(fr[side + "_pre"].astype(np.complex128))                     # [n][64]
        w = (rng.standard_normal((n, 64)) + 1j * rng.standard_normal((n, 64))) * (lts_noise if side == "rx" else 0.0)
        out[side + "_lptot"] = np.concatenate([t[:, 32:], t - w, t + w], axis=-1)      # p2 = lptot[32:96], p1 = lptot[96:160]
    return out
