"""Element indices beyond 2^31: 41 Mi frames of [n][53] block vectors (2.3e9 complex values per array) through LT_LS and the
tcgen05 shared-filter GEMM, 2.8 Mi whole frames (2.3e9 values) through the equalizer.  Device-generated frames; the tail of
each batch -- where a 32-bit index would have wrapped -- is compared with the oracle.  Needs ~70 GB of HBM."""
import importlib

import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu
NSC, NBLK = 53, 15


def test_indices_beyond_2_31(oracle):
    import torch
    wifi = importlib.import_module("80211parallelestimation_b200")
    free, total = torch.cuda.mem_get_info()
    if free < 80e9:
        pytest.skip("needs 80 GB of free HBM")
    ctx = wifi.WifiContext(0)
    n = 41 * (1 << 20) + 3                       # 43 M frames: 2.28e9 complex values per [n][53] array, odd count
    assert n * NSC > 2 ** 31
    chunk = 1 << 20
    tx = torch.empty((n, NSC), dtype=torch.complex64, device="cuda"); rx = torch.empty_like(tx)
    for f0 in range(0, n, chunk):                # block-0 vectors of the global synthetic sequence
        m = min(chunk, n - f0)
        fr = ctx.synth_frames(m, "f32", first_frame=f0, want=("tx_symb", "rx_symb"))
        tx[f0:f0 + m] = fr["tx_symb"][:, 0, :]; rx[f0:f0 + m] = fr["rx_symb"][:, 0, :]
        del fr
    tail = slice(n - 300, n)
    t64 = tx[tail].cpu().numpy().astype(np.complex128); r64 = rx[tail].cpu().numpy().astype(np.complex128)
    # LT_LS (flat float4 kernel + odd tail element)
    H = ctx.lt_ls(tx, rx)
    assert rel_err(H[tail].cpu().numpy(), oracle.lt_ls(t64, r64)) < 1e-4
    assert rel_err(H[:300].cpu().numpy(), oracle.lt_ls(tx[:300].cpu().numpy().astype(complex), rx[:300].cpu().numpy().astype(complex))) < 1e-4
    # shared-filter GEMM
    R = ctx.synth_covariance()
    d = torch.full((NSC,), synth.OW2 / synth.AMP ** 2, dtype=torch.float64, device="cuda"); d[26] = synth.OW2 / 1e-8
    W = ctx.mmse_filter_form(R, d).cpu().numpy()
    ctx.mmse_shared(tx, rx, out=H)
    assert rel_err(H[tail].cpu().numpy(), oracle.mmse_apply(W, r64 / t64), 1e-2) < 1e-4
    mid = slice(40 * (1 << 20) + 517, 40 * (1 << 20) + 517 + 200)        # element index just past 2^31
    assert mid.start * NSC > 2 ** 31 - 2 ** 26
    ref = oracle.mmse_apply(W, rx[mid].cpu().numpy().astype(complex) / tx[mid].cpu().numpy().astype(complex))
    assert rel_err(H[mid].cpu().numpy(), ref, 1e-2) < 1e-4
    del tx, rx, H
    torch.cuda.empty_cache()
    # equalizer on whole frames: 2.8 Mi frames x 795 = 2.3e9 values
    nf = 2_800_001
    assert nf * NBLK * NSC > 2 ** 31
    rxs = torch.empty((nf, NBLK, NSC), dtype=torch.complex64, device="cuda")
    hl = torch.empty((nf, NSC), dtype=torch.complex64, device="cuda"); hp = torch.empty_like(hl)
    for f0 in range(0, nf, chunk):
        m = min(chunk, nf - f0)
        fr = ctx.synth_frames(m, "f32", first_frame=f0, want=("rx_symb", "H_true"))
        rxs[f0:f0 + m] = fr["rx_symb"]; hl[f0:f0 + m] = fr["H_true"]; hp[f0:f0 + m] = fr["H_true"] * 1.01
        del fr
    eq = ctx.equalize(rxs, hl, hp)
    t = slice(nf - 40, nf)
    c = lambda x: x[t].cpu().numpy().astype(np.complex128)
    assert rel_err(eq[t].cpu().numpy(), oracle.equalize(c(rxs), c(hl), c(hp)), floor=1e-6) < 1e-4
    assert float(eq[:, :, 26].abs().max()) == 0.0
