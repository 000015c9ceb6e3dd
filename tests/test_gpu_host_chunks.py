"""The *_host entry points with MANY chunks in flight (wifi_set_host_chunk_bytes(1 MiB): two pipeline streams alternate, chunks
of ~2 400 frames): every result must equal the device-pointer call bit for bit -- this is what catches per-context scratch
shared by the two streams.  Runs in a subprocess so that a device fault cannot take the pytest process with it."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import importlib, os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, synth
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
ctx.set_host_chunk_bytes(1 << 20)
n = 20011
for prec, cdt in (("f32", np.complex64), ("f64", np.complex128)):
    fr = ctx.synth_frames(n, prec, per_frame_sigma=True)
    h = {k: v.cpu().numpy() for k, v in fr.items()}
    tx0, rx0 = np.ascontiguousarray(h["tx_symb"][:, 0, :]), np.ascontiguousarray(h["rx_symb"][:, 0, :])
    dtx0, drx0 = fr["tx_symb"][:, 0, :].contiguous(), fr["rx_symb"][:, 0, :].contiguous()
    eq = lambda a, b, what: (_ for _ in ()).throw(AssertionError(what + " " + prec)) if not np.array_equal(a, b.cpu().numpy()) else None
    lt = ctx.lt_ls(fr["tx_pre"], fr["rx_pre"]); eq(ctx.lt_ls(h["tx_pre"], h["rx_pre"]), lt, "lt_ls")
    ps = ctx.ps(fr["tx_symb"], fr["rx_symb"]); hps = ctx.ps(h["tx_symb"], h["rx_symb"])
    for k in ps: eq(hps[k], ps[k], "ps " + k)
    psm = ctx.ps(fr["tx_symb"], fr["rx_symb"], ("cubic",), matlab=True); eq(ctx.ps(h["tx_symb"], h["rx_symb"], ("cubic",), matlab=True)["cubic"], psm["cubic"], "ps matlab")
    R = ctx.synth_covariance()
    d = torch.full((53,), synth.OW2 / synth.AMP ** 2, dtype=torch.float64, device="cuda"); d[26] = synth.OW2 / 1e-8
    ctx.mmse_filter_form(R, d, want_W=False)
    eq(ctx.mmse_shared(tx0, rx0), ctx.mmse_shared(dtx0, drx0), "mmse_shared")
    eq(ctx.mmse_shared(h["tx_symb"].reshape(-1), h["rx_symb"].reshape(-1), frame_stride=795, n_frames=n), ctx.mmse_shared(dtx0, drx0), "mmse_shared strided")
    ctx.mmse_eig_prepare(R, (dtx0[0].abs().to(torch.float64)) ** 2)
    eq(ctx.mmse_perframe_eig(tx0, rx0, h["sigma2"]), ctx.mmse_perframe_eig(dtx0, drx0, fr["sigma2"]), "mmse_perframe_eig")
    Rp = R if prec == "f64" else R.to(torch.complex64)
    m = 3000
    eq(ctx.mmse_perframe(Rp.cpu().numpy(), tx0[:m], rx0[:m], h["sigma2"][:m], flags=wifi.SOLVE_HPD),
       ctx.mmse_perframe(Rp, dtx0[:m], drx0[:m], fr["sigma2"][:m], flags=wifi.SOLVE_HPD), "mmse_perframe")
    eq(ctx.mmse_cconv(tx0, rx0, h["sigma2"], lt.cpu().numpy()), ctx.mmse_cconv(dtx0, drx0, fr["sigma2"], lt), "mmse_cconv")
    eq(ctx.mmse_matlab(h["tx_symb"], h["rx_symb"], h["sigma2"], lt.cpu().numpy()), ctx.mmse_matlab(fr["tx_symb"], fr["rx_symb"], fr["sigma2"], lt), "mmse_matlab")
    eq(ctx.equalize(h["rx_symb"], lt.cpu().numpy(), ps["linear"].cpu().numpy()), ctx.equalize(fr["rx_symb"], lt, ps["linear"]), "equalize")
    rng = np.random.default_rng(1)
    pk = (rng.standard_normal((4001, 1200)) + 1j * rng.standard_normal((4001, 1200))).astype(cdt)
    lp = (rng.standard_normal((4001, 160)) + 1j * rng.standard_normal((4001, 160))).astype(cdt)
    a = ctx.frontend(pk, lp); b = ctx.frontend(torch.from_numpy(pk).cuda(), torch.from_numpy(lp).cuda())
    for x, y in zip(a, b): eq(x, y, "frontend")
print("host chunks ok")
'''


def test_host_entry_points_with_many_chunks():
    out = subprocess.run([sys.executable, "-c", SCRIPT % (ROOT, ROOT)], capture_output=True, text=True, timeout=600)
    print(out.stdout[-2000:], out.stderr[-3000:])
    assert out.returncode == 0 and "host chunks ok" in out.stdout
