"""Timing + accuracy probe of the eigen-domain per-frame MMSE (GPU box; not a pytest file)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from oracle.pyoracle import Oracle
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0); o = Oracle()
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
n = 1 << 20
for prec, cdt in (("f32", np.complex64), ("f64", np.complex128)):
    fr = synth.make_frames(96, seed=77, sigma2="perframe", dtype=cdt)
    tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
    s2 = fr["sigma2"].astype(np.float32 if prec == "f32" else np.float64)
    R = synth.channel_covariance().astype(cdt).astype(np.complex128)
    ctx.mmse_eig_prepare(R, np.abs(tx[0].astype(complex)) ** 2)
    got = ctx.mmse_perframe_eig(dev(tx), dev(rx), dev(s2)).cpu().numpy()
    ref = o.mmse_perframe(R, tx.astype(complex), rx.astype(complex), s2.astype(np.float64))
    err = synth.rel_err(got, ref)
    frd = ctx.synth_frames(n, prec, per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
    tx0 = frd["tx_symb"][:, 0, :].contiguous(); rx0 = frd["rx_symb"][:, 0, :].contiguous(); sd = frd["sigma2"]
    del frd
    H = torch.empty_like(tx0)
    for _ in range(3): ctx.mmse_perframe_eig(tx0, rx0, sd, out=H)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.mmse_perframe_eig(tx0, rx0, sd, out=H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("eig %s: rel_err %.2e | %.3f ms for %d frames = %.3e frames/s" % (prec, err, ms, n, n / ms * 1e3), flush=True)
