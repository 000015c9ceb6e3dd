"""Timing probe for the per-frame solvers (GPU box; not a pytest file)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 18
R = ctx.synth_covariance()
for prec, peak in (("f32", 74.0), ("f64", 37.0)):
    fr = ctx.synth_frames(n, prec, per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
    tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous(); s2 = fr["sigma2"]
    Rp = R if prec == "f64" else R.to(torch.complex64)
    H = torch.empty_like(tx0)
    for _ in range(3): ctx.mmse_perframe(Rp, tx0, rx0, s2, flags=wifi.SOLVE_HPD, out=H)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.mmse_perframe(Rp, tx0, rx0, s2, flags=wifi.SOLVE_HPD, out=H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("hpd %s: %.3f ms for %d frames = %.3e frames/s = %.2f TFLOP/s algorithmic (%.1f%% of %g)" % (
        prec, ms, n, n / ms * 1e3, n * 441949 / ms / 1e9, 100 * n * 441949 / ms / 1e9 / peak, peak))
