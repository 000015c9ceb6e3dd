"""Timing probe of the equalizer (GPU box; not a pytest file)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 19
for prec, cb in (("f32", 8), ("f64", 16)):
    fr = ctx.synth_frames(n, prec, want=("tx_pre", "rx_pre", "rx_symb"))
    H = ctx.lt_ls(fr["tx_pre"], fr["rx_pre"]); H2 = H.clone()
    eq = torch.empty_like(fr["rx_symb"])
    for _ in range(3): ctx.equalize(fr["rx_symb"], H, H2, out=eq)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.equalize(fr["rx_symb"], H, H2, out=eq)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("equalize %s: %.3f ms for %d frames = %.0f GB/s (%.1f%% of 6554)" % (prec, ms, n, n * 1696 * cb / ms / 1e6, 100 * n * 1696 * cb / ms / 1e6 / 6554.2), flush=True)
    del fr, eq
