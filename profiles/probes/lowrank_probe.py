"""Timing probe of the low-rank per-frame PS_MMSE kernel (GPU box; not a pytest file): rank 4 and rank 7, both precisions, 1 Mi frames."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 1:                      # A/B builds of the library
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[1]); print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0)
n = 1 << 20
for prec, cb in (("f32", 8), ("f64", 16)):
    fr = ctx.synth_frames(n, prec, per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
    tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous(); s2 = fr["sigma2"]
    del fr
    H = torch.empty_like(tx0)
    for taps in (4, 7):
        ctx.mmse_lowrank_prepare(synth.channel_covariance(taps))
        for _ in range(3): ctx.mmse_perframe_lowrank(tx0, rx0, s2, out=H)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): ctx.mmse_perframe_lowrank(tx0, rx0, s2, out=H)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        bpf = 159 * cb + cb // 2
        print("lowrank %s rank %d: %.4f ms for %d frames = %.3e frames/s = %.0f GB/s (%.1f%% of 6554)" % (
            prec, taps, ms, n, n / ms * 1e3, n * bpf / ms / 1e6, 100 * n * bpf / ms / 1e6 / 6554.2), flush=True)
    del tx0, rx0, H
