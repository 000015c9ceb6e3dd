"""Shared-filter MMSE reading block 0 of whole frames in place (frame_stride 795) vs dense block vectors (GPU box)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 19
R = ctx.synth_covariance()
d = torch.full((53,), 9.6172e-08 / 8.875 ** 2, dtype=torch.float64, device="cuda"); d[26] = 9.6172e-08 / 1e-8
ctx.mmse_filter_form(R, d, want_W=False)
for prec, cb in (("f32", 8), ("f64", 16)):
    fr = ctx.synth_frames(n, prec, want=("tx_symb", "rx_symb"))
    tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous()
    H = torch.empty_like(tx0); H2 = torch.empty_like(tx0)
    for name, fn in (("dense [n][53]", lambda: ctx.mmse_shared(tx0, rx0, out=H)),
                     ("in place, stride 795", lambda: ctx.mmse_shared(fr["tx_symb"].reshape(-1), fr["rx_symb"].reshape(-1), frame_stride=795, n_frames=n, out=H2))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s %-22s %.3f ms for %d frames = %.3e frames/s (%.0f GB/s algorithmic)" % (prec, name, ms, n, n / ms * 1e3, n * 159 * cb / ms / 1e6), flush=True)
    print("   max |dense - strided| =", float((H - H2).abs().max()))
    del fr
