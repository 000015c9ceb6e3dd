import importlib, sys
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 18
for prec, cdt, cb in (("f32", torch.complex64, 8), ("f64", torch.complex128, 16)):
    pk = torch.randn(n, 1200, dtype=cdt, device="cuda"); lp = torch.randn(n, 160, dtype=cdt, device="cuda")
    out = ctx.frontend(pk, lp)
    for _ in range(3): ctx.frontend(pk, lp, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.frontend(pk, lp, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("frontend %s: %.4f ms = %.0f GB/s (%.1f %% of 6554)" % (prec, ms, n * ((1088 + 848) * cb + cb // 2) / ms / 1e6, 100 * n * ((1088 + 848) * cb + cb // 2) / ms / 1e6 / 6554.2))
    del pk, lp, out
