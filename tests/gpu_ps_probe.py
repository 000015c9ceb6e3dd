import importlib, os, sys, time
sys.path.insert(0, '/root/repo')
import torch, numpy as np
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 20
for prec in ("f32", "f64"):
    fr = ctx.synth_frames(n, prec, want=("tx_symb", "rx_symb"))
    tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous()
    outs = {k: torch.empty_like(tx0) for k in ("linear", "cubic", "sinc")}
    for name, (a, b) in (("frames795", (fr["tx_symb"], fr["rx_symb"])), ("vec53", (tx0, rx0))):
        for which in (("linear", "cubic", "sinc"), ("linear",)):
            o = {k: outs[k] for k in which}
            for _ in range(3): ctx.ps(a, b, which, out=o)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ctx.ps(a, b, which, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            cb = 8 if prec == "f32" else 16
            alg = (8 + 53 * len(which)) * cb
            print(os.environ.get("WIFI_B200_L2FETCH", "default32"), prec, name, len(which), "ms %.4f  alg GB/s %.0f  frames/s %.3e" % (ms, n * alg / ms / 1e6, n / ms * 1e3))
