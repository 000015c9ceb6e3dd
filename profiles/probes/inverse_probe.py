"""Batched inverse(): residual and throughput per order and precision.   python profiles/probes/inverse_probe.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 1:                      # A/B builds of the library
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[1]); print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0)
nb = 8192
g = torch.Generator(device="cuda").manual_seed(7)
for cdt, peak in ((torch.complex64, 74.4), (torch.complex128, 37.2)):
    for n in (53, 64, 40):
        A = torch.randn(nb, n, n, dtype=cdt, device="cuda", generator=g)
        Y = ctx.inverse(A)
        res = float((Y[:64] @ A[:64] - torch.eye(n, dtype=cdt, device="cuda")).abs().max())
        ref = torch.linalg.inv(A[:64].to(torch.complex128))
        err = float(((Y[:64].to(torch.complex128) - ref).abs().amax(dim=(1, 2)) / ref.abs().amax(dim=(1, 2))).max())
        for _ in range(3): ctx.inverse(A)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ctx.inverse(A)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("%s n=%d: |YA - I| %.2e, rel err vs torch f64 %.2e | %.3f ms per %d = %.2f M matrices/s = %.1f %% of the %s peak on 8 n^3" %
              (str(cdt)[6:], n, res, err, ms, nb, nb / ms / 1e3, 100 * nb * 8 * n ** 3 / ms / 1e9 / peak, "FP32" if peak > 50 else "FP64"), flush=True)
