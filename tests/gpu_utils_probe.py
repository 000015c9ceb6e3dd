"""Timing + accuracy probe of the batched utils (GPU box; not a pytest file)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 1:                      # A/B builds of the library
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[1]); print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0)
nb = 8192
g = torch.Generator(device="cuda").manual_seed(7)
for cdt in (torch.complex64, torch.complex128):
    for n in (53, 64, 17):
        A = torch.randn(nb, n, n, dtype=cdt, device="cuda", generator=g); B = torch.randn(nb, n, n, dtype=cdt, device="cuda", generator=g)
        ref = (A[:64].to(torch.complex128) @ B[:64].to(torch.complex128))
        got = ctx.multiply(A, B)
        err = float((got[:64].to(torch.complex128) - ref).abs().max() / ref.abs().max())
        for _ in range(3): ctx.multiply(A, B)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ctx.multiply(A, B)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("multiply %s n=%d: err %.2e | %.3f ms = %.3e matrices/s = %.2f TFLOP/s" % (str(cdt)[6:], n, err, ms, nb / ms * 1e3, nb * 8 * n ** 3 / ms / 1e9), flush=True)
    A = torch.randn(nb, 53, 53, dtype=cdt, device="cuda", generator=g)
    A = A @ A.conj().transpose(1, 2) / 53 + torch.eye(53, dtype=cdt, device="cuda")
    Y = ctx.inverse(A)
    err = float((Y[:64] @ A[:64] - torch.eye(53, dtype=cdt, device="cuda")).abs().max())
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.inverse(A)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("inverse %s n=53: |YA-I| %.2e | %.3f ms = %.3e matrices/s" % (str(cdt)[6:], err, ms, nb / ms * 1e3), flush=True)
    for n in (33, 47, 53, 64):                                   # general (non-Hermitian) matrices: the pivot order is not the identity
        G = torch.randn(256, n, n, dtype=cdt, device="cuda", generator=g)
        Yg = ctx.inverse(G)
        ref = torch.linalg.inv(G.to(torch.complex128))
        print("inverse %s n=%d general: |YG-I| %.2e, rel. to torch.linalg.inv(f64) %.2e" % (str(cdt)[6:], n,
              float((Yg @ G - torch.eye(n, dtype=cdt, device="cuda")).abs().max()), float((Yg.to(torch.complex128) - ref).abs().max() / ref.abs().max())), flush=True)
