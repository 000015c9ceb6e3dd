"""Timing + accuracy probe of the FP32 shared-filter kernel (GPU box; not a pytest file)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from oracle.pyoracle import Oracle
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 1:                      # A/B builds of the library
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[1]); print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0); o = Oracle()
R = synth.channel_covariance()
fr = synth.make_frames(18949, seed=3, dtype=np.complex64)
tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
d = synth.OW2 / np.abs(tx[0].astype(complex)) ** 2
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
W = ctx.mmse_filter_form(dev(R), dev(d)).cpu().numpy()
ref = o.mmse_apply(W, rx.astype(complex) / tx.astype(complex))
got = ctx.mmse_shared(dev(tx), dev(rx)).cpu().numpy()
err = synth.rel_err(got, ref, 1e-2); err3 = synth.rel_err(got, ref, 1e-3); errp = float((np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)).max())
for n in (1 << 20, 1 << 22):
    frd = ctx.synth_frames(n, "f32", want=("tx_symb", "rx_symb")) if n <= (1 << 20) else None
    if frd is not None:
        tx0 = frd["tx_symb"][:, 0, :].contiguous(); rx0 = frd["rx_symb"][:, 0, :].contiguous()
        del frd
    else:
        tx0 = tx0.repeat(4, 1); rx0 = rx0.repeat(4, 1)
    H = torch.empty_like(tx0)
    for _ in range(3): ctx.mmse_shared(tx0, rx0, out=H)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ctx.mmse_shared(tx0, rx0, out=H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("f32 shared tc: rel_err %.2e (floor 1e-3: %.2e, of peak %.2e) | %.4f ms for %d frames = %.3e frames/s = %.0f GB/s (%.1f%% of 6554)" % (
        err, err3, errp, ms, n, n / ms * 1e3, n * 1272 / ms / 1e6, 100 * n * 1272 / ms / 1e6 / 6554.2), flush=True)
