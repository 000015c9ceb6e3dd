"""Limits of the eigen-domain per-frame MMSE that isolate its two tensor-core products (debugging aid):
   sigma2 -> 0   (full-rank R): s_i -> 0, H -> y           : tests the y path and the epilogue
   sigma2 -> inf (no null bin): s_i -> 1, H -> y - M G^H G y = 0 : tests the second product's transposed read of the images
   python profiles/probes/eig_limits_probe.py [alternative libwifi_b200.so]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch, synth
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 1:
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[1]); print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0)
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
n = 300
fr = synth.make_frames(n, seed=5, dtype=np.complex64)
tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
tx[:, 26] = 8.875                       # no null bin
y = rx.astype(complex) / tx.astype(complex)
R = synth.random_hpd(np.random.default_rng(3))
ctx.mmse_eig_prepare(R, np.abs(tx[0].astype(complex)) ** 2)
for s2v in (1e-20, 1e6):
    got = ctx.mmse_perframe_eig(dev(tx), dev(rx), dev(np.full(n, s2v, np.float32))).cpu().numpy()
    ref = y if s2v < 1 else np.zeros_like(y)
    print("sigma2 = %g: max |H - expected| / max |y| = %.3e" % (s2v, np.abs(got - ref).max() / np.abs(y).max()), flush=True)
