import importlib, sys
sys.path[:0] = ["/root/repo", "/root/repo/tests"]
import numpy as np, torch, synth
from oracle.pyoracle import Oracle
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0); o = Oracle()
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
for seed, n in ((1, 256), (77, 96), (5, 1024)):
    fr = synth.make_frames(n, seed=seed, sigma2="perframe")
    tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
    R = synth.channel_covariance()
    tx32, rx32 = tx.astype(np.complex64), rx.astype(np.complex64)
    s2 = fr["sigma2"].astype(np.float32)
    ctx.mmse_eig_prepare(R, np.abs(tx[0]) ** 2)
    He = ctx.mmse_perframe_eig(d(tx32), d(rx32), d(s2)).cpu().numpy()
    ref = o.mmse_perframe(R, tx32.astype(complex), rx32.astype(complex), s2.astype(np.float64))
    sc = np.abs(ref).max(axis=1, keepdims=True)
    e3 = np.abs(He - ref) / np.maximum(np.abs(ref), 1e-3 * sc)
    e2 = np.abs(He - ref) / np.maximum(np.abs(ref), 1e-2 * sc)
    ep = np.abs(He - ref) / sc
    f, k = np.unravel_index(e3.argmax(), e3.shape)
    print("seed %d n %d: floor1e-3 %.2e (frame %d bin %d |ref|/peak %.1e sigma2 %.1e)  floor1e-2 %.2e  of-peak %.2e  frames>1e-4: %d" %
          (seed, n, e3.max(), f, k, abs(ref[f, k]) / sc[f, 0], s2[f], e2.max(), ep.max(), int((e3.max(axis=1) > 1e-4).sum())))
