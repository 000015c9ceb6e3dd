"""Device eigen-domain result vs a numpy evaluation of the same formulas on the worst frames (GPU box)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from oracle.pyoracle import Oracle
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0); o = Oracle()
frames = [247993, 105443, 89256, 129727, 5]
R = ctx.synth_covariance().cpu().numpy()
tx = []; rx = []; s2 = []
for f in frames:
    fr = ctx.synth_frames(1, "f64", first_frame=f, per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
    tx.append(fr["tx_symb"][0, 0].cpu().numpy()); rx.append(fr["rx_symb"][0, 0].cpu().numpy()); s2.append(float(fr["sigma2"][0]))
tx = np.array(tx); rx = np.array(rx); s2 = np.array(s2)
ref = o.mmse_perframe(R, tx, rx, s2)
dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
absx2 = np.abs(tx[0]) ** 2
ctx.mmse_eig_prepare(R, absx2)
He = ctx.mmse_perframe_eig(dev(tx), dev(rx), dev(s2)).cpu().numpy()
# numpy evaluation with LAPACK eigh
DC = 26; N = np.array([k for k in range(53) if k != DC])
ax = np.sqrt(absx2[N])
S = (R[np.ix_(N, N)] * ax[:, None]) * ax[None, :]
lam, V = np.linalg.eigh(S)
G = V.conj().T * ax[None, :]; G2 = V / ax[:, None]
b = R[N, DC]; p = G @ b
y = rx / tx
u = y[:, N] @ G.T
inv = 1 / (lam[None, :] + s2[:, None])
beta = (np.conj(p)[None, :] * u * inv).sum(1); gamma = ((np.abs(p) ** 2)[None, :] * inv).sum(1)
q = R[DC, DC].real - gamma; den = s2 / absx2[DC] + q
zd = (y[:, DC] - beta) / den
v = (s2[:, None] * inv) * (u - p[None, :] * zd[:, None])
Hn = np.empty_like(y); Hn[:, N] = y[:, N] - v @ G2.T; Hn[:, DC] = beta + (y[:, DC] - beta) * (q / den)
for i, f in enumerate(frames):
    print("frame %7d s2 %.2e: device eig err %.2e | numpy(eigh) err %.2e | device vs numpy %.2e | q %.3e den %.3e zd %.3e" % (
        f, s2[i], synth.rel_err(He[i], ref[i]), synth.rel_err(Hn[i], ref[i]), synth.rel_err(He[i], Hn[i]), q[i], den[i], abs(zd[i])))
print("lam top", lam[-5:], "garbage max", np.abs(lam[:-4]).max())
