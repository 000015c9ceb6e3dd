"""Which of the two per-frame algorithms is off on the worst frames of a 256 Ki batch (GPU box)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, synth
from oracle.pyoracle import Oracle
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0); o = Oracle()
n = 1 << 18
fr = ctx.synth_frames(n, "f64", per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous(); s2 = fr["sigma2"]
R = ctx.synth_covariance()
Hs = ctx.mmse_perframe(R, tx0, rx0, s2, flags=wifi.SOLVE_HPD)
ctx.mmse_eig_prepare(R, (tx0[0].abs()) ** 2)
He = ctx.mmse_perframe_eig(tx0, rx0, s2)
scale = Hs.abs().amax(dim=1, keepdim=True)
d = ((Hs - He).abs() / torch.maximum(Hs.abs(), 1e-3 * scale))
per = d.amax(dim=1)
top = torch.topk(per, 12).indices
c = lambda t: t[top].cpu().numpy()
ref = o.mmse_perframe(R.cpu().numpy(), c(tx0), c(rx0), c(s2))
for i, f in enumerate(top.tolist()):
    k = int(d[f].argmax())
    print("frame %7d sigma2 %.2e bin %2d diff %.2e | solve err %.2e  eig err %.2e | |y_dc| %.2e" % (
        f, float(s2[f]), k, float(per[f]), synth.rel_err(c(Hs)[i], ref[i]), synth.rel_err(c(He)[i], ref[i]), float((rx0[f, 26] / tx0[f, 26]).abs())))
