import importlib, os, sys
ROOT = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 16
R = ctx.synth_covariance()
frd = ctx.synth_frames(n, "f64", per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
tx0 = frd["tx_symb"][:, 0, :].contiguous(); rx0 = frd["rx_symb"][:, 0, :].contiguous(); sd = frd["sigma2"]
H = torch.empty_like(tx0)
for _ in range(2): ctx.mmse_perframe(R, tx0, rx0, sd, flags=wifi.SOLVE_HPD, out=H)
torch.cuda.synchronize()
