"""One launch of the batched inverse per precision (GPU box; the command ncu captures for profiles/*_ncu_inverse.txt)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
nb = 8192
g = torch.Generator(device="cuda").manual_seed(7)
for cdt in (torch.complex64, torch.complex128):
    A = torch.randn(nb, 53, 53, dtype=cdt, device="cuda", generator=g)
    Y = ctx.inverse(A)
    torch.cuda.synchronize()
    print(str(cdt), float((Y[:16] @ A[:16] - torch.eye(53, dtype=cdt, device="cuda")).abs().max()))
