"""General (pivoted) per-frame PS_MMSE solve: accuracy against the Hermitian kernel and throughput (GPU box; not a pytest file)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 16
for prec in ("f32", "f64"):
    fr = ctx.synth_frames(n, prec, per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
    tx0 = fr["tx_symb"][:, 0, :].contiguous(); rx0 = fr["rx_symb"][:, 0, :].contiguous(); s2 = fr["sigma2"]
    R = ctx.synth_covariance()
    Rp = R if prec == "f64" else R.to(torch.complex64)
    Hh = ctx.mmse_perframe(Rp, tx0, rx0, s2, flags=wifi.SOLVE_HPD)
    Hp = ctx.mmse_perframe(Rp, tx0, rx0, s2, flags=wifi.SOLVE_PIVOT)
    sc = Hh.abs().amax(dim=1, keepdim=True)
    err = float(((Hh - Hp).abs() / torch.maximum(Hh.abs(), 1e-3 * sc)).max())
    for _ in range(2): ctx.mmse_perframe(Rp, tx0, rx0, s2, flags=wifi.SOLVE_PIVOT, out=Hp)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.mmse_perframe(Rp, tx0, rx0, s2, flags=wifi.SOLVE_PIVOT, out=Hp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print("pivot %s: max rel diff vs HPD %.2e | %.3f ms per %d = %.2f M frames/s = %.1f %% of the FP64 peak on 441 949 flop" %
          (prec, err, ms, n, n / ms / 1e3, 100 * n * 441949 / ms / 1e9 / 37.2), flush=True)
