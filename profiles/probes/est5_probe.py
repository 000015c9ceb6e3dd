"""wifi_estimate_all_batch (BASELINE configs[4]) against its stand-alone kernels, one by one (used for the A/B runs of the rejected fusions).
   python profiles/probes/est5_probe.py [n_frames]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 2:                      # A/B builds of the library (profiles/tmp_variants/, not part of the product)
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[2])
    print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20


def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for prec, cb in (("f32", 8), ("f64", 16)):
    fr = ctx.synth_frames(n, prec, want=("tx_pre", "rx_pre", "tx_symb", "rx_symb"))
    R = ctx.synth_covariance()
    d = torch.full((53,), 9.6172e-08 / 8.875 ** 2, dtype=torch.float64, device="cuda"); d[26] = 9.6172e-08 / 1e-8
    ctx.mmse_filter_form(R, d, want_W=False)
    mk = lambda: torch.empty_like(fr["tx_pre"])
    o = {k: mk() for k in ("lt_ls", "linear", "cubic", "sinc", "mmse")}
    eq = torch.empty_like(fr["rx_symb"])
    tx0, rx0 = fr["tx_symb"][:, 0, :].contiguous(), fr["rx_symb"][:, 0, :].contiguous()
    ms = t(lambda: ctx.estimate_all(fr["tx_pre"], fr["rx_pre"], fr["tx_symb"], fr["rx_symb"], equalize=False, out=o))
    print("%s est5 whole frames (stride 795): %.4f ms  %.0f GB/s on 477 c" % (prec, ms, n * 477 * cb / ms / 1e6))
    ms = t(lambda: ctx.estimate_all(fr["tx_pre"], fr["rx_pre"], tx0, rx0, out=o))
    print("%s est5 block vectors (stride 53): %.4f ms  %.0f GB/s on 477 c" % (prec, ms, n * 477 * cb / ms / 1e6))
    o2 = dict(o, eq=eq)
    ms = t(lambda: ctx.estimate_all(fr["tx_pre"], fr["rx_pre"], fr["tx_symb"], fr["rx_symb"], out=o2))
    print("%s est5 + equalizer:               %.4f ms  %.0f GB/s on 2014 c" % (prec, ms, n * 2014 * cb / ms / 1e6))
    a = t(lambda: ctx.lt_ls(fr["tx_pre"], fr["rx_pre"], out=o["lt_ls"]))
    b = t(lambda: ctx.ps(fr["tx_symb"], fr["rx_symb"], out={k: o[k] for k in ("linear", "cubic", "sinc")}))
    c = t(lambda: ctx.mmse_shared(fr["tx_symb"].reshape(-1), fr["rx_symb"].reshape(-1), frame_stride=795, n_frames=n, out=o["mmse"]))
    c2 = t(lambda: ctx.mmse_shared(tx0, rx0, out=o["mmse"]))
    e = t(lambda: ctx.equalize(fr["rx_symb"], o["lt_ls"], o["linear"], out=eq))
    print("%s stand-alone: lt_ls %.4f + ps %.4f + mmse(795) %.4f [mmse(53) %.4f] + eq %.4f = %.4f ms" % (prec, a, b, c, c2, e, a + b + c + e), flush=True)
    del fr, o, o2, eq, tx0, rx0
