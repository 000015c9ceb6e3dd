"""Timing of the fused receiver chain against the unfused path (GPU box; not a pytest file)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
if len(sys.argv) > 1:                      # A/B builds of the library
    wifi._lib.LIB_PATH = os.path.abspath(sys.argv[1]); print("library:", wifi._lib.LIB_PATH)
ctx = wifi.WifiContext(0)
n = int(os.environ.get("N", 1 << 18))


def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for prec, cdt, cb in (("f32", torch.complex64, 8), ("f64", torch.complex128, 16)):
    g = torch.Generator(device="cuda").manual_seed(3)
    mk = lambda w: torch.randn(n, w, dtype=cdt, device="cuda", generator=g)
    tp, tl, rp, rl = mk(1200), mk(160), mk(1200), mk(160)
    for want, c_per_frame in ((("lt_ls", "linear", "cubic", "sinc", "mmse_cconv", "eq", "ow2"), 1280 + 5 * 53 + 795),
                              (("lt_ls", "linear", "cubic", "sinc", "mmse_cconv", "ow2"), 64 * 2 + 128 * 2 + 5 * 53)):
        out = ctx.rx_chain(tp, tl, rp, rl, want=want)
        ms = timed(lambda: ctx.rx_chain(tp, tl, rp, rl, want=want, out=out))
        print("%s chain %-38s %.4f ms  %.1f M frames/s  %.0f GB/s on %d c" % (prec, "+".join(w[:4] for w in want), ms, n / ms / 1e3,
                                                                             n * c_per_frame * cb / ms / 1e6, c_per_frame), flush=True)
    # unfused: front-end x 2 + LT_LS + PS x 3 + cconv (on a gathered block 0) + equalizer
    def unfused():
        ts, tpre, _ = ctx.frontend(tp, tl, want_ow2=False)
        rs, rpre, ow2 = ctx.frontend(rp, rl)
        lt = ctx.lt_ls(tpre, rpre)
        ps = ctx.ps(ts, rs)
        mm = ctx.mmse_cconv(ts[:, 0, :].contiguous(), rs[:, 0, :].contiguous(), ow2, lt)
        return ctx.equalize(rs, lt, ps["linear"])
    ms = timed(unfused, 5)
    print("%s unfused (allocations included)              %.4f ms  %.1f M frames/s" % (prec, ms, n / ms / 1e3), flush=True)
    del tp, tl, rp, rl, out
    torch.cuda.empty_cache()
