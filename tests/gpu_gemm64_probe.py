"""Timing + accuracy probe of the FP64 shared-filter kernels (GPU box; not a pytest file)."""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def one():
    import numpy as np, torch, synth
    from oracle.pyoracle import Oracle
    wifi = importlib.import_module("80211parallelestimation_b200")
    ctx = wifi.WifiContext(0); o = Oracle()
    R = synth.channel_covariance()
    fr = synth.make_frames(777, seed=3)
    tx, rx = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy()
    d = synth.OW2 / np.abs(tx[0]) ** 2
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    W = ctx.mmse_filter_form(dev(R), dev(d)).cpu().numpy()
    ref = o.mmse_apply(W, rx / tx)
    got = ctx.mmse_shared(dev(tx), dev(rx)).cpu().numpy()
    err = synth.rel_err(got, ref)
    n = 1 << 20
    frd = ctx.synth_frames(n, "f64", want=("tx_symb", "rx_symb"))
    tx0 = frd["tx_symb"][:, 0, :].contiguous(); rx0 = frd["rx_symb"][:, 0, :].contiguous()
    H = torch.empty_like(tx0)
    for _ in range(3): ctx.mmse_shared(tx0, rx0, out=H)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ctx.mmse_shared(tx0, rx0, out=H)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("gemm=%s f64 shared: rel_err %.2e | %.3f ms for %d frames = %.3e frames/s = %.2f TFLOP/s (%.1f%% of 37.0) = %.0f GB/s" % (
        os.environ.get("WIFI_B200_GEMM", "dmma"), err, ms, n, n / ms * 1e3, n * 22472 / ms / 1e9, 100 * n * 22472 / ms / 1e9 / 37.0, n * 2544 / ms / 1e6), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for g in ("dmma", "simt"):
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=dict(os.environ, WIFI_B200_GEMM=g), timeout=300)
