"""Prints measured error levels of the FP32 paths against the oracle (run on the GPU box; not a pytest file)."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, synth
from synth import rel_err
from oracle.pyoracle import Oracle

def main():
    wifi = importlib.import_module("80211parallelestimation_b200")
    o = Oracle()
    n = 40000
    fr = synth.make_frames(n, seed=11, sigma2="perframe")
    tx = fr["tx_symb"][:, 0, :].astype(np.complex64); rx = fr["rx_symb"][:, 0, :].astype(np.complex64)
    R = synth.channel_covariance()
    d = synth.OW2 / np.abs(fr["tx_symb"][0, 0, :]) ** 2
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for mode in ("tc", "simt"):
        os.environ["WIFI_B200_GEMM"] = mode
        ctx = wifi.WifiContext(0)
        W = ctx.mmse_filter_form(dev(R), dev(d)).cpu().numpy()
        ref = o.mmse_apply(W, rx.astype(complex) / tx.astype(complex))
        got = ctx.mmse_shared(dev(tx), dev(rx)).cpu().numpy()
        print("shared f32 %-4s: floor1e-3 %.2e  floor1e-2 %.2e  rel-to-frame-max %.2e" % (
            mode, rel_err(got, ref, 1e-3), rel_err(got, ref, 1e-2), (np.abs(got - ref) / np.abs(ref).max(axis=1, keepdims=True)).max()))
        ctx.close()
    os.environ["WIFI_B200_GEMM"] = "tc"
    ctx = wifi.WifiContext(0)
    m = 4096
    R32 = R.astype(np.complex64)
    for lo, hi in ((-8, -7), (-7, -6), (-6, -5), (-5, -4)):
        s2 = (10.0 ** np.random.default_rng(1).uniform(lo, hi, m)).astype(np.float32)
        fr = synth.make_frames(m, seed=3, sigma2=1e-6)
        t = fr["tx_symb"][:, 0, :]; H = fr["H_true"]
        noise = (np.random.default_rng(2).standard_normal((m, 53)) + 1j * np.random.default_rng(3).standard_normal((m, 53))) * np.sqrt(s2 / 2)[:, None]
        tx32 = t.astype(np.complex64); rx32 = (H * t + noise).astype(np.complex64)
        ref = o.mmse_perframe(R32.astype(complex), tx32.astype(complex), rx32.astype(complex), s2.astype(np.float64))
        for name, fl in (("hpd", wifi.SOLVE_HPD), ("pivot", wifi.SOLVE_PIVOT)):
            got = ctx.mmse_perframe(dev(R32), dev(tx32), dev(rx32), dev(s2), flags=fl).cpu().numpy()
            print("perframe f32 %-5s sigma2 1e%d..1e%d: floor1e-3 %.2e floor1e-2 %.2e" % (name, lo, hi, rel_err(got, ref, 1e-3), rel_err(got, ref, 1e-2)))
        t64 = t; r64 = H * t + noise
        ref64 = o.mmse_perframe(R, t64, r64, s2.astype(np.float64))
        for name, fl in (("hpd", wifi.SOLVE_HPD), ("pivot", wifi.SOLVE_PIVOT)):
            got = ctx.mmse_perframe(dev(R), dev(t64), dev(r64), dev(s2.astype(np.float64)), flags=fl).cpu().numpy()
            print("perframe f64 %-5s sigma2 1e%d..1e%d: floor1e-3 %.2e" % (name, lo, hi, rel_err(got, ref64, 1e-3)))

if __name__ == "__main__":
    main()
