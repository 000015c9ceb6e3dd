"""Timing + accuracy probe for the per-frame solver variants (GPU box; not a pytest file).
    python tests/gpu_hpd_probe.py            # loops over WIFI_HPD_CFG variants in subprocesses
"""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))

def one():
    import numpy as np, torch, synth
    from oracle.pyoracle import Oracle
    wifi = importlib.import_module("80211parallelestimation_b200")
    ctx = wifi.WifiContext(0)
    o = Oracle()
    cfg = os.environ.get("WIFI_HPD_CFG", "-1")
    n = 1 << 18
    R = ctx.synth_covariance()
    # accuracy on host-generated frames
    fr = synth.make_frames(96, seed=77, sigma2="perframe")
    tx, rx, s2 = fr["tx_symb"][:, 0, :].copy(), fr["rx_symb"][:, 0, :].copy(), fr["sigma2"]
    Rn = synth.channel_covariance()
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for prec, flags, peak in (("f32", wifi.SOLVE_HPD, 74.0), ("f32", wifi.SOLVE_HPD | wifi.SOLVE_WIDE, 37.0), ("f64", wifi.SOLVE_HPD, 37.0)):
        if prec == "f32":
            t32, r32, s32, R32 = tx.astype(np.complex64), rx.astype(np.complex64), s2.astype(np.float32), Rn.astype(np.complex64)
            ref = o.mmse_perframe(R32.astype(complex), t32.astype(complex), r32.astype(complex), s32.astype(np.float64))
            got = ctx.mmse_perframe(dev(R32), dev(t32), dev(r32), dev(s32), flags=flags).cpu().numpy()
        else:
            ref = o.mmse_perframe(Rn, tx, rx, s2)
            got = ctx.mmse_perframe(dev(Rn), dev(tx), dev(rx), dev(s2), flags=flags).cpu().numpy()
        err = synth.rel_err(got, ref)
        frd = ctx.synth_frames(n, prec, per_frame_sigma=True, want=("tx_symb", "rx_symb", "sigma2"))
        tx0 = frd["tx_symb"][:, 0, :].contiguous(); rx0 = frd["rx_symb"][:, 0, :].contiguous(); sd = frd["sigma2"]
        Rp = R if prec == "f64" else R.to(torch.complex64)
        H = torch.empty_like(tx0)
        for _ in range(3): ctx.mmse_perframe(Rp, tx0, rx0, sd, flags=flags, out=H)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ctx.mmse_perframe(Rp, tx0, rx0, sd, flags=flags, out=H)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("cfg %s hpd %s flags %d: rel_err %.2e | %.3f ms for %d frames = %.3e frames/s = %.2f TFLOP/s algorithmic (%.1f%% of %g)" % (
            cfg, prec, flags, err, ms, n, n / ms * 1e3, n * 441949 / ms / 1e9, 100 * n * 441949 / ms / 1e9 / peak, peak), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for cfg in (sys.argv[1:] or ["-1", "1", "2", "3", "4"]):
            env = dict(os.environ, WIFI_HPD_CFG=cfg)
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=env, timeout=300)
