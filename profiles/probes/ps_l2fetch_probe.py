"""PS estimators with the device-wide L2 fetch granularity at its default (128 B) and at 32 / 64 B.  The library no longer
touches this device-global limit (round 1 set it to 32 in wifi_create); this probe measures what that is worth for the pilot
gather, whose loads already carry the per-instruction L2::64B qualifier.   python profiles/probes/ps_l2fetch_probe.py"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch
wifi = importlib.import_module("80211parallelestimation_b200")
rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")
cudaLimitMaxL2FetchGranularity = 0x05
ctx = wifi.WifiContext(0)
n = 1 << 20
for gran in (128, 32, 64, 128):
    assert rt.cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, ctypes.c_size_t(gran)) == 0
    for prec in ("f32", "f64"):
        fr = ctx.synth_frames(n, prec, want=("tx_symb", "rx_symb"))
        a, b = fr["tx_symb"], fr["rx_symb"]
        for which in (("linear",), ("linear", "cubic", "sinc")):
            o = {k: torch.empty(n, 53, dtype=a.dtype, device="cuda") for k in which}
            for _ in range(3): ctx.ps(a, b, which, out=o)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): ctx.ps(a, b, which, out=o)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            cb = 8 if prec == "f32" else 16
            alg = (8 + 53 * len(which)) * cb
            print("L2 fetch granularity %3d  %s  %d estimator(s): %.4f ms  algorithmic %.0f GB/s  %.3e frames/s" % (gran, prec, len(which), ms, n * alg / ms / 1e6, n / ms * 1e3), flush=True)
        del fr, a, b, o
