"""e2e (host buffers) throughput of the shared-filter MMSE vs the host pipeline chunk size (GPU box)."""
import importlib, os, sys, time, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
sys.argv = ["x"]
import bench
wifi = importlib.import_module("80211parallelestimation_b200")
ctx = wifi.WifiContext(0)
n = 1 << 20
R = ctx.synth_covariance()
d = torch.full((53,), 9.6172e-08 / 8.875 ** 2, dtype=torch.float64, device="cuda"); d[26] = 9.6172e-08 / 1e-8
ctx.mmse_filter_form(R, d, want_W=False)
fr = ctx.synth_frames(n, "f32", want=("tx_symb", "rx_symb"))
tx = fr["tx_symb"][:, 0, :].contiguous(); rx = fr["rx_symb"][:, 0, :].contiguous()
htx = bench.pinned(wifi, (n, 53), np.complex64); hrx = bench.pinned(wifi, (n, 53), np.complex64); hH = bench.pinned(wifi, (n, 53), np.complex64)
htx[:] = tx.cpu().numpy(); hrx[:] = rx.cpu().numpy()
for _ in range(2): ctx.mmse_shared(htx, hrx, out=hH)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8): ctx.mmse_shared(htx, hrx, out=hH)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 8
print("chunk %s MB: %.2f ms per pass = %.3e frames/s" % (os.environ.get("WIFI_B200_HOST_CHUNK_MB", "16"), dt * 1e3, n / dt), flush=True)
