# large frame-file run of the streaming pipeline: generate N frames on the GPU, write a FREQ file, run host/wifi_host_main --file, spot-check
import importlib, os, subprocess, sys, time, shutil
ROOT = "/root/repo"
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
wifi = importlib.import_module("80211parallelestimation_b200")
n = int(sys.argv[1]); d = sys.argv[2]
free = shutil.disk_usage(d).free
need = n * (13568 + 8484) * 1.05
print("free %.1f GB, need %.1f GB" % (free / 1e9, need / 1e9))
ctx = wifi.WifiContext(0)
fin, fout = os.path.join(d, "frames.bin"), os.path.join(d, "est.bin")
t0 = time.time()
with open(fin, "wb") as f:
    f.write(b"WIFIFRM1" + np.array([0, 0], np.uint32).tobytes() + np.array([n], np.uint64).tobytes() + bytes(40))
    CH = 1 << 16
    for key in ("tx_pre", "rx_pre", "tx_symb", "rx_symb"):
        for f0 in range(0, n, CH):
            fr = ctx.synth_frames(min(CH, n - f0), "f32", first_frame=f0, want=(key,))
            f.write(fr[key].cpu().numpy().tobytes())
print("wrote %s in %.1f s" % (fin, time.time() - t0))
subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
for rep in range(2):
    out = subprocess.run([os.path.join(ROOT, "host", "wifi_host_main"), "--file", fin, fout], capture_output=True, text=True)
    print(out.stdout, out.stderr)
# spot check the last chunk's LT_LS plane
fr = ctx.synth_frames(1000, "f32", first_frame=n - 1000, want=("tx_pre", "rx_pre"))
ref = ctx.lt_ls(fr["tx_pre"], fr["rx_pre"]).cpu().numpy()
got = np.fromfile(fout, np.complex64, 1000 * 53, offset=64 + (n - 1000) * 53 * 8).reshape(1000, 53)
print("LT_LS plane tail equal:", np.array_equal(got, ref))
os.remove(fin); os.remove(fout)
