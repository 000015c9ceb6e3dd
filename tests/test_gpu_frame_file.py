"""Frame files (include/wifi_frame_file.h) through the C host driver: frequency-domain and time-domain inputs, every
output plane against the oracle.  Runs on the B200 box."""
import os
import subprocess

import numpy as np
import pytest

import synth
from synth import rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NSC, NBLK = 53, 15


def write_file(path, kind, dtype, planes):
    n = planes[0].shape[0]
    with open(path, "wb") as f:
        f.write(b"WIFIFRM1" + np.array([0 if dtype == np.complex64 else 1, kind], np.uint32).tobytes() + np.array([n], np.uint64).tobytes() + bytes(40))
        for p in planes:
            f.write(np.ascontiguousarray(p.astype(dtype)).tobytes())


def read_est(path, n, dtype):
    raw = open(path, "rb").read()
    assert raw[:8] == b"WIFIFRM1" and np.frombuffer(raw[12:16], np.uint32)[0] == 2
    out, o = [], 64
    for w in (NSC, NSC, NSC, NSC, NSC, NBLK * NSC):
        out.append(np.frombuffer(raw, dtype, n * w, o).reshape(n, w)); o += n * w * np.dtype(dtype).itemsize
    rdt = np.float32 if dtype == np.complex64 else np.float64
    out.append(np.frombuffer(raw, rdt, n, o))
    return out


@pytest.mark.parametrize("kind", ["freq", "time"])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_frame_file_pipeline(tmp_path, oracle, kind, dtype):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "host")])
    n = 777
    fr = synth.make_frames(n, seed=31)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    r = lambda x: x.astype(dtype).astype(np.complex128)
    if kind == "freq":
        write_file(fin, 0, dtype, [fr["tx_pre"], fr["rx_pre"], fr["tx_symb"].reshape(n, -1), fr["rx_symb"].reshape(n, -1)])
        tx_pre, rx_pre, tx_symb, rx_symb = r(fr["tx_pre"]), r(fr["rx_pre"]), r(fr["tx_symb"]), r(fr["rx_symb"])
        ow2 = np.full(n, np.float32(9.6172e-08) if dtype == np.complex64 else 9.6172e-08, np.float64)
    else:
        td = synth.to_time_domain(fr)
        write_file(fin, 1, dtype, [td["tx_packet"], td["rx_packet"], td["tx_lptot"], td["rx_lptot"]])
        tx_symb, tx_pre, _ = oracle.frontend(r(td["tx_packet"]), r(td["tx_lptot"]))
        rx_symb, rx_pre, ow2 = oracle.frontend(r(td["rx_packet"]), r(td["rx_lptot"]))
        assert rel_err(tx_symb, fr["tx_symb"]) < 1e-3 and rel_err(rx_pre, fr["rx_pre"]) < 1e-3      # the synthetic time frames invert the front-end (FP32-rounded samples: 1e-4 on the -1e-4 DC bin)
    out = subprocess.run([os.path.join(ROOT, "host", "wifi_host_main"), "--file", fin, fout], capture_output=True, text=True, timeout=300, cwd=ROOT)
    print(out.stdout, out.stderr)
    assert out.returncode == 0 and "frames/s" in out.stdout
    lt, lin, cub, sinc, mmse, eq, ow2_got = read_est(fout, n, dtype)
    tol = 1e-10 if dtype == np.complex128 else 1e-4
    if kind == "time":        # the estimators run on the device's own front-end output; compare stage by stage through the oracle
        assert np.allclose(ow2_got, ow2, rtol=1e-10 if dtype == np.complex128 else 1e-4)
    ref_lt = oracle.lt_ls(tx_pre, rx_pre)
    ftol = tol if kind == "freq" else max(tol, 1e-9) * 20          # time files: FFT rounding is amplified by the LS divide of small bins
    assert rel_err(lt, ref_lt) < ftol
    tx0, rx0 = tx_symb[:, 0, :], rx_symb[:, 0, :]
    for got, name in ((lin, "linear"), (cub, "cubic"), (sinc, "sinc")):
        assert rel_err(got, getattr(oracle, "ps_" + name)(tx0, rx0)) < ftol, name
    sub = slice(0, 48)                                                # the long-double per-frame oracle solve is slow: a sample
    ref_mmse = oracle.mmse_cconv_batch(tx0[sub], rx0[sub], ow2[sub], ref_lt[sub])
    # the device's PS_MMSE consumes the device's own LT_LS plane (H_ls argument of main.c:148); feed the oracle the same values
    ref_mmse_dev = oracle.mmse_cconv_batch(tx0[sub], rx0[sub], ow2[sub], lt[sub].astype(np.complex128))
    e_dev, e_ref = rel_err(mmse[sub], ref_mmse_dev), rel_err(mmse[sub], ref_mmse)
    print("PS_MMSE plane: rel_err %.3g vs oracle(on device LT_LS), %.3g vs oracle(on oracle LT_LS)" % (e_dev, e_ref))
    assert e_dev < ftol and e_ref < ftol
    ref_eq = oracle.equalize(rx_symb, lt.astype(np.complex128), lin.astype(np.complex128))
    assert rel_err(eq.reshape(n, NBLK, NSC), ref_eq, floor=1e-6) < (tol if kind == "freq" else ftol) * 5
