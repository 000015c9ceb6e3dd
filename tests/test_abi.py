"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads, and exports exactly the symbols
include/wifi_b200.h declares; the product fails loudly without a GPU; nothing in the package touches oracle/."""
import ctypes as C
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "80211parallelestimation_b200")


@pytest.fixture(scope="module")
def wifi():
    mod = importlib.import_module("80211parallelestimation_b200")
    from importlib import import_module
    import_module("80211parallelestimation_b200.build").build()
    return mod


def header_symbols(name):
    text = open(os.path.join(ROOT, "include", name)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wifi_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(wifi):
    lib = wifi._lib.load()
    declared = header_symbols("wifi_b200.h")
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), "declared in include/wifi_b200.h but not exported: " + name
    # and the ctypes signature table covers the whole header
    assert sorted(wifi._lib.SIGNATURES) == declared


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", os.path.join(PKG, "libwifi_b200.so")], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_gpu_fails_loudly(wifi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(wifi.WifiError, match="no CPU fallback"):
        wifi.WifiContext(0)
    h = C.c_void_p()
    assert wifi._lib.load().wifi_create(0, C.byref(h)) == wifi._lib.ERR_NO_DEVICE


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), "%s mentions the oracle" % f


def test_dropin_symbols(wifi):
    """The C99 drop-in shim exports the reference's own names (main.c:4-8, utils.h:38-60)."""
    so = os.path.join(PKG, "libwifi_dropin.so")
    if not os.path.exists(so):
        pytest.skip("drop-in shim not built")
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    for name in ("WiFi_channel_estimation_LT_LS", "WiFi_channel_estimation_PS_Linear", "WiFi_channel_estimation_PS_Cubic",
                 "WiFi_channel_estimation_PS_Sinc", "WiFi_channel_estimation_PS_MMSE", "hermitian", "multiply",
                 "multiplyVxVeqM", "identity", "addition", "inverse"):
        assert re.search(r"\bT %s\b" % name, out), name
