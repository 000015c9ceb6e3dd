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


def test_dropin_cxx_mangled_symbols(wifi):
    """The C++-linkage shim carries the names an UNMODIFIED reference object file asks for: the reference is C++ compiled by
    g++ / mpiCC (compile.c:25-30), so main.c / main_openmp.c reference the mangled utils.h:38-60 and main.c:4-8 names."""
    so = os.path.join(PKG, "libwifi_dropin_cxx.so")
    if not os.path.exists(so):
        pytest.skip("C++ drop-in shim not built")
    mangled = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
    for name in ("_Z7inversePPCeiS1_", "_Z8multiplyPPCeiiS1_iiS1_", "_Z9hermitianPPCeiiS1_"):          # SURVEY 8(b), nm of the rebuilt utils.o
        assert re.search(r"\bT %s\b" % name, mangled), name
    demangled = subprocess.run(["nm", "-DC", "--defined-only", so], capture_output=True, text=True).stdout
    ldc = r"long double _Complex"
    for proto in (r"hermitian\(%s\*\*, int, int, %s\*\*\)" % (ldc, ldc),
                  r"multiply\(%s\*\*, int, int, %s\*\*, int, int, %s\*\*\)" % (ldc, ldc, ldc),
                  r"multiplyVxVeqM\(%s\*\*, int, int, %s\*\*, int, int, %s\*\*\)" % (ldc, ldc, ldc),
                  r"identity\(%s\*\*, int, double\)" % ldc,
                  r"addition\(%s\*\*, int, int, %s\*\*, int, int, %s\*\*\)" % (ldc, ldc, ldc),
                  r"inverse\(%s\*\*, int, %s\*\*\)" % (ldc, ldc),
                  r"WiFi_channel_estimation_LT_LS\(%s\*, %s\*, %s\*\)" % (ldc, ldc, ldc),
                  r"WiFi_channel_estimation_PS_Linear\(%s\*, %s\*, %s\*\)" % (ldc, ldc, ldc),
                  r"WiFi_channel_estimation_PS_Cubic\(%s\*, %s\*, %s\*\)" % (ldc, ldc, ldc),
                  r"WiFi_channel_estimation_PS_Sinc\(%s\*, %s\*, %s\*\)" % (ldc, ldc, ldc),
                  r"WiFi_channel_estimation_PS_MMSE\(%s\*, %s\*, %s\*\*, double, %s\*, %s\*\)" % (ldc, ldc, ldc, ldc, ldc)):
        assert re.search(r"\bT " + proto, demangled), proto


REF = "/root/reference"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main.c")), reason="the reference is not mounted on this box")
def test_unmodified_reference_main_links_against_the_cxx_dropin(wifi, tmp_path):
    """the reference's own main.c, compiled as it is (g++ -std=gnu++98, stub mpi.h because utils.h:1
    includes it), links against libwifi_dropin_cxx.so with NO utils.o -- every estimator and utils symbol it needs comes from
    the drop-in.  (Running it needs a GPU; tests/test_gpu_host_driver.py does that on the B200 box.)"""
    so = os.path.join(PKG, "libwifi_dropin_cxx.so")
    if not os.path.exists(so):
        pytest.skip("C++ drop-in shim not built")
    obj, exe = str(tmp_path / "main.o"), str(tmp_path / "main")
    subprocess.check_call(["g++", "-std=gnu++98", "-w", "-I" + os.path.join(ROOT, "oracle", "stub"), "-I" + REF, "-c", os.path.join(REF, "main.c"), "-o", obj])
    r = subprocess.run(["g++", "-o", exe, obj, "-L" + PKG, "-lwifi_dropin_cxx", "-lwifi_dropin", "-lwifi_b200", "-Wl,-rpath," + PKG, "-lm"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    dyn = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    for name in ("_Z8multiplyPPCeiiS1_iiS1_", "_Z7inversePPCeiS1_", "_Z9hermitianPPCeiiS1_"):       # main.c:183-205 calls them
        assert name in dyn, name
