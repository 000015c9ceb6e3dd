"""In-tree build of the CUDA library (sm_100a only) and the reference-named drop-in shims.

    python 80211parallelestimation_b200/build.py [--force]

Outputs (git-ignored, shipped to the GPU box with the snapshot):
    80211parallelestimation_b200/libwifi_b200.so        kernels + C-ABI (include/wifi_b200.h)
    80211parallelestimation_b200/libwifi_dropin.so      C99 single-frame drop-ins (include/wifi_dropin.h)
    80211parallelestimation_b200/libwifi_dropin_cxx.so  the same entry points with the reference's C++ linkage
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
INC = os.path.join(os.path.dirname(HERE), "include")
LIB = os.path.join(HERE, "libwifi_b200.so")
DROPIN = os.path.join(HERE, "libwifi_dropin.so")
DROPIN_CXX = os.path.join(HERE, "libwifi_dropin_cxx.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets"]
CU = ["wifi_ls.cu", "wifi_frontend.cu", "wifi_solve.cu", "wifi_inverse_tc.cu", "wifi_solve_hpd.cu", "wifi_gemm.cu", "wifi_gemm_tc.cu", "wifi_gemm_dmma.cu", "wifi_eig.cu", "wifi_lowrank.cu", "wifi_synth.cu", "wifi_peaks.cu", "wifi_capi.cu"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))] + \
              [os.path.join(INC, f) for f in os.listdir(INC)]
    cus = [f for f in CU if os.path.exists(os.path.join(CSRC, f))]
    jobs = []
    for f in cus:
        src, obj = os.path.join(CSRC, f), os.path.join(OBJ, f[:-3] + ".o")
        if force or _newer(obj, [src] + headers):
            jobs.append([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=8) as ex:
        for out in ex.map(_run, jobs):
            if verbose and out.strip():
                print(out)
    objs = [os.path.join(OBJ, f[:-3] + ".o") for f in cus]
    if force or jobs or _newer(LIB, objs):
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs)
    c_src = os.path.join(CSRC, "wifi_dropin.c")
    if os.path.exists(c_src) and (force or _newer(DROPIN, [c_src, LIB] + headers)):
        _run(["gcc", "-std=gnu99", "-O2", "-fPIC", "-shared", "-I" + INC, "-o", DROPIN, c_src,
              "-L" + HERE, "-lwifi_b200", "-Wl,-rpath,$ORIGIN", "-lm"])
    cxx_src = os.path.join(CSRC, "wifi_dropin_cxx.cpp")
    if os.path.exists(cxx_src) and (force or _newer(DROPIN_CXX, [cxx_src, DROPIN] + headers)):
        _run(["g++", "-std=gnu++98", "-w", "-O2", "-fPIC", "-shared", "-I" + INC, "-o", DROPIN_CXX, cxx_src,
              "-L" + HERE, "-lwifi_dropin", "-lwifi_b200", "-Wl,-rpath,$ORIGIN", "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
