"""ctypes bindings for the CPU ORACLE (test infrastructure, not product code).

Loads oracle/libwifi_oracle.so (the C99 long-double restatement, wifi_oracle.c) and,
when present, oracle/_ref/libwifi_ref.so (the reference's own sequential code compiled
in place by oracle/Makefile).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
legs may import this module; the product package never does.

All arrays are numpy complex128 (interleaved re/im doubles in memory, which is what the
C side takes); results are complex128 rounded from long double.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NSC, NBLK, PILOTS, DC = 53, 15, (5, 19, 33, 47), 26

_dp = C.POINTER(C.c_double)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    so = os.path.join(HERE, "libwifi_oracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(os.path.join(HERE, "wifi_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    elif os.path.exists("/root/reference/main.c") and not all(
            os.path.exists(os.path.join(HERE, "_ref", n)) for n in ("libwifi_ref.so", "libwifi_ref_O0.so", "libwifi_ref_omp.so")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.ctypes.data_as(_dp)


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class Oracle:
    def __init__(self):
        build()
        self.lib = C.CDLL(os.path.join(HERE, "libwifi_oracle.so"))
        assert self.lib.orc_sizeof_long_double() == 16, "oracle expects x87 80-bit long double"

    # ---- estimators, C semantics ----
    def lt_ls(self, tx_pre, rx_pre):
        tx, ptx = _c(tx_pre); rx, prx = _c(rx_pre)
        n = tx.size // NSC
        H = np.empty((n, NSC), np.complex128)
        self.lib.orc_lt_ls(ptx, prx, H.ctypes.data_as(_dp), C.c_long(n))
        return H.reshape(tx.shape)

    def _ps(self, fn, tx, rx, frame_stride=None):
        tx, ptx = _c(tx); rx, prx = _c(rx)
        if frame_stride is None:
            frame_stride = NSC
        n = tx.size // frame_stride
        H = np.empty((n, NSC), np.complex128)
        fn(ptx, prx, C.c_long(frame_stride), H.ctypes.data_as(_dp), C.c_long(n))
        return H.reshape(NSC) if tx.shape == (NSC,) else H

    def ps_linear(self, tx, rx, frame_stride=None):
        return self._ps(self.lib.orc_ps_linear, tx, rx, frame_stride)

    def ps_cubic(self, tx, rx, frame_stride=None):
        return self._ps(self.lib.orc_ps_cubic, tx, rx, frame_stride)

    def ps_sinc(self, tx, rx, frame_stride=None):
        return self._ps(self.lib.orc_ps_sinc, tx, rx, frame_stride)

    def ps_matlab(self, which, tx_frames, rx_frames):
        tx, ptx = _c(tx_frames); rx, prx = _c(rx_frames)
        n = tx.size // (NSC * NBLK)
        H = np.empty((n, NSC), np.complex128)
        self.lib.orc_ps_matlab(C.c_int({"linear": 0, "cubic": 1, "sinc": 2}[which]), ptx, prx, H.ctypes.data_as(_dp), C.c_long(n))
        return H

    def frontend(self, packet, lptot):
        """WiFi_blocks_extraction.m + WiFi_RX.m:19-31: time samples -> (symb [n][15][53], pre_fft [n][53], ow2 [n])."""
        pk, ppk = _c(packet); lp, plp = _c(lptot)
        n = pk.size // 1200
        symb = np.empty((n, NBLK, NSC), np.complex128); pre = np.empty((n, NSC), np.complex128); ow2 = np.empty(n, np.float64)
        self.lib.orc_frontend(ppk, plp, symb.ctypes.data_as(_dp), pre.ctypes.data_as(_dp), ow2.ctypes.data_as(_dp), C.c_long(n))
        return symb, pre, ow2

    def equalize(self, rx_frames, H_lt, H_ps):
        rx, prx = _c(rx_frames); a, pa = _c(H_lt); b, pb = _c(H_ps)
        n = rx.size // (NSC * NBLK)
        eq = np.empty((n, NBLK, NSC), np.complex128)
        self.lib.orc_equalize(prx, pa, pb, eq.ctypes.data_as(_dp), C.c_long(n))
        return eq

    # ---- utils.c ----
    def multiply(self, M1, M2):
        a, pa = _c(M1); b, pb = _c(M2)
        res = np.empty((a.shape[0], b.shape[1]), np.complex128)
        rc = self.lib.orc_multiply(pa, a.shape[0], a.shape[1], pb, b.shape[0], b.shape[1], res.ctypes.data_as(_dp))
        if rc:
            raise ValueError("Matrices dimension missmatch")
        return res

    def hermitian_as_written(self, M):
        a, pa = _c(M)
        res = np.empty((a.shape[1], a.shape[0]), np.complex128)
        self.lib.orc_hermitian_as_written(pa, a.shape[0], a.shape[1], res.ctypes.data_as(_dp))
        return res

    def conj_transpose(self, M):
        a, pa = _c(M)
        res = np.empty((a.shape[1], a.shape[0]), np.complex128)
        self.lib.orc_conj_transpose(pa, a.shape[0], a.shape[1], res.ctypes.data_as(_dp))
        return res

    def outer(self, M1, M2):
        a, pa = _c(M1); b, pb = _c(M2)
        res = np.empty((a.shape[0], b.shape[1]), np.complex128)
        rc = self.lib.orc_outer(pa, a.shape[0], a.shape[1], pb, b.shape[0], b.shape[1], res.ctypes.data_as(_dp))
        if rc:
            raise ValueError("Matrices dimension missmatch")
        return res

    def identity(self, size, scalar):
        res = np.empty((size, size), np.complex128)
        self.lib.orc_identity(res.ctypes.data_as(_dp), size, C.c_double(scalar))
        return res

    def addition_as_written(self, M1, M2):
        a, pa = _c(M1); b, pb = _c(M2)
        res = np.empty(a.shape, np.complex128)
        rc = self.lib.orc_addition_as_written(pa, a.shape[0], a.shape[1], pb, b.shape[0], b.shape[1], res.ctypes.data_as(_dp))
        if rc:
            raise ValueError("Matrices dimension missmatch")
        return res

    def add(self, M1, M2):
        a, pa = _c(M1); b, pb = _c(M2)
        res = np.empty(a.shape, np.complex128)
        rc = self.lib.orc_add(pa, a.shape[0], a.shape[1], pb, b.shape[0], b.shape[1], res.ctypes.data_as(_dp))
        if rc:
            raise ValueError("Matrices dimension missmatch")
        return res

    def inverse_cofactor(self, A):
        a, pa = _c(A)
        Y = np.empty(a.shape, np.complex128)
        rc = self.lib.orc_inverse_cofactor(pa, a.shape[0], Y.ctypes.data_as(_dp))
        assert rc == 0
        return Y

    def inverse_gj(self, A):
        a, pa = _c(A)
        Y = np.empty(a.shape, np.complex128)
        rc = self.lib.orc_inverse_gj(pa, a.shape[0], Y.ctypes.data_as(_dp))
        if rc:
            raise np.linalg.LinAlgError("singular")
        return Y

    # ---- intended MMSE ----
    def mmse_filter(self, R, d):
        r, pr = _c(R); dd, pd = _d(d)
        W = np.empty((NSC, NSC), np.complex128)
        rc = self.lib.orc_mmse_filter(pr, pd, W.ctypes.data_as(_dp))
        assert rc == 0
        return W

    def mmse_apply(self, W, H_ls):
        w, pw = _c(W); h, ph = _c(H_ls)
        n = h.size // NSC
        H = np.empty((n, NSC), np.complex128)
        self.lib.orc_mmse_apply(pw, ph, H.ctypes.data_as(_dp), C.c_long(n))
        return H.reshape(h.shape)

    def mmse_perframe(self, R, tx, rx, sigma2):
        r, pr = _c(R); t, pt = _c(tx); x, px = _c(rx)
        n = t.size // NSC
        s, ps = _d(np.broadcast_to(np.asarray(sigma2, np.float64), (n,)))
        H = np.empty((n, NSC), np.complex128)
        rc = self.lib.orc_mmse_perframe(pr, pt, px, ps, H.ctypes.data_as(_dp), C.c_long(n))
        assert rc == 0
        return H

    def mmse_cconv_batch(self, tx, rx, ow2, H_ls):
        t, pt = _c(tx); x, px = _c(rx); h, ph = _c(H_ls)
        n = t.size // NSC
        s, ps = _d(np.broadcast_to(np.asarray(ow2, np.float64), (n,)))
        H = np.empty((n, NSC), np.complex128)
        rc = self.lib.orc_mmse_cconv_batch(pt, px, ps, ph, H.ctypes.data_as(_dp), C.c_long(n))
        assert rc == 0
        return H

    def mmse_cconv(self, tx, rx, ow2, H_ls):
        t, pt = _c(tx); x, px = _c(rx); h, ph = _c(H_ls)
        H = np.empty(NSC, np.complex128)
        rc = self.lib.orc_mmse_cconv(pt, px, C.c_double(ow2), ph, H.ctypes.data_as(_dp))
        assert rc == 0
        return H

    def mmse_matlab_block(self, tx, rx, ow2, H_ls):
        t, pt = _c(tx); x, px = _c(rx); h, ph = _c(H_ls)
        H = np.empty(NSC, np.complex128)
        rc = self.lib.orc_mmse_matlab_block(pt, px, C.c_double(ow2), ph, H.ctypes.data_as(_dp))
        assert rc == 0
        return H

    def mmse_rank1(self, tx, rx, ow2, H_ls):
        t, pt = _c(tx); x, px = _c(rx); h, ph = _c(H_ls)
        H = np.empty(NSC, np.complex128)
        self.lib.orc_mmse_rank1(pt, px, C.c_double(ow2), ph, H.ctypes.data_as(_dp))
        return H


class Reference:
    """The reference's own compiled sequential code (oracle/_ref/libwifi_ref.so, -O2; opt="O0": the same sources at -O0, which
    is what the reference's build helper compile.c:25-30 produces)."""

    path = os.path.join(HERE, "_ref", "libwifi_ref.so")
    path_O0 = os.path.join(HERE, "_ref", "libwifi_ref_O0.so")

    @classmethod
    def available(cls, opt="O2"):
        p = cls.path if opt == "O2" else cls.path_O0
        if not os.path.exists(p) and os.path.exists("/root/reference/main.c"):
            build()
        return os.path.exists(p)

    def __init__(self, opt="O2"):
        if not self.available(opt):
            raise RuntimeError("oracle/_ref/libwifi_ref%s.so not built (needs /root/reference)" % ("" if opt == "O2" else "_O0"))
        self.lib = C.CDLL(self.path if opt == "O2" else self.path_O0)
        self.lib.ref_ow2.restype = C.c_double
        self.lib.ref_sinc.restype = C.c_double

    def set_threads(self, n):
        """OpenMP team size of the *_omp timing loops; returns the size in effect."""
        self.lib.ref_set_threads(C.c_int(int(n)))
        return int(self.lib.ref_get_max_threads())

    def estimate(self, which, a, b, omp=False):
        code = {"lt_ls": 0, "ps_linear": 1, "ps_cubic": 2, "ps_sinc": 3}[which]
        x, px = _c(a); y, py = _c(b)
        n = x.size // NSC
        H = np.empty((n, NSC), np.complex128)
        fn = self.lib.ref_estimate_omp if omp else self.lib.ref_estimate
        fn(C.c_int(code), px, py, H.ctypes.data_as(_dp), C.c_long(n))
        return H.reshape(x.shape)

    def mmse_as_written(self, tx, rx, ow2, H_ls):
        t, pt = _c(tx); x, px = _c(rx); h, ph = _c(H_ls)
        H = np.empty(NSC, np.complex128)
        self.lib.ref_mmse_as_written(pt, px, C.c_double(ow2), ph, H.ctypes.data_as(_dp))
        return H

    def mmse_shared_omp(self, W, tx, rx):
        w, pw = _c(W); t, pt = _c(tx); x, px = _c(rx)
        n = t.size // NSC
        H = np.empty((n, NSC), np.complex128)
        self.lib.ref_mmse_shared_omp(pw, pt, px, H.ctypes.data_as(_dp), C.c_long(n))
        return H

    def _mm(self, fn, M1, M2, shape):
        a, pa = _c(M1); b, pb = _c(M2)
        res = np.full(shape, np.nan + 0j, np.complex128)
        fn(pa, a.shape[0], a.shape[1], pb, b.shape[0], b.shape[1], res.ctypes.data_as(_dp))
        return res

    def multiply(self, M1, M2):
        return self._mm(self.lib.ref_multiply, M1, M2, (np.shape(M1)[0], np.shape(M2)[1]))

    def outer(self, M1, M2):
        return self._mm(self.lib.ref_outer, M1, M2, (np.shape(M1)[0], np.shape(M2)[1]))

    def addition(self, M1, M2):
        return self._mm(self.lib.ref_addition, M1, M2, np.shape(M1))

    def hermitian(self, M):
        a, pa = _c(M)
        res = np.empty((a.shape[1], a.shape[0]), np.complex128)
        self.lib.ref_hermitian(pa, a.shape[0], a.shape[1], res.ctypes.data_as(_dp))
        return res

    def identity(self, size, scalar):
        res = np.empty((size, size), np.complex128)
        self.lib.ref_identity(res.ctypes.data_as(_dp), size, C.c_double(scalar))
        return res

    def inverse(self, A):
        a, pa = _c(A)
        Y = np.empty(a.shape, np.complex128)
        self.lib.ref_inverse(pa, a.shape[0], Y.ctypes.data_as(_dp))
        return Y

    def sinc(self, x):
        return self.lib.ref_sinc(C.c_double(x))

    def inputs(self):
        tp = np.empty(NSC, np.complex128); rp = np.empty(NSC, np.complex128)
        ts = np.empty(NSC * NBLK, np.complex128); rs = np.empty(NSC * NBLK, np.complex128)
        self.lib.ref_inputs(*(a.ctypes.data_as(_dp) for a in (tp, rp, ts, rs)))
        return dict(ow2=self.lib.ref_ow2(), tx_preamble_fft=tp, rx_preamble_fft=rp, tx_symb=ts, rx_symb=rs)


class ReferenceOpenMP:
    """The reference's OpenMP build as it is (main_openmp.c:70-176 + the _omp twins of utils.c; oracle/_ref/libwifi_ref_omp.so):
    every estimator call spawns its own 53-thread intra-frame team.  Timed by bench.py on a small N'."""

    path = os.path.join(HERE, "_ref", "libwifi_ref_omp.so")

    @classmethod
    def available(cls):
        if not os.path.exists(cls.path) and os.path.exists("/root/reference/main.c"):
            build()
        return os.path.exists(cls.path)

    def __init__(self):
        if not self.available():
            raise RuntimeError("oracle/_ref/libwifi_ref_omp.so not built (needs /root/reference)")
        self.lib = C.CDLL(self.path)

    def set_threads(self, n):
        self.lib.refomp_set_threads(C.c_int(int(n)))

    def estimate(self, which, a, b):
        code = {"lt_ls": 0, "ps_linear": 1, "ps_cubic": 2, "ps_sinc": 3}[which]
        x, px = _c(a); y, py = _c(b)
        n = x.size // NSC
        H = np.empty((n, NSC), np.complex128)
        self.lib.refomp_estimate(C.c_int(code), px, py, H.ctypes.data_as(_dp), C.c_long(n))
        return H.reshape(x.shape)
