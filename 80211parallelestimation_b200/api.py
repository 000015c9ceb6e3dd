"""Host-side mirror of the reference interface (main.c:4-8, utils.h:38-60, WiFi_Equalization.m) over the C-ABI.

Every method accepts either
  * torch CUDA tensors (complex64 / complex128, contiguous): the `_batch` entry points run on the tensors' memory
    on torch's current stream, results are new CUDA tensors; or
  * numpy arrays (complex64 / complex128): the `_host` entry points copy H2D, run the kernels and copy D2H.
Argument meaning and error behaviour follow the reference: a dimension mismatch raises WifiError("Matrices
dimension missmatch") and writes nothing (utils.c:18-19); singular systems raise instead of returning NaN silently.
There is no CPU implementation behind any of these calls.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import (F32, F64, PS_LINEAR, PS_CUBIC, PS_SINC, PS_MATLAB, SOLVE_PIVOT, SOLVE_HPD, SOLVE_REFINE, SOLVE_WIDE, SOLVE_FAST32, AS_WRITTEN, INTENDED)

NSC, NBLK, FRAME = 53, 15, 795


class WifiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("wifi_b200 error %d: %s" % (code, msg))
        self.code = code


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class _Arg:
    """A complex/real array crossing the C-ABI: pointer + dtype code + a reference that keeps it alive."""

    def __init__(self, x, real=False):
        if _is_torch(x):
            import torch
            if not x.is_cuda:
                raise TypeError("torch tensors must live on the GPU (pass numpy arrays for host data)")
            want = (torch.float32, torch.float64) if real else (torch.complex64, torch.complex128)
            if x.dtype not in want:
                raise TypeError("unsupported dtype %s" % x.dtype)
            x = x.contiguous()
            self.dt = F32 if x.dtype in (torch.complex64, torch.float32) else F64
            self.ptr = x.data_ptr()
            self.device = True
        else:
            x = np.asarray(x)
            if real:
                if x.dtype not in (np.float32, np.float64):
                    x = x.astype(np.float64)
            elif x.dtype not in (np.complex64, np.complex128):
                x = x.astype(np.complex128)
            x = np.ascontiguousarray(x)
            self.dt = F32 if x.dtype in (np.complex64, np.float32) else F64
            self.ptr = x.ctypes.data
            self.device = False
        self.x = x
        self.shape = tuple(x.shape)
        self.size = int(np.prod(self.shape)) if self.shape else 1

    def empty_like(self, shape, real=False):
        if self.device:
            import torch
            cd = {F32: torch.float32 if real else torch.complex64, F64: torch.float64 if real else torch.complex128}[self.dt]
            return torch.empty(shape, dtype=cd, device=self.x.device)
        cd = {F32: np.float32 if real else np.complex64, F64: np.float64 if real else np.complex128}[self.dt]
        return np.empty(shape, dtype=cd)


def _ptr(x):
    if x is None:
        return None
    return x.data_ptr() if _is_torch(x) else x.ctypes.data


def _same(*args):
    dev = {a.device for a in args}
    dts = {a.dt for a in args}
    if len(dev) != 1 or len(dts) != 1:
        raise TypeError("all arrays of one call must share residency (host/device) and precision")
    return args[0].device, args[0].dt


class WifiContext:
    """One context per GPU (wifi_create).  Not thread-safe per context; create one per thread/stream."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.wifi_create(int(device), C.byref(h))
        if rc != _lib.OK:
            raise WifiError(rc, "wifi_create(%d) failed (%s): no CUDA device means no service, there is no CPU fallback"
                            % (device, {5: "no CUDA device", 2: "CUDA error", 1: "bad device index"}.get(rc, "?")))
        self.h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.wifi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing ----
    def _ck(self, rc):
        if rc != _lib.OK:
            raise WifiError(rc, self.lib.wifi_last_error(self.h).decode())

    def _sync_stream(self, device_call):
        if device_call:
            import torch
            self.lib.wifi_set_stream(self.h, C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))

    def set_host_chunk_bytes(self, nbytes):
        """Per-array staging target of one chunk of the host-pointer pipeline (default 48 MiB)."""
        self._ck(self.lib.wifi_set_host_chunk_bytes(self.h, int(nbytes)))

    def synchronize(self):
        self._ck(self.lib.wifi_synchronize(self.h))

    @property
    def launches(self):
        return int(self.lib.wifi_launch_count(self.h))

    def enable_kernel_timing(self, on=True):
        self._ck(self.lib.wifi_enable_kernel_timing(self.h, int(on)))

    def last_kernel_ms(self):
        ms = C.c_float()
        self._ck(self.lib.wifi_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    # ---- estimators ----
    def lt_ls(self, tx_pre, rx_pre, out=None):
        """WiFi_channel_estimation_LT_LS (main.c:66-75) over [n][53] preambles."""
        tx, rx = _Arg(tx_pre), _Arg(rx_pre)
        dev, dt = _same(tx, rx)
        n = tx.size // NSC
        H = out if out is not None else tx.empty_like(tx.shape)
        self._sync_stream(dev)
        fn = self.lib.wifi_lt_ls_batch if dev else self.lib.wifi_lt_ls_host
        self._ck(fn(self.h, dt, tx.ptr, rx.ptr, _ptr(H), n))
        return H

    def ps(self, tx_symbols, rx_symbols, which=("linear", "cubic", "sinc"), frame_stride=None, n_frames=None, out=None, matlab=False):
        """PS_Linear / PS_Cubic / PS_Sinc (main.c:77-146) fused over one pilot-LS pass.  tx/rx: [n][53] block vectors
        (frame_stride 53) or whole frames [n][15][53] with frame_stride=795 (block 0 is used, main.c:30-33).
        matlab=True: WiFi_channel_estimation_PS_*.m semantics (average over blocks 1..4 of whole frames, true cubic spans)."""
        tx, rx = _Arg(tx_symbols), _Arg(rx_symbols)
        dev, dt = _same(tx, rx)
        if frame_stride is None:
            frame_stride = FRAME if (len(tx.shape) == 3 and tx.shape[-2:] == (NBLK, NSC)) else NSC
        if n_frames is None:
            n_frames = tx.size // frame_stride
        mask = sum({"linear": PS_LINEAR, "cubic": PS_CUBIC, "sinc": PS_SINC}[w] for w in which) | (PS_MATLAB if matlab else 0)
        outs = {w: (out[w] if out else tx.empty_like((n_frames, NSC))) for w in which}
        self._sync_stream(dev)
        fn = self.lib.wifi_ps_batch if dev else self.lib.wifi_ps_host
        self._ck(fn(self.h, dt, mask, tx.ptr, rx.ptr, frame_stride, _ptr(outs.get("linear")), _ptr(outs.get("cubic")),
                    _ptr(outs.get("sinc")), n_frames))
        return outs

    def ps_linear(self, tx, rx, **kw):
        return self.ps(tx, rx, ("linear",), **kw)["linear"]

    def ps_cubic(self, tx, rx, **kw):
        return self.ps(tx, rx, ("cubic",), **kw)["cubic"]

    def ps_sinc(self, tx, rx, **kw):
        return self.ps(tx, rx, ("sinc",), **kw)["sinc"]

    def estimate_all(self, tx_pre, rx_pre, tx_symbols, rx_symbols, frame_stride=None, n_frames=None, equalize=True, out=None):
        """All five estimators (LT_LS, PS_Linear, PS_Cubic, PS_Sinc, shared-filter PS_MMSE) and, for whole frames, the equalizer
        (WiFi_RX.m:47-60) in one call.  Returns a dict with keys lt_ls, linear, cubic, sinc, mmse (+ eq)."""
        tp, rp, tx, rx = _Arg(tx_pre), _Arg(rx_pre), _Arg(tx_symbols), _Arg(rx_symbols)
        dev, dt = _same(tp, rp, tx, rx)
        if frame_stride is None:
            frame_stride = FRAME if (len(tx.shape) == 3 and tx.shape[-2:] == (NBLK, NSC)) else NSC
        if n_frames is None:
            n_frames = tp.size // NSC
        do_eq = bool(equalize) and frame_stride == FRAME
        o = dict(out) if out else {}
        for k in ("lt_ls", "linear", "cubic", "sinc", "mmse"):
            if k not in o:
                o[k] = tp.empty_like((n_frames, NSC))
        if do_eq and "eq" not in o:
            o["eq"] = tp.empty_like((n_frames, NBLK, NSC))
        self._sync_stream(dev)
        if dev:
            self._ck(self.lib.wifi_estimate_all_batch(self.h, dt, tp.ptr, rp.ptr, tx.ptr, rx.ptr, frame_stride, _ptr(o["lt_ls"]), _ptr(o["linear"]),
                                                      _ptr(o["cubic"]), _ptr(o["sinc"]), _ptr(o["mmse"]), _ptr(o["eq"]) if do_eq else None, n_frames))
        else:
            if frame_stride != FRAME:
                raise TypeError("host arrays: pass whole frames [n][15][53]")
            self._ck(self.lib.wifi_estimate_all_host(self.h, dt, tp.ptr, rp.ptr, tx.ptr, rx.ptr, _ptr(o["lt_ls"]), _ptr(o["linear"]), _ptr(o["cubic"]),
                                                     _ptr(o["sinc"]), _ptr(o["mmse"]), _ptr(o["eq"]) if do_eq else None, n_frames))
        return o

    def frontend(self, packet, lptot, want_ow2=True, out=None):
        """Receiver front-end (WiFi_blocks_extraction.m, WiFi_RX.m:19-31) for one side of n frames: packet [n][1200],
        lptot [n][160] -> (symb [n][15][53], pre_fft [n][53], ow2 [n] or None)."""
        pk, lp = _Arg(packet), _Arg(lptot)
        dev, dt = _same(pk, lp)
        n = pk.size // 1200
        if lp.size != n * 160:
            raise ValueError("packet is [n][1200] and lptot [n][160] for the same n")
        if out is not None:
            symb, pre, ow2 = out
        else:
            symb, pre = pk.empty_like((n, NBLK, NSC)), pk.empty_like((n, NSC))
            ow2 = pk.empty_like((n,), real=True) if want_ow2 else None
        self._sync_stream(dev)
        fn = self.lib.wifi_frontend_batch if dev else self.lib.wifi_frontend_host
        self._ck(fn(self.h, dt, pk.ptr, lp.ptr, _ptr(symb), _ptr(pre), _ptr(ow2), n))
        return symb, pre, ow2

    RX_CHAIN_PLANES = {"lt_ls": "H_lt", "linear": "H_linear", "cubic": "H_cubic", "sinc": "H_sinc", "mmse_cconv": "H_mmse_cconv",
                       "ls0": "H_ls0", "mmse": "H_mmse_shared", "eq": "eq", "rx_symb": "rx_symb", "ow2": "ow2"}

    def rx_chain(self, tx_packet, tx_lptot, rx_packet, rx_lptot, want=("lt_ls", "linear", "cubic", "sinc", "mmse_cconv", "eq", "ow2"), out=None):
        """Fused receiver chain (WiFi_RX.m:17-60 with the C estimators of main.c:66-146): time samples of n frames -> the planes
        named in `want`, one launch (two with "mmse", the shared-filter PS_MMSE).  Packets [n][1200], lptot [n][160]."""
        tp, tl, rp, rl = _Arg(tx_packet), _Arg(tx_lptot), _Arg(rx_packet), _Arg(rx_lptot)
        dev, dt = _same(tp, tl, rp, rl)
        n = rp.size // 1200
        if tp.size != n * 1200 or tl.size != n * 160 or rl.size != n * 160:
            raise ValueError("packets are [n][1200] and lptot [n][160] for the same n")
        o = dict(out) if out is not None else {}
        st = _lib.RxChainOut()
        for name in want:
            if name not in o:
                o[name] = rp.empty_like((n,), real=True) if name == "ow2" else rp.empty_like((n, NBLK, NSC) if name in ("eq", "rx_symb") else (n, NSC))
            setattr(st, self.RX_CHAIN_PLANES[name], _ptr(o[name]))
        self._sync_stream(dev)
        fn = self.lib.wifi_rx_chain_batch if dev else self.lib.wifi_rx_chain_host
        self._ck(fn(self.h, dt, tp.ptr, tl.ptr, rp.ptr, rl.ptr, C.byref(st), n))
        return o

    def equalize(self, rx_frames, H_lt, H_ps, out=None):
        """WiFi_Equalization.m: rx [n][15][53], H_lt / H_ps [n][53] -> [n][15][53]."""
        rx, a, b = _Arg(rx_frames), _Arg(H_lt), _Arg(H_ps)
        dev, dt = _same(rx, a, b)
        n = rx.size // FRAME
        eq = out if out is not None else rx.empty_like(rx.shape)
        self._sync_stream(dev)
        fn = self.lib.wifi_equalize_batch if dev else self.lib.wifi_equalize_host
        self._ck(fn(self.h, dt, rx.ptr, a.ptr, b.ptr, _ptr(eq), n))
        return eq

    # ---- MMSE ----
    def mmse_filter_form(self, R, d, want_W=True):
        """W = R (R + diag d)^-1 in FP64 on the device; installs W for mmse_shared*.  R: 53x53 complex128, d: 53 float64."""
        r = _Arg(R)
        dd = _Arg(d, real=True)
        if r.dt != F64 or dd.dt != F64 or r.shape != (NSC, NSC) or dd.size != NSC:
            raise TypeError("R must be 53x53 complex128 and d 53 float64")
        if r.device != dd.device:
            raise TypeError("R and d must share residency")
        W = r.empty_like((NSC, NSC)) if want_W else None
        self._sync_stream(r.device)
        fn = self.lib.wifi_mmse_filter_form if r.device else self.lib.wifi_mmse_filter_form_host
        self._ck(fn(self.h, r.ptr, dd.ptr, _ptr(W)))
        return W

    def mmse_filter_set(self, W):
        w = _Arg(W)
        if not w.device or w.dt != F64 or w.shape != (NSC, NSC):
            raise TypeError("W must be a 53x53 complex128 CUDA tensor")
        self._sync_stream(True)
        self._ck(self.lib.wifi_mmse_filter_set(self.h, w.ptr))

    def mmse_shared_apply(self, H_ls, out=None):
        h = _Arg(H_ls)
        if not h.device:
            raise TypeError("mmse_shared_apply takes CUDA tensors (use mmse_shared for host arrays)")
        H = out if out is not None else h.empty_like(h.shape)
        self._sync_stream(True)
        self._ck(self.lib.wifi_mmse_shared_apply_batch(self.h, h.dt, h.ptr, _ptr(H), h.size // NSC))
        return H

    def mmse_shared(self, tx_symbols, rx_symbols, frame_stride=NSC, n_frames=None, out=None):
        """Shared-filter PS_MMSE: per-block LS divide fused with the GEMM H = (rx/tx) W^T."""
        tx, rx = _Arg(tx_symbols), _Arg(rx_symbols)
        dev, dt = _same(tx, rx)
        if n_frames is None:
            n_frames = tx.size // frame_stride
        H = out if out is not None else tx.empty_like((n_frames, NSC))
        self._sync_stream(dev)
        fn = self.lib.wifi_mmse_shared_batch if dev else self.lib.wifi_mmse_shared_host
        self._ck(fn(self.h, dt, tx.ptr, rx.ptr, frame_stride, _ptr(H), n_frames))
        return H

    def mmse_filter_fold_tx(self, tx_block):
        """Shared, known tx block vector [53]: fold rx/tx into the installed filter (W' = W diag(1/tx)) for mmse_shared_rx."""
        if _is_torch(tx_block):
            import torch
            t = tx_block.to(torch.complex128).contiguous()
            self._sync_stream(True)
            self._ck(self.lib.wifi_mmse_filter_fold_tx(self.h, t.data_ptr()))
        else:
            t = np.ascontiguousarray(np.asarray(tx_block, np.complex128))
            self._ck(self.lib.wifi_mmse_filter_fold_tx_host(self.h, t.ctypes.data))

    def mmse_shared_rx(self, rx_symbols, frame_stride=NSC, n_frames=None, out=None):
        """Shared-filter PS_MMSE for frames that share their (known) tx block vector: H = rx W'^T, only rx is read."""
        rx = _Arg(rx_symbols)
        if n_frames is None:
            n_frames = rx.size // frame_stride
        H = out if out is not None else rx.empty_like((n_frames, NSC))
        self._sync_stream(rx.device)
        fn = self.lib.wifi_mmse_shared_rx_batch if rx.device else self.lib.wifi_mmse_shared_rx_host
        self._ck(fn(self.h, rx.dt, rx.ptr, frame_stride, _ptr(H), n_frames))
        return H

    def pcie_probe(self, h_src, h_dst):
        """Wall time (ms) of one pinned H2D copy of h_src and one D2H copy into h_dst issued together (either may be None)."""
        ms = C.c_double()
        self._ck(self.lib.wifi_pcie_probe(self.h, _ptr(h_src), _ptr(h_dst), h_src.nbytes if h_src is not None else 0,
                                          h_dst.nbytes if h_dst is not None else 0, C.byref(ms)))
        return ms.value

    def mmse_perframe(self, R, tx_symbols, rx_symbols, sigma2, frame_stride=NSC, n_frames=None, flags=SOLVE_HPD, out=None):
        """Per-frame PS_MMSE: A_f = R + diag(sigma2_f/|tx|^2), A_f z = rx/tx, H = R z.  complex64 arrays are solved in FP64
        arithmetic (1e-4-accurate); SOLVE_FAST32 opts into FP32 arithmetic (~4e-3 with SOLVE_HPD, ~1e-1 with SOLVE_PIVOT)."""
        r, tx, rx = _Arg(R), _Arg(tx_symbols), _Arg(rx_symbols)
        dev, dt = _same(r, tx, rx)
        if n_frames is None:
            n_frames = tx.size // frame_stride
        if _is_torch(sigma2):
            s = _Arg(sigma2, real=True)
        else:
            s = _Arg(np.broadcast_to(np.asarray(sigma2, np.float32 if dt == F32 else np.float64), (n_frames,)).copy(), real=True)
            if dev:
                import torch
                s = _Arg(torch.from_numpy(s.x).to(tx.x.device), real=True)
        if s.dt != dt or s.device != dev:
            raise TypeError("sigma2 must match the precision and residency of the frames")
        H = out if out is not None else tx.empty_like((n_frames, NSC))
        self._sync_stream(dev)
        fn = self.lib.wifi_mmse_perframe_batch if dev else self.lib.wifi_mmse_perframe_host
        self._ck(fn(self.h, dt, r.ptr, tx.ptr, rx.ptr, frame_stride, s.ptr, _ptr(H), n_frames, flags))
        return H

    def mmse_eig_prepare(self, R, absx2):
        """Eigen-domain operands for mmse_perframe_eig: R 53x53 complex128, absx2 = the shared |tx_k|^2 pattern (53 float64)."""
        import torch
        r = R if _is_torch(R) else torch.from_numpy(np.ascontiguousarray(np.asarray(R, np.complex128))).cuda(self.device)
        a = absx2 if _is_torch(absx2) else torch.from_numpy(np.ascontiguousarray(np.asarray(absx2, np.float64))).cuda(self.device)
        if r.dtype != torch.complex128 or a.dtype != torch.float64:
            raise TypeError("R must be complex128 and absx2 float64")
        r, a = r.contiguous(), a.contiguous()
        self._sync_stream(True)
        self._ck(self.lib.wifi_mmse_eig_prepare(self.h, r.data_ptr(), a.data_ptr()))

    def mmse_perframe_eig(self, tx_symbols, rx_symbols, sigma2, frame_stride=NSC, n_frames=None, out=None):
        """Per-frame PS_MMSE in the eigen domain (frames share |tx_k|^2; see mmse_eig_prepare)."""
        tx, rx = _Arg(tx_symbols), _Arg(rx_symbols)
        dev, dt = _same(tx, rx)
        if n_frames is None:
            n_frames = tx.size // frame_stride
        s = _Arg(sigma2, real=True)
        if s.dt != dt or s.device != dev:
            raise TypeError("sigma2 must match the precision and residency of the frames")
        H = out if out is not None else tx.empty_like((n_frames, NSC))
        self._sync_stream(dev)
        fn = self.lib.wifi_mmse_perframe_eig_batch if dev else self.lib.wifi_mmse_perframe_eig_host
        self._ck(fn(self.h, dt, tx.ptr, rx.ptr, frame_stride, s.ptr, _ptr(H), n_frames))
        return H

    def mmse_lowrank_prepare(self, R):
        """Operands of mmse_perframe_lowrank from the covariance R (53x53 complex128): returns its numerical rank; raises WifiError
        (INVALID) when the rank is 0 or above LOWRANK_MAX (use mmse_perframe / mmse_perframe_eig then)."""
        import ctypes
        import torch
        r = R if _is_torch(R) else torch.from_numpy(np.ascontiguousarray(np.asarray(R, np.complex128))).cuda(self.device)
        if r.dtype != torch.complex128:
            raise TypeError("R must be complex128")
        r = r.contiguous()
        rank = ctypes.c_int(0)
        self._sync_stream(True)
        self._ck(self.lib.wifi_mmse_lowrank_prepare(self.h, r.data_ptr(), ctypes.addressof(rank)))
        return rank.value

    def mmse_perframe_lowrank(self, tx_symbols, rx_symbols, sigma2, frame_stride=NSC, n_frames=None, out=None):
        """Per-frame PS_MMSE for a low-rank covariance (see mmse_lowrank_prepare): sigma2 and |tx_k|^2 free per frame, one launch."""
        tx, rx = _Arg(tx_symbols), _Arg(rx_symbols)
        dev, dt = _same(tx, rx)
        if n_frames is None:
            n_frames = tx.size // frame_stride
        s = _Arg(sigma2, real=True)
        if s.dt != dt or s.device != dev:
            raise TypeError("sigma2 must match the precision and residency of the frames")
        H = out if out is not None else tx.empty_like((n_frames, NSC))
        self._sync_stream(dev)
        fn = self.lib.wifi_mmse_perframe_lowrank_batch if dev else self.lib.wifi_mmse_perframe_lowrank_host
        self._ck(fn(self.h, dt, tx.ptr, rx.ptr, frame_stride, s.ptr, _ptr(H), n_frames))
        return H

    def mmse_cconv(self, tx_symbols, rx_symbols, ow2, H_ls, out=None):
        """PS_MMSE in the calling convention of main.c:148 (R_f = H_ls H_ls^H), batched over [n][53]."""
        tx, rx, h = _Arg(tx_symbols), _Arg(rx_symbols), _Arg(H_ls)
        dev, dt = _same(tx, rx, h)
        n = tx.size // NSC
        rdt = np.float32 if dt == F32 else np.float64
        s = np.broadcast_to(np.asarray(ow2, rdt), (n,)).copy() if not _is_torch(ow2) else ow2
        if dev and not _is_torch(s):
            import torch
            s = torch.from_numpy(s).to(tx.x.device)
        s = _Arg(s, real=True)
        H = out if out is not None else tx.empty_like(tx.shape)
        self._sync_stream(dev)
        fn = self.lib.wifi_mmse_cconv_batch if dev else self.lib.wifi_mmse_cconv_host
        self._ck(fn(self.h, dt, tx.ptr, rx.ptr, s.ptr, h.ptr, _ptr(H), n))
        return H

    def mmse_matlab(self, tx_frames, rx_frames, ow2, H_ls, out=None):
        """WiFi_channel_estimation_PS_MMSE.m as written, averaged over OFDM blocks 1..4: tx/rx [n][15][53], H_ls [n][53]."""
        tx, rx, h = _Arg(tx_frames), _Arg(rx_frames), _Arg(H_ls)
        dev, dt = _same(tx, rx, h)
        n = tx.size // FRAME
        rdt = np.float32 if dt == F32 else np.float64
        s = np.broadcast_to(np.asarray(ow2, rdt), (n,)).copy() if not _is_torch(ow2) else ow2
        if dev and not _is_torch(s):
            import torch
            s = torch.from_numpy(s).to(tx.x.device)
        s = _Arg(s, real=True)
        H = out if out is not None else h.empty_like((n, NSC))
        self._sync_stream(dev)
        fn = self.lib.wifi_mmse_matlab_batch if dev else self.lib.wifi_mmse_matlab_host
        self._ck(fn(self.h, dt, tx.ptr, rx.ptr, s.ptr, h.ptr, _ptr(H), n))
        return H

    # ---- utils.h ----
    @staticmethod
    def _mat(a):
        shape = a.shape
        if len(shape) == 2:
            return 1, shape[0], shape[1]
        if len(shape) == 3:
            return shape
        raise TypeError("matrices must be [r][c] or [batch][r][c]")

    def multiply(self, M1, M2):
        a, b = _Arg(M1), _Arg(M2)
        dev, dt = _same(a, b)
        ba, r1, c1 = self._mat(a)
        bb, r2, c2 = self._mat(b)
        if ba != bb:
            raise TypeError("batch mismatch")
        res = a.empty_like((r1, c2) if len(a.shape) == 2 else (ba, r1, c2))
        self._sync_stream(dev)
        fn = self.lib.wifi_cmatmul_batch if dev else self.lib.wifi_cmatmul_host
        self._ck(fn(self.h, dt, a.ptr, r1, c1, b.ptr, r2, c2, _ptr(res), ba))
        return res

    def hermitian(self, M, mode=AS_WRITTEN):
        a = _Arg(M)
        ba, r, c = self._mat(a)
        res = a.empty_like((c, r) if len(a.shape) == 2 else (ba, c, r))
        self._sync_stream(a.device)
        fn = self.lib.wifi_chermitian_batch if a.device else self.lib.wifi_chermitian_host
        self._ck(fn(self.h, a.dt, mode, a.ptr, r, c, _ptr(res), ba))
        return res

    def addition(self, M1, M2, mode=AS_WRITTEN):
        a, b = _Arg(M1), _Arg(M2)
        dev, dt = _same(a, b)
        ba, r1, c1 = self._mat(a)
        bb, r2, c2 = self._mat(b)
        res = a.empty_like(a.shape)
        self._sync_stream(dev)
        fn = self.lib.wifi_cadd_batch if dev else self.lib.wifi_cadd_host
        self._ck(fn(self.h, dt, mode, a.ptr, r1, c1, b.ptr, r2, c2, _ptr(res), ba))
        return res

    def multiplyVxVeqM(self, M1, M2):
        a, b = _Arg(M1), _Arg(M2)
        dev, dt = _same(a, b)
        ba, r1, c1 = self._mat(a)
        bb, r2, c2 = self._mat(b)
        res = a.empty_like((r1, c2) if len(a.shape) == 2 else (ba, r1, c2))
        self._sync_stream(dev)
        fn = self.lib.wifi_couter_batch if dev else self.lib.wifi_couter_host
        self._ck(fn(self.h, dt, a.ptr, r1, c1, b.ptr, r2, c2, _ptr(res), ba))
        return res

    def identity(self, size, scalar, dtype=np.complex128, batch=1, like=None):
        if like is not None and _is_torch(like):
            import torch
            res = torch.empty((batch, size, size) if batch > 1 else (size, size), dtype=like.dtype, device=like.device)
            dt = F32 if like.dtype == torch.complex64 else F64
            self._sync_stream(True)
            self._ck(self.lib.wifi_cidentity_batch(self.h, dt, res.data_ptr(), size, float(scalar), batch))
            return res
        res = np.empty((batch, size, size) if batch > 1 else (size, size), dtype=dtype)
        dt = F32 if res.dtype == np.complex64 else F64
        self._ck(self.lib.wifi_cidentity_host(self.h, dt, res.ctypes.data, size, float(scalar), batch))
        return res

    def inverse(self, A):
        a = _Arg(A)
        ba, n, m = self._mat(a)
        if n != m:
            raise WifiError(_lib.ERR_INVALID, "inverse needs square matrices")
        Y = a.empty_like(a.shape)
        if a.device:
            import torch
            info = torch.zeros(ba, dtype=torch.int32, device=a.x.device)
            self._sync_stream(True)
            self._ck(self.lib.wifi_cinverse_batch(self.h, a.dt, a.ptr, n, _ptr(Y), ba, info.data_ptr()))
            if bool(info.any()):
                raise WifiError(_lib.ERR_SINGULAR, "singular matrix in batch")
            return Y
        info = np.zeros(ba, np.int32)
        self._ck(self.lib.wifi_cinverse_host(self.h, a.dt, a.ptr, n, _ptr(Y), ba, info.ctypes.data))
        return Y

    # ---- synthetic frames / statistics (device only) ----
    def synth_frames(self, n_frames, dtype="f32", seed=0x80211, first_frame=0, per_frame_sigma=False,
                     want=("tx_pre", "rx_pre", "tx_symb", "rx_symb", "H_true", "sigma2")):
        import torch
        dev = torch.device("cuda", self.device)
        cd, rd, dt = (torch.complex64, torch.float32, F32) if dtype in ("f32", F32) else (torch.complex128, torch.float64, F64)
        shapes = {"tx_pre": (n_frames, NSC), "rx_pre": (n_frames, NSC), "tx_symb": (n_frames, NBLK, NSC),
                  "rx_symb": (n_frames, NBLK, NSC), "H_true": (n_frames, NSC)}
        out = {k: torch.empty(shapes[k], dtype=cd, device=dev) for k in want if k in shapes}
        if "sigma2" in want:
            out["sigma2"] = torch.empty((n_frames,), dtype=rd, device=dev)
        self._sync_stream(True)
        self._ck(self.lib.wifi_synth_frames(self.h, dt, seed, first_frame, n_frames, int(per_frame_sigma),
                                            *(_ptr(out.get(k)) for k in ("tx_pre", "rx_pre", "tx_symb", "rx_symb", "H_true", "sigma2"))))
        return out

    def synth_covariance(self):
        import torch
        R = torch.empty((NSC, NSC), dtype=torch.complex128, device=torch.device("cuda", self.device))
        self._sync_stream(True)
        self._ck(self.lib.wifi_synth_covariance(self.h, R.data_ptr()))
        return R

    def error_stats(self, H, H_ref, stats=None):
        """Accumulates [sum|H-Href|^2, sum|Href|^2, count, max|H-Href|] into a 4-double CUDA tensor."""
        import torch
        a, b = _Arg(H), _Arg(H_ref)
        dev, dt = _same(a, b)
        if stats is None:
            stats = torch.zeros(4, dtype=torch.float64, device=a.x.device)
        self._sync_stream(True)
        self._ck(self.lib.wifi_error_stats(self.h, dt, a.ptr, b.ptr, a.size, stats.data_ptr()))
        return stats

    def measure_peaks(self):
        """Measured ceilings of this GPU: FP32/FP64 FMA and FP64 DMMA in TFLOP/s, streaming copy in GB/s."""
        out = {}
        for which, name in enumerate(("fp32_fma_tflops", "fp64_fma_tflops", "fp64_dmma_tflops", "copy_gbs")):
            v = C.c_double()
            self._ck(self.lib.wifi_measure_peak(self.h, which, C.byref(v)))
            out[name] = v.value
        return out


_default = None


def default_context():
    """Process-wide context on $WIFI_B200_DEVICE (default 0), created on first use."""
    global _default
    if _default is None:
        import os
        _default = WifiContext(int(os.environ.get("WIFI_B200_DEVICE", "0")))
    return _default


# ---- the reference's own names (main.c:4-8, WiFi_Equalization.m, utils.h:38-60) ----
def WiFi_channel_estimation_LT_LS(tx_pre, rx_pre):
    return default_context().lt_ls(tx_pre, rx_pre)


def WiFi_channel_estimation_PS_Linear(tx_symbols, rx_symbols):
    H = default_context().ps_linear(tx_symbols, rx_symbols)
    return H.reshape(np.shape(tx_symbols)) if np.ndim(tx_symbols) == 1 else H


def WiFi_channel_estimation_PS_Cubic(tx_symbols, rx_symbols):
    H = default_context().ps_cubic(tx_symbols, rx_symbols)
    return H.reshape(np.shape(tx_symbols)) if np.ndim(tx_symbols) == 1 else H


def WiFi_channel_estimation_PS_Sinc(tx_symbols, rx_symbols):
    H = default_context().ps_sinc(tx_symbols, rx_symbols)
    return H.reshape(np.shape(tx_symbols)) if np.ndim(tx_symbols) == 1 else H


def WiFi_channel_estimation_PS_MMSE(tx_symbols, rx_symbols, F, ow2, H_EST_LS):
    """main.c:148 argument list.  F (the 53-point DFT matrix of main.c:22-26) is accepted for signature
    compatibility; R = F (F^-1 H_ls)(F^-1 H_ls)^H F^H = H_ls H_ls^H does not depend on it."""
    del F
    return default_context().mmse_cconv(tx_symbols, rx_symbols, ow2, H_EST_LS)


def WiFi_Equalization(rx, H_EST_LT, H_EST_PS):
    return default_context().equalize(rx, H_EST_LT, H_EST_PS)


def hermitian(M):
    return default_context().hermitian(M, AS_WRITTEN)


def multiply(M1, M2):
    return default_context().multiply(M1, M2)


def multiplyVxVeqM(M1, M2):
    return default_context().multiplyVxVeqM(M1, M2)


def identity(size, scalar):
    return default_context().identity(size, scalar)


def addition(M1, M2):
    return default_context().addition(M1, M2, AS_WRITTEN)


def inverse(A):
    return default_context().inverse(A)
