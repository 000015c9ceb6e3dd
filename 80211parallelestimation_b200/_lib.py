"""ctypes binding of libwifi_b200.so (include/wifi_b200.h).  Fails loudly: no library, no product."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libwifi_b200.so")

F32, F64 = 0, 1
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_SINGULAR, ERR_NO_DEVICE, ERR_STATE = range(7)
PS_LINEAR, PS_CUBIC, PS_SINC, PS_MATLAB = 1, 2, 4, 8
SOLVE_PIVOT, SOLVE_HPD, SOLVE_REFINE = 0, 1, 2
SOLVE_WIDE = SOLVE_REFINE
SOLVE_FAST32 = 4
AS_WRITTEN, INTENDED = 0, 1

_vp, _i, _i64, _d, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint64

class RxChainOut(C.Structure):
    """wifi_rx_chain_out (include/wifi_b200.h): the optional output planes of the fused receiver chain."""
    _fields_ = [(k, _vp) for k in ("H_lt", "H_linear", "H_cubic", "H_sinc", "H_mmse_cconv", "H_ls0", "H_mmse_shared", "eq", "rx_symb", "ow2")]


# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/wifi_b200.h line by line
SIGNATURES = {
    "wifi_create": [_i, C.POINTER(_vp)],
    "wifi_destroy": [_vp],
    "wifi_set_stream": [_vp, _vp],
    "wifi_synchronize": [_vp],
    "wifi_last_error": [_vp],
    "wifi_version": [],
    "wifi_launch_count": [_vp],
    "wifi_enable_kernel_timing": [_vp, _i],
    "wifi_last_kernel_ms": [_vp, C.POINTER(C.c_float)],
    "wifi_lt_ls_batch": [_vp, _i, _vp, _vp, _vp, _i64],
    "wifi_ps_batch": [_vp, _i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _i64],
    "wifi_equalize_batch": [_vp, _i, _vp, _vp, _vp, _vp, _i64],
    "wifi_estimate_all_batch": [_vp, _i, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_frontend_batch": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_rx_chain_batch": [_vp, _i, _vp, _vp, _vp, _vp, C.POINTER(RxChainOut), _i64],
    "wifi_mmse_filter_form": [_vp, _vp, _vp, _vp],
    "wifi_mmse_filter_set": [_vp, _vp],
    "wifi_mmse_shared_apply_batch": [_vp, _i, _vp, _vp, _i64],
    "wifi_mmse_shared_batch": [_vp, _i, _vp, _vp, _i64, _vp, _i64],
    "wifi_mmse_filter_fold_tx": [_vp, _vp],
    "wifi_mmse_shared_rx_batch": [_vp, _i, _vp, _i64, _vp, _i64],
    "wifi_mmse_perframe_batch": [_vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i],
    "wifi_mmse_eig_prepare": [_vp, _vp, _vp],
    "wifi_mmse_perframe_eig_batch": [_vp, _i, _vp, _vp, _i64, _vp, _vp, _i64],
    "wifi_mmse_lowrank_prepare": [_vp, _vp, _vp],
    "wifi_mmse_perframe_lowrank_batch": [_vp, _i, _vp, _vp, _i64, _vp, _vp, _i64],
    "wifi_mmse_cconv_batch": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_mmse_matlab_batch": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_cmatmul_batch": [_vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_chermitian_batch": [_vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_cadd_batch": [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_couter_batch": [_vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_cidentity_batch": [_vp, _i, _vp, _i, _d, _i64],
    "wifi_cinverse_batch": [_vp, _i, _vp, _i, _vp, _i64, _vp],
    "wifi_synth_frames": [_vp, _i, _u64, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _vp],
    "wifi_synth_covariance": [_vp, _vp],
    "wifi_error_stats": [_vp, _i, _vp, _vp, _i64, _vp],
    "wifi_measure_peak": [_vp, _i, C.POINTER(_d)],
    "wifi_lt_ls_host": [_vp, _i, _vp, _vp, _vp, _i64],
    "wifi_ps_host": [_vp, _i, _i, _vp, _vp, _i64, _vp, _vp, _vp, _i64],
    "wifi_equalize_host": [_vp, _i, _vp, _vp, _vp, _vp, _i64],
    "wifi_estimate_all_host": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_frontend_host": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_rx_chain_host": [_vp, _i, _vp, _vp, _vp, _vp, C.POINTER(RxChainOut), _i64],
    "wifi_mmse_filter_form_host": [_vp, _vp, _vp, _vp],
    "wifi_mmse_shared_host": [_vp, _i, _vp, _vp, _i64, _vp, _i64],
    "wifi_mmse_filter_fold_tx_host": [_vp, _vp],
    "wifi_mmse_shared_rx_host": [_vp, _i, _vp, _i64, _vp, _i64],
    "wifi_pcie_probe": [_vp, _vp, _vp, C.c_size_t, C.c_size_t, C.POINTER(_d)],
    "wifi_mmse_perframe_host": [_vp, _i, _vp, _vp, _vp, _i64, _vp, _vp, _i64, _i],
    "wifi_mmse_perframe_eig_host": [_vp, _i, _vp, _vp, _i64, _vp, _vp, _i64],
    "wifi_mmse_perframe_lowrank_host": [_vp, _i, _vp, _vp, _i64, _vp, _vp, _i64],
    "wifi_mmse_cconv_host": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_mmse_matlab_host": [_vp, _i, _vp, _vp, _vp, _vp, _vp, _i64],
    "wifi_cmatmul_host": [_vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_chermitian_host": [_vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_cadd_host": [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_couter_host": [_vp, _i, _vp, _i, _i, _vp, _i, _i, _vp, _i64],
    "wifi_cidentity_host": [_vp, _i, _vp, _i, _d, _i64],
    "wifi_cinverse_host": [_vp, _i, _vp, _i, _vp, _i64, _vp],
    "wifi_set_host_chunk_bytes": [_vp, C.c_size_t],
    "wifi_host_alloc": [C.POINTER(_vp), C.c_size_t],
    "wifi_host_free": [_vp],
    "wifi_default_ctx": [],
}
_RESTYPES = {"wifi_last_error": C.c_char_p, "wifi_version": C.c_char_p, "wifi_launch_count": _i64, "wifi_default_ctx": _vp}

_lib = None


def load():
    """Load the CUDA library.  Raises if it has not been built: there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libwifi_b200.so is missing (%s): run `python 80211parallelestimation_b200/build.py`; "
            "this package has no CPU or PyTorch fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library disagree
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, _i)
    _lib = lib
    return lib
