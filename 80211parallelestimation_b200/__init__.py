"""b200-native 802.11 channel-estimation hot path (LT_LS, PS_Linear/Cubic/Sinc, PS_MMSE, equalizer and the
complex 53x53 matrix routines) behind the reference's own function names.

The package name starts with a digit, so import it with
    wifi = importlib.import_module("80211parallelestimation_b200")
All compute happens in hand-written sm_100a CUDA kernels in libwifi_b200.so (C-ABI: include/wifi_b200.h);
PyTorch is only used for device memory / streams by callers that pass CUDA tensors.
"""
from . import _lib
from ._lib import (F32, F64, PS_LINEAR, PS_CUBIC, PS_SINC, PS_MATLAB, SOLVE_PIVOT, SOLVE_HPD, SOLVE_REFINE, SOLVE_WIDE, SOLVE_FAST32, AS_WRITTEN, INTENDED)
from .api import (WifiContext, WifiError, default_context,
                  WiFi_channel_estimation_LT_LS, WiFi_channel_estimation_PS_Linear, WiFi_channel_estimation_PS_Cubic,
                  WiFi_channel_estimation_PS_Sinc, WiFi_channel_estimation_PS_MMSE, WiFi_Equalization,
                  hermitian, multiply, multiplyVxVeqM, identity, addition, inverse)
from .shard import shard_range, ShardedEstimator

NSC, NBLK, DC, PILOTS = 53, 15, 26, (5, 19, 33, 47)
