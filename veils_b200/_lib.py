"""ctypes binding of libveils_cuda.so (include/veils_cuda.h).  There is no fallback: if the
library is missing this module raises, and every compute call needs a CUDA device."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libveils_cuda.so")

OK, ERR_INVALID, ERR_NOT_INVERTIBLE, ERR_CUDA, ERR_UNSUPPORTED, ERR_NCCL = range(6)
F32, F64 = 0, 1


class PlanDesc(C.Structure):
    _fields_ = [("win", C.POINTER(C.c_double)), ("m_num", C.c_int64), ("hop", C.c_int64),
                ("fs", C.c_double), ("fft_mode", C.c_char_p), ("mfft", C.c_int64),
                ("dual_win", C.POINTER(C.c_double)), ("scale_to", C.c_char_p),
                ("phase_shift", C.c_int64), ("precision", C.c_int32), ("device", C.c_int32)]


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)
_DP = C.POINTER(C.c_double)

# name -> (restype, argtypes); mirrors include/veils_cuda.h one to one
PROTOTYPES = {
    "veils_abi_version": (C.c_int32, []),
    "veils_plan_create": (C.c_int32, [C.POINTER(PlanDesc), C.POINTER(_P)]),
    "veils_plan_destroy": (None, [_P]),
    "veils_last_error": (C.c_char_p, []),
    "veils_m_num": (C.c_int64, [_P]), "veils_m_num_mid": (C.c_int64, [_P]),
    "veils_hop": (C.c_int64, [_P]), "veils_mfft": (C.c_int64, [_P]),
    "veils_f_pts": (C.c_int64, [_P]), "veils_phase_shift": (C.c_int64, [_P]),
    "veils_fs": (C.c_double, [_P]), "veils_delta_t": (C.c_double, [_P]),
    "veils_delta_f": (C.c_double, [_P]), "veils_fft_mode": (C.c_int32, [_P]),
    "veils_onesided_fft": (C.c_int32, [_P]), "veils_scaling": (C.c_int32, [_P]),
    "veils_precision": (C.c_int32, [_P]), "veils_device": (C.c_int32, [_P]),
    "veils_win": (C.c_int32, [_P, _DP]), "veils_dual_win": (C.c_int32, [_P, _DP]),
    "veils_invertible": (C.c_int32, [_P]),
    "veils_p_min": (C.c_int64, [_P]), "veils_k_min": (C.c_int64, [_P]),
    "veils_p_max": (C.c_int32, [_P, C.c_int64, _I64P]),
    "veils_k_max": (C.c_int32, [_P, C.c_int64, _I64P]),
    "veils_p_num": (C.c_int32, [_P, C.c_int64, _I64P]),
    "veils_p_range": (C.c_int32, [_P, C.c_int64, _I64P, _I64P, _I64P, _I64P]),
    "veils_istft_range": (C.c_int32, [_P, C.c_int64, _I64P, _I64P, _I64P, _I64P]),
    "veils_t": (C.c_int32, [_P, C.c_int64, _I64P, _I64P, C.c_int64, _DP]),
    "veils_f": (C.c_int32, [_P, _DP]),
    "veils_lower_border_end": (C.c_int32, [_P, _I64P, _I64P]),
    "veils_upper_border_begin": (C.c_int32, [_P, C.c_int64, _I64P, _I64P]),
    "veils_extent": (C.c_int32, [_P, C.c_int64, C.c_int32, C.c_int32, _DP]),
    "veils_nearest_k_p": (C.c_int64, [_P, C.c_int64, C.c_int32]),
    "veils_debug_ifft_func": (C.c_int32, [_P, _DP, _DP]),
    "veils_stft_dev": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P,
                                    C.c_int64, _P, C.c_int64, _P]),
    "veils_spectrogram_dev": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P,
                                           C.c_int64, _P, C.c_int64, _P]),
    "veils_istft_dev": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P, _P,
                                     C.c_int64, _P]),
    "veils_pad_extent": (C.c_int32, [_P, C.c_int64, _I64P, _I64P, C.c_int64, _I64P, _I64P]),
    "veils_stft_dev_padded": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P, C.c_int64,
                                           C.c_char_p, _P, C.c_int64, _P, C.c_int64, _P, C.c_int32]),
    "veils_stft_detrend_dev": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P, C.c_int64,
                                            C.c_char_p, C.c_char_p, _P, C.c_int64, _P, _P, C.c_int64, _P]),
    "veils_cross_multiply_dev": (C.c_int32, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    "veils_complex_op_dev": (C.c_int32, [_P, C.c_int32, _P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_int64, _P]),
    "veils_stft_dev_checked": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P,
                                            C.c_int64, _P, C.c_int64, _P, C.c_int32, _P]),
    "veils_stft_host": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P,
                                     C.c_int64, _P]),
    "veils_spectrogram_host": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _I64P, _I64P,
                                            C.c_int64, _P]),
    "veils_istft_host": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, _I64P, _I64P, _P]),
    "veils_device_count": (C.c_int32, [C.POINTER(C.c_int32)]),
    "veils_dev_alloc": (C.c_int32, [C.c_int32, C.c_uint64, C.POINTER(_P)]),
    "veils_dev_free": (C.c_int32, [C.c_int32, _P]),
    "veils_dev_upload": (C.c_int32, [_P, _P, C.c_uint64, _P]),
    "veils_dev_download": (C.c_int32, [_P, _P, C.c_uint64, _P]),
    "veils_host_alloc_pinned": (C.c_int32, [C.c_uint64, C.POINTER(_P)]),
    "veils_host_free_pinned": (C.c_int32, [_P]),
    "veils_stream_create": (C.c_int32, [C.c_int32, C.POINTER(_P)]),
    "veils_stream_destroy": (C.c_int32, [_P]),
    "veils_stream_sync": (C.c_int32, [_P]),
    "veils_debug_launch_cost_us": (C.c_int32, [_P, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64, _P, C.c_int32, _DP]),
    "veils_kernel_launch_count": (C.c_uint64, []),
    "veils_plan_kernel_name": (C.c_char_p, [_P, C.c_int32]),
}

_lib = None


def lib():
    """Load libveils_cuda.so (once).  Raises if it has not been built -- no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m veils_b200.build` "
                "(veils_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error():
    return lib().veils_last_error().decode("utf-8", "replace")


def opt_i64(v):
    """Option<i64> -> nullable pointer."""
    return None if v is None else C.byref(C.c_int64(int(v)))
