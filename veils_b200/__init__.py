"""veils_b200 -- B200-native (sm_100a) STFT / iSTFT / spectrogram engine behind the API of
ultrasaw/veils' `StandaloneSTFT`.  The product is libveils_cuda.so (CUDA kernels + C ABI,
include/veils_cuda.h); this package is the thin host mirror used by the tests and the bench."""
from ._lib import LIB_PATH, lib  # noqa: F401
from .stft import StandaloneSTFT, VeilsError  # noqa: F401

__all__ = ["StandaloneSTFT", "VeilsError", "lib", "LIB_PATH"]
