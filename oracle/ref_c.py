"""ctypes wrapper of oracle/stft_ref.c -- the C restatement of the reference's CPU path.
TEST INFRASTRUCTURE / CPU BASELINE ONLY (see the header of stft_ref.c)."""
import ctypes as C

import numpy as np

from . import build_c
from .stft_oracle import OracleSTFT

_MODE = {"twosided": 0, "centered": 1, "onesided": 2, "onesided2X": 3}
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_c.build())
        _lib.ref_num_threads.restype = C.c_int
    return _lib


def num_threads():
    return lib().ref_num_threads()


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class RefC:
    """Batched stft / istft with the per-slice structure of src/lib.rs, geometry from the oracle."""

    def __init__(self, oracle: OracleSTFT):
        self.o = oracle

    def stft(self, x, p0=None, p1=None, k_offset=0, threads=0):
        o = self.o
        x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        batch, n = x.shape
        p0, p1 = o.p_range(n, p0, p1)
        S = np.empty((batch, o.f_pts, p1 - p0), dtype=np.complex128)
        win = np.ascontiguousarray(o.win)
        rc = lib().ref_stft(_dp(x), C.c_int64(n), C.c_int64(batch), _dp(win), C.c_int64(o.m_num),
                            C.c_int64(o.hop), C.c_int64(o.mfft), C.c_int(_MODE[o.fft_mode]),
                            C.c_int64(o._p_s), C.c_double(o._fac2x), C.c_int64(p0), C.c_int64(p1),
                            C.c_int64(k_offset), C.c_void_p(S.ctypes.data), C.c_int(threads))
        assert rc == 0
        return S

    def istft(self, S, k0=None, k1=None, threads=0):
        o = self.o
        S = np.ascontiguousarray(S, dtype=np.complex128)
        if S.ndim == 2:
            S = S[None]
        batch, F, P = S.shape
        assert F == o.f_pts
        k0, k1, q0, q1 = o.istft_range(P, k0, k1)
        y = np.empty((batch, k1 - k0))
        dual = np.ascontiguousarray(o.dual_win)
        rc = lib().ref_istft(C.c_void_p(S.ctypes.data), C.c_int64(P), C.c_int64(batch), _dp(dual),
                             C.c_int64(o.m_num), C.c_int64(o.hop), C.c_int64(o.mfft),
                             C.c_int(_MODE[o.fft_mode]), C.c_int64(o._p_s), C.c_double(o._fac2x),
                             C.c_int64(o.p_min), C.c_int64(k0), C.c_int64(k1), C.c_int64(q0),
                             C.c_int64(q1), _dp(y), C.c_int(threads))
        assert rc == 0
        return y
