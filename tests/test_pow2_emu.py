"""CPU check of the pow2 CUDA kernel family's index algebra: the kernels' own phase functions
(veils_b200/csrc/pow2_tile_impl.cuh) are compiled for the host (tests/emu) and run block by
block, thread by thread, against the oracle.  The GPU tests repeat these cases on the device."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from golden_util import rel_err
from oracle.stft_oracle import OracleSTFT

import build_emu

TOL = {1: 1e-12, 0: 1e-5}


@pytest.fixture(scope="module")
def emu():
    lib = C.CDLL(build_emu.build())
    lib.emu_last_error.restype = C.c_char_p
    return lib


def _i64(v):
    return None if v is None else C.byref(C.c_int64(v))


def emu_stft(lib, o, prec, x, p0=None, p1=None, k_offset=0, spectrogram=False, family=0):
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=rdt)
    a, b = o.p_range(x.shape[1], p0, p1)
    out = np.zeros((x.shape[0], o.f_pts, b - a), dtype=rdt if spectrogram else cdt)
    win = np.ascontiguousarray(o._win0, dtype=np.float64)
    rc = lib.emu_stft(win.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(win)), C.c_int64(o.hop),
                      C.c_double(o.fs), o.fft_mode.encode(), C.c_int64(o.mfft),
                      None if o.scaling is None else o.scaling.encode(), C.c_int64(o.phase_shift),
                      C.c_int(prec), C.c_void_p(x.ctypes.data), C.c_int64(x.shape[1]),
                      C.c_int64(x.shape[0]), _i64(p0), _i64(p1), C.c_int64(k_offset),
                      C.c_void_p(out.ctypes.data), C.c_int(int(spectrogram)), C.c_int(family))
    if rc == 2:
        return None
    assert rc == 0, (rc, lib.emu_last_error())
    return out


def emu_istft(lib, o, prec, S, k0=None, k1=None, family=0):
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    S = np.ascontiguousarray(S, dtype=cdt)
    a, b, _, _ = o.istft_range(S.shape[2], k0, k1)
    out = np.full((S.shape[0], b - a), np.nan, dtype=rdt)
    win = np.ascontiguousarray(o._win0, dtype=np.float64)
    rc = lib.emu_istft(win.ctypes.data_as(C.POINTER(C.c_double)), C.c_int64(len(win)), C.c_int64(o.hop),
                       C.c_double(o.fs), o.fft_mode.encode(), C.c_int64(o.mfft),
                       None if o.scaling is None else o.scaling.encode(), C.c_int64(o.phase_shift),
                       C.c_int(prec), C.c_void_p(S.ctypes.data), C.c_int64(S.shape[2]),
                       C.c_int64(S.shape[0]), _i64(k0), _i64(k1), C.c_void_p(out.ctypes.data),
                       C.c_int(family))
    if rc == 2:
        return None                          # not covered by the pow2 family -> generic kernels
    assert rc == 0, (rc, lib.emu_last_error())
    return out


def mk(win, hop, **kw):
    o = OracleSTFT(win, hop, 16000.0, **kw)
    o._win0 = np.asarray(win, float)        # unscaled window (the emu applies scale_to itself)
    return o


CASES = [  # (mfft, m_num, hop, mode, phase_shift, scale)
    (32, 32, 8, "onesided", 0, None), (64, 64, 16, "twosided", 0, None),
    (128, 128, 32, "centered", 0, None), (256, 256, 64, "onesided2X", 0, "psd"),
    (512, 512, 128, "onesided", 0, None), (512, 400, 160, "onesided", 0, None),
    (512, 511, 97, "onesided", 3, None), (1024, 1024, 256, "twosided", 0, None),
    (2048, 2048, 512, "onesided", 0, None), (4096, 4096, 1024, "onesided", 0, None),
    (8192, 8192, 2048, "onesided2X", 0, "magnitude"), (256, 100, 25, "centered", -7, None),
    (64, 33, 11, "onesided", 5, None), (64, 64, 2, "onesided", 0, None), (32, 20, 3, "twosided", 0, None),
]


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("case", CASES, ids=[f"M{c[0]}_N{c[1]}_H{c[2]}_{c[3]}" for c in CASES])
def test_emulated_kernels_match_oracle(emu, case, prec):
    M, N, H, mode, ps, scale = case
    rng = np.random.default_rng(M + N + H)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(N) / N)) + 0.02
    o = mk(win, H, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale)
    n = 3 * N + 2 * H + 13
    x = rng.uniform(-1, 1, (2, n))
    if prec == 0:
        x = x.astype(np.float32).astype(np.float64)
    Sref = np.stack([o.stft(r) for r in x])
    S = emu_stft(emu, o, prec, x)
    assert rel_err(S, Sref) <= TOL[prec]
    if M <= 512:
        sp = emu_stft(emu, o, prec, x, spectrogram=True)
        assert rel_err(sp, np.abs(Sref) ** 2) <= 4 * TOL[prec]
        a, b = o.p_min + 1, o.p_min + 4
        S2 = emu_stft(emu, o, prec, x, p0=a, p1=b, k_offset=3)
        assert rel_err(S2, np.stack([o.stft(r, a, b, 3) for r in x])) <= TOL[prec]
    yref = np.stack([o.istft(s) for s in Sref])
    y = emu_istft(emu, o, prec, Sref)
    if y is None:
        assert (M, prec) == (8192, 1)        # f64 tile of 2 slices < overlap halo of 3
        return
    assert not np.isnan(y).any()
    assert rel_err(y, yref) <= TOL[prec]
    if M <= 512:
        k0, k1 = H + 3, 2 * N + 5
        y2 = emu_istft(emu, o, prec, Sref, k0, k1)
        assert rel_err(y2, np.stack([o.istft(s, k0, k1) for s in Sref])) <= TOL[prec]


@pytest.mark.parametrize("case", CASES, ids=[f"M{c[0]}_N{c[1]}_H{c[2]}_{c[3]}" for c in CASES])
def test_emulated_packed_kernels_match_oracle(emu, case):
    """Packed f32 family (two slices per lane, FADD2/FFMA2): same cases, f32 tolerance."""
    M, N, H, mode, ps, scale = case
    rng = np.random.default_rng(M + N + H + 1)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(N) / N)) + 0.02
    o = mk(win, H, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale)
    n = 3 * N + 2 * H + 13
    x = rng.uniform(-1, 1, (2, n)).astype(np.float32).astype(np.float64)
    Sref = np.stack([o.stft(r) for r in x])
    S = emu_stft(emu, o, 0, x, family=1)
    fwd_ok = ps % 2 == 0 and H % 2 == 0 and N % 2 == 0 and H >= 4
    assert (S is not None) == fwd_ok
    if S is not None:
        assert rel_err(S, Sref) <= 1e-5
        sp = emu_stft(emu, o, 0, x, spectrogram=True, family=1)
        assert rel_err(sp, np.abs(Sref) ** 2) <= 4e-5
        a, b = o.p_min + 1, o.p_min + 4
        S2 = emu_stft(emu, o, 0, x, p0=a, p1=b, k_offset=2, family=1)
        assert rel_err(S2, np.stack([o.stft(r, a, b, 2) for r in x])) <= 1e-5
        # odd row pitch / odd slice counts are covered by P = S.shape[2] being odd or even here
    yref = np.stack([o.istft(s) for s in Sref])
    y = emu_istft(emu, o, 0, Sref, family=1)
    assert y is not None and not np.isnan(y).any()
    assert rel_err(y, yref) <= 1e-5
    if M <= 512:
        k0, k1 = H + 3, 2 * N + 5
        y2 = emu_istft(emu, o, 0, Sref, k0, k1, family=1)
        assert rel_err(y2, np.stack([o.istft(s, k0, k1) for s in Sref])) <= 1e-5


@pytest.mark.parametrize("M,N,H,n", [(64, 64, 16, 5000), (512, 512, 128, 30000), (256, 200, 50, 9000),
                                     (128, 128, 64, 6000), (64, 64, 64, 3000), (1024, 1024, 256, 40000)])
def test_emulated_packed_inverse_streaming_overlap_add(emu, M, N, H, n):
    """Long signals: several chunks of several tiles each -> the carried overlap tails, the chunk
    starts (halo re-computation) and clipped (k0, k1) ranges of the packed inverse."""
    rng = np.random.default_rng(M + H)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(N) / N)) + 0.02
    o = mk(win, H, fft_mode="onesided", mfft=M)
    x = rng.uniform(-1, 1, (2, n)).astype(np.float32).astype(np.float64)
    Sref = np.stack([o.stft(r) for r in x])
    y = emu_istft(emu, o, 0, Sref, family=1)
    assert y is not None and not np.isnan(y).any()
    assert rel_err(y, np.stack([o.istft(s) for s in Sref])) <= 1e-5
    assert np.max(np.abs(y[:, :n] - x)) <= 2e-5
    # odd k0: row-based overlap-add; even k0: the accumulator path (75 % overlap cases) with clipping
    for k0, k1 in ((3 * H + 5, n - 2 * H - 7), (3 * H + 6, n - 2 * H - 7), (2 * H, 5 * N)):
        y2 = emu_istft(emu, o, 0, Sref, k0, k1, family=1)
        assert not np.isnan(y2).any()
        assert rel_err(y2, np.stack([o.istft(s, k0, k1) for s in Sref])) <= 1e-5
    if N == 4 * H and M == N:
        # the two overlap-add paths agree bit for bit (same summation order, products rounded alone)
        ya = emu_istft(emu, o, 0, Sref, 2 * H, n - 3, family=1)
        yb = emu_istft(emu, o, 0, Sref, 2 * H - 1, n - 3, family=1)
        assert np.array_equal(ya, yb[:, 1:])
