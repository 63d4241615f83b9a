"""CPU check of the wide CUDA kernel family (mfft 1024 / 2048 / 4096, default geometry): the kernels'
own phase functions (veils_b200/csrc/wide_impl.cuh) are compiled for the host (tests/emu/wide_emu.cu)
and run tile by tile, thread by thread, against the oracle; the emulation also replays every
shared-memory access against the 32-bank model and the tests hold the design to its claim of
conflict-free passes.  The GPU tests repeat these cases on the device."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from golden_util import rel_err
from oracle.stft_oracle import OracleSTFT

import build_wide_emu

MODES = {"twosided": 0, "centered": 1, "onesided": 2, "onesided2X": 3}
PHASES = ["fwd pass1", "fwd pass2", "fwd pass3", "inv pass1", "inv pass2", "inv pass3", "inv ola", "inv out"]


@pytest.fixture(scope="module")
def emu():
    return C.CDLL(build_wide_emu.build())


def hann(m):
    return 0.5 * (1 - np.cos(2 * np.pi * np.arange(m) / m))


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def emu_stft(lib, o, tf, x, p0=None, p1=None, k_offset=0, spectrogram=False, ldp_align=2):
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float32)
    a, b = o.p_range(x.shape[1], p0, p1)
    P = b - a
    ldp = -(-P // ldp_align) * ldp_align
    out = np.full((x.shape[0], o.f_pts, ldp), np.nan, dtype=np.float32 if spectrogram else np.complex64)
    win = np.ascontiguousarray(o.win, dtype=np.float64)
    fac2x = o._fac2x if o.fft_mode == "onesided2X" else 1.0
    rc = lib.wide_emu_forward(C.c_int64(o.mfft), C.c_int(tf), dptr(win), C.c_void_p(x.ctypes.data), C.c_int64(x.shape[1]),
                              C.c_int64(x.shape[0]), C.c_int64(x.shape[1]), C.c_int64(a), C.c_int64(P), C.c_int64(k_offset),
                              C.c_int(MODES[o.fft_mode]), C.c_double(fac2x), C.c_int(int(spectrogram)),
                              C.c_void_p(out.ctypes.data), C.c_int64(ldp), C.c_int64(o.f_pts))
    assert rc == 0
    assert not np.isnan(out[:, :, :P]).any()
    return out[:, :, :P]


def emu_istft(lib, o, tf, S, k0=None, k1=None, chunk_tiles=3, ldp_align=2, land=1):
    S = np.asarray(S, dtype=np.complex64)
    B, F, P = S.shape
    ldp = -(-P // ldp_align) * ldp_align
    Sp = np.full((B, F, ldp), np.nan, dtype=np.complex64)
    Sp[:, :, :P] = S
    a, b, q0, q1 = o.istft_range(P, k0, k1)
    out = np.full((B, b - a), np.nan, dtype=np.float32)
    dual = np.ascontiguousarray(o.dual_win, dtype=np.float64)
    fac2x = o._fac2x if o.fft_mode == "onesided2X" else 1.0
    rc = lib.wide_emu_inverse(C.c_int64(o.mfft), C.c_int(tf), dptr(dual), C.c_void_p(Sp.ctypes.data), C.c_int64(P),
                              C.c_int64(ldp), C.c_int64(B), C.c_int64(F), C.c_int64(o.p_min), C.c_int64(a), C.c_int64(b),
                              C.c_int64(q0), C.c_int64(q1), C.c_int(MODES[o.fft_mode]), C.c_double(fac2x),
                              C.c_int(chunk_tiles), C.c_void_p(out.ctypes.data), C.c_int64(b - a), C.c_int(land))
    assert rc == 0
    assert not np.isnan(out).any()
    return out


def conflicts(lib):
    w = (C.c_longlong * 8)()
    i = (C.c_longlong * 8)()
    lib.wide_emu_conflicts(w, i)
    return {PHASES[k]: (w[k], i[k]) for k in range(8) if i[k]}


CASES = [(1024, 16), (2048, 16), (2048, 8), (4096, 8)]


@pytest.mark.parametrize("M,tf", CASES)
@pytest.mark.parametrize("mode", ["onesided", "twosided", "centered", "onesided2X"])
def test_forward_matches_oracle(emu, M, tf, mode):
    rng = np.random.default_rng(M + tf)
    o = OracleSTFT(hann(M), M // 4, 16000.0, fft_mode=mode)
    n = 5 * M + 37 * (M // 64) + 3
    x = rng.uniform(-1, 1, (2, n)).astype(np.float32)
    emu.wide_emu_reset()
    S = emu_stft(emu, o, tf, x)
    ref = np.stack([o.stft(r.astype(np.float64)) for r in x])
    assert S.shape == ref.shape
    assert rel_err(S, ref) <= 1e-5
    for name, (w, i) in conflicts(emu).items():
        assert w == i, f"{name}: {w} wavefronts for {i} conflict-free ones"


@pytest.mark.parametrize("M,tf", CASES)
def test_forward_ranges_and_spectrogram(emu, M, tf):
    rng = np.random.default_rng(7 * M + tf)
    o = OracleSTFT(hann(M) + 0.01, M // 4, 16000.0, fft_mode="onesided2X", scale_to="psd")
    n = 4 * M + 5
    x = rng.uniform(-1, 1, n).astype(np.float32)
    S = emu_stft(emu, o, tf, x, p0=1, p1=9, k_offset=3, ldp_align=1)       # odd k_offset, dense odd pitch
    ref = o.stft(x.astype(np.float64), 1, 9, 3)
    assert rel_err(S[0], ref) <= 1e-5
    Sx = emu_stft(emu, o, tf, x, spectrogram=True)
    refx = o.spectrogram(x.astype(np.float64))
    assert rel_err(Sx[0], refx) <= 2e-5


@pytest.mark.parametrize("M,tf", CASES)
@pytest.mark.parametrize("mode", ["onesided", "twosided", "centered", "onesided2X"])
def test_inverse_matches_oracle(emu, M, tf, mode):
    rng = np.random.default_rng(3 * M + tf)
    o = OracleSTFT(hann(M), M // 4, 16000.0, fft_mode=mode)
    n = (3 * tf + 5) * (M // 4) + 11
    x = rng.uniform(-1, 1, (2, n))
    S = np.stack([o.stft(r) for r in x])
    emu.wide_emu_reset()
    y = emu_istft(emu, o, tf, S)
    ref = np.stack([o.istft(s) for s in S.astype(np.complex64).astype(np.complex128)])
    assert y.shape == ref.shape
    assert rel_err(y, ref) <= 1e-5
    for name, (w, i) in conflicts(emu).items():
        assert w == i, f"{name}: {w} wavefronts for {i} conflict-free ones"
    # sub-range, dense odd pitch, other chunking: bit-identical where they overlap (same summation order)
    y2 = emu_istft(emu, o, tf, S, k0=M // 4 + 3, k1=n - 5, chunk_tiles=2, ldp_align=1)
    assert np.array_equal(y2, y[:, M // 4 + 3:n - 5])
    # the direct-load path (no tensor map) gives the same bits as the landed-tile path
    emu.wide_emu_reset()
    y3 = emu_istft(emu, o, tf, S, land=0)
    assert np.array_equal(y3, y)
    for name, (w, i) in conflicts(emu).items():
        assert w == i, f"{name}: {w} wavefronts for {i} conflict-free ones"
