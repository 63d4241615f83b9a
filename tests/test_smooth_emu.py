"""CPU check of the mixed-radix ("smooth") CUDA kernel family: its small in-register DFTs (every length
2 .. 16), the factorisation, and the kernels' own phase functions (veils_b200/csrc/smooth_impl.cuh,
compiled for the host) tile by tile against the oracle -- non-power-of-two mfft, mfft > m_num, odd hops,
phase shifts, all layouts, f32 and f64.  The GPU tests repeat these cases on the device."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
from golden_util import rel_err
from oracle.stft_oracle import OracleSTFT

import build_smooth_emu

MODES = {"twosided": 0, "centered": 1, "onesided": 2, "onesided2X": 3}


@pytest.fixture(scope="module")
def emu():
    return C.CDLL(build_smooth_emu.build())


def dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("R", range(2, 17))
def test_small_dfts(emu, R):
    rng = np.random.default_rng(R)
    x = rng.standard_normal(R) + 1j * rng.standard_normal(R)
    for inverse in (0, 1):
        buf = np.ascontiguousarray(x.copy()).view(np.float64)
        assert emu.smooth_emu_dft(R, inverse, dptr(buf)) == 1
        ref = np.fft.ifft(x) * R if inverse else np.fft.fft(x)
        assert rel_err(buf.view(np.complex128), ref) <= 1e-14


def test_factorisation(emu):
    for M, ok in ((15, True), (400, True), (441, True), (1000, True), (17, False), (2 * 19, False), (30030, True), (60060, False),
                  (65536, True), (3 * 5 * 7 * 11 * 13, True), (13 ** 4, True), (13 ** 5, False)):
        npass, R = C.c_int(0), (C.c_int * 4)()
        got = emu.smooth_emu_factors(C.c_int64(M), C.byref(npass), R)
        assert bool(got) == ok, M
        if ok:
            assert int(np.prod([R[i] for i in range(4)])) == M and all(1 <= R[i] <= 16 for i in range(4))


def run_fwd(emu, o, prec, x, p0=None, p1=None, k_offset=0, spec=False):
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=rdt)
    a, b = o.p_range(x.shape[1], p0, p1)
    out = np.full((x.shape[0], o.f_pts, b - a), np.nan, dtype=rdt if spec else cdt)
    win = np.ascontiguousarray(o.win, dtype=np.float64)
    fac = o._fac2x if o.fft_mode == "onesided2X" else 1.0
    rc = emu.smooth_emu_forward(prec, C.c_void_p(x.ctypes.data), C.c_int64(x.shape[1]), C.c_int64(x.shape[0]), dptr(win),
                                C.c_int64(o.mfft), C.c_int64(o.m_num), C.c_int64(o.hop), C.c_int64(o.f_pts),
                                C.c_int64(o._p_s), C.c_int64(o.m_num_mid), C.c_int64(a), C.c_int64(b - a),
                                C.c_int64(k_offset), MODES[o.fft_mode], C.c_double(fac), int(spec),
                                C.c_void_p(out.ctypes.data), C.c_int64(b - a))
    assert rc == 0
    assert not np.isnan(out).any()
    return out


def run_inv(emu, o, prec, S, k0=None, k1=None):
    rdt, cdt = (np.float32, np.complex64) if prec == 0 else (np.float64, np.complex128)
    S = np.ascontiguousarray(S, dtype=cdt)
    B, F, P = S.shape
    a, b, q0, q1 = o.istft_range(P, k0, k1)
    out = np.full((B, b - a), np.nan, dtype=rdt)
    dual = np.ascontiguousarray(o.dual_win, dtype=np.float64)
    fac = o._fac2x if o.fft_mode == "onesided2X" else 1.0
    rc = emu.smooth_emu_inverse(prec, C.c_void_p(S.ctypes.data), C.c_int64(P), C.c_int64(P), C.c_int64(B), dptr(dual),
                                C.c_int64(o.mfft), C.c_int64(o.m_num), C.c_int64(o.hop), C.c_int64(F), C.c_int64(o._p_s),
                                C.c_int64(o.m_num_mid), C.c_int64(o.p_min), C.c_int64(a), C.c_int64(b), C.c_int64(q0),
                                C.c_int64(q1), MODES[o.fft_mode], C.c_double(fac), C.c_void_p(out.ctypes.data),
                                C.c_int64(b - a))
    assert rc == 0
    assert not np.isnan(out).any()
    return out


CASES = [  # m_num, hop, mfft, phase_shift
    (15, 8, 15, 0), (400, 160, 400, 0), (50, 10, 60, 3), (21, 4, 63, -5), (100, 25, 100, 0), (33, 7, 33, 0),
    (9, 3, 441, 2), (30, 6, 1000, 0), (12, 5, 26, 0),
    # even mfft = half-length transform + split: single-pass halves, odd halves, odd rolls, mfft = 2 (no half)
    (4, 2, 4, 0), (6, 2, 6, 1), (30, 10, 30, 0), (20, 5, 882, 7), (2, 1, 2, 0), (14, 3, 18, -3), (320, 160, 320, 0),
]


@pytest.mark.parametrize("m,hop,mfft,ps", CASES)
@pytest.mark.parametrize("mode", ["onesided", "twosided", "centered", "onesided2X"])
@pytest.mark.parametrize("prec,tol", [(1, 1e-12), (0, 2e-5)])
def test_smooth_matches_oracle(emu, m, hop, mfft, ps, mode, prec, tol):
    rng = np.random.default_rng(m * 7 + hop)
    win = 0.5 * (1 - np.cos(2 * np.pi * (np.arange(m) + 0.5) / m)) + 0.05
    o = OracleSTFT(win, hop, 1000.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
    n = 23 * hop + m + 5
    x = rng.uniform(-1, 1, (2, n))
    S = run_fwd(emu, o, prec, x)
    xr = x.astype(np.float32 if prec == 0 else np.float64).astype(np.float64)
    ref = np.stack([o.stft(r) for r in xr])
    assert S.shape == ref.shape
    assert rel_err(S, ref) <= tol
    S2 = run_fwd(emu, o, prec, x[0], 1, 9, k_offset=3)
    assert rel_err(S2[0], o.stft(xr[0], 1, 9, 3)) <= tol
    sp = run_fwd(emu, o, prec, x[1], spec=True)
    assert rel_err(sp[0], o.spectrogram(xr[1])) <= 2 * tol
    y = run_inv(emu, o, prec, ref)
    Sr = ref.astype(np.complex64 if prec == 0 else np.complex128).astype(np.complex128)
    assert rel_err(y, np.stack([o.istft(s) for s in Sr])) <= tol
    y2 = run_inv(emu, o, prec, ref, k0=hop + 1, k1=n - 3)
    assert rel_err(y2, np.stack([o.istft(s, hop + 1, n - 3) for s in Sr])) <= tol
