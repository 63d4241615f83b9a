"""GPU parity of the wide kernel family (mfft 1024 / 2048 / 4096, m_num = mfft = 4 hop) through the
C ABI: every layout against the oracle, sub-ranges, odd offsets, dense (unaligned) row pitches, the
row tail of a BASELINE c3 row, spectrogram, bit-identical repeat runs and shard-wise inverse."""
import os

import numpy as np
import pytest

from golden_util import rel_err
from oracle.stft_oracle import OracleSTFT
from veils_b200 import StandaloneSTFT, sharding

pytestmark = pytest.mark.gpu


def hann(m):
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(m) / m))


@pytest.mark.parametrize("M", [1024, 2048, 4096])
@pytest.mark.parametrize("mode", ["onesided", "twosided", "centered", "onesided2X"])
def test_wide_matches_oracle(M, mode):
    rng = np.random.default_rng(M + len(mode))
    st = StandaloneSTFT(hann(M), M // 4, 16000.0, fft_mode=mode, precision="f32")
    assert st.kernel_name() == "pow2_wide"
    assert st.kernel_name(True) == ("pow2_wide" if mode.startswith("onesided") else "pow2_packed")
    o = OracleSTFT(hann(M), M // 4, 16000.0, fft_mode=mode)
    n = 37 * (M // 4) + 123
    x = rng.uniform(-1, 1, (3, n)).astype(np.float32)
    S = st.stft(x)                                                  # host arrays: dense pitch, staged
    ref = np.stack([o.stft(r.astype(np.float64)) for r in x])
    assert S.shape == ref.shape
    assert rel_err(S, ref) <= 1e-5
    y = st.istft(S)
    refy = np.stack([o.istft(s.astype(np.complex128)) for s in S])
    assert y.shape == refy.shape
    assert rel_err(y, refy) <= 1e-5
    assert np.max(np.abs(y[:, :n] - x)) <= 4e-6                     # perfect reconstruction
    # sub-ranges: odd k_offset (unaligned staging), p0 / p1, k0 / k1
    S2 = st.stft(x[1], 2, 29, k_offset=5)
    assert rel_err(S2, o.stft(x[1].astype(np.float64), 2, 29, 5)) <= 1e-5
    y2 = st.istft(S[1], M // 4 + 7, n - 9)
    assert np.array_equal(y2, y[1, M // 4 + 7:n - 9])               # same summation order -> same bits


@pytest.mark.parametrize("M", [1024, 4096])
def test_wide_device_tensors_spectrogram_and_determinism(M):
    import torch
    g = torch.Generator(device="cuda").manual_seed(M)
    st = StandaloneSTFT(hann(M) + 0.01, M // 4, 16000.0, fft_mode="onesided2X", scale_to="psd", precision="f32")
    o = OracleSTFT(hann(M) + 0.01, M // 4, 16000.0, fft_mode="onesided2X", scale_to="psd")
    n = 53 * (M // 4) + 17
    x = torch.rand((5, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    S = st.stft(x)
    assert torch.equal(S, st.stft(x))
    xr = x.double().cpu().numpy()
    assert rel_err(S[3].cpu().numpy(), o.stft(xr[3])) <= 1e-5
    sp = st.spectrogram(x)
    assert rel_err(sp[4].cpu().numpy(), o.spectrogram(xr[4])) <= 2e-5
    y = st.istft(S)
    assert torch.equal(y, st.istft(S))
    assert float((y[:, :n] - x).abs().max()) <= 1e-5 * float(x.abs().max()) * 4


def test_wide_c3_row_tail_and_interior():
    """BASELINE c3 shape (Hann 4096 / hop 1024, 30 s @ 48 kHz): the LAST slices of a row (clipped tile,
    zero padding behind the signal) and an interior range against the oracle, on two rows of a batch."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(33)
    n = 1440000
    x = torch.rand((6, n), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    st = StandaloneSTFT(hann(4096), 1024, 48000.0, precision="f32")
    o = OracleSTFT(hann(4096), 1024, 48000.0)
    S = st.stft(x)
    P = S.shape[2]
    assert S.shape == (6, 2049, 1410)
    for row in (0, 5):
        xr = x[row].double().cpu().numpy()
        # last 40 slices from the signal's tail
        p1 = o.p_max(n)
        p0 = p1 - 40
        tail0 = p0 * 1024 - 2048
        ref = o.stft(xr[tail0:], p0, p1, k_offset=-tail0)
        assert rel_err(S[row, :, P - 40:].cpu().numpy(), ref) <= 1e-5
        # interior
        pa = 700
        s_a, s_b = pa * 1024 - 2048, (pa + 24) * 1024 - 2048 + 4096
        ref = o.stft(xr[s_a:s_b], pa, pa + 25, k_offset=-s_a)
        assert rel_err(S[row, :, pa - st.p_min: pa - st.p_min + 25].cpu().numpy(), ref) <= 1e-5
    y = st.istft(S)
    assert float((y[:, :n] - x).abs().max()) <= 4e-6
    # the tail of the reconstruction against the oracle's istft of the last slices
    S_tail = S[5, :, P - 12:].cpu().numpy().astype(np.complex128)
    q_a = o.p_min + P - 12                                          # first slice of the tail block
    k_a = (q_a + 3) * 1024 - 2048                                    # samples that only these slices touch
    sub = OracleSTFT(hann(4096), 1024, 48000.0)
    shift = (q_a - o.p_min) * 1024
    ref_tail = sub.istft(S_tail, k_a - shift, None)
    got = y[5, k_a:].cpu().numpy()
    assert got.shape == ref_tail.shape
    assert rel_err(got, ref_tail) <= 1e-5


def test_wide_sharded_inverse_is_bit_identical():
    rng = np.random.default_rng(8)
    for M, mode in ((4096, "onesided"), (1024, "twosided")):
        st = StandaloneSTFT(hann(M), M // 4, 16000.0, fft_mode=mode, precision="f32")
        x = rng.uniform(-1, 1, 300 * (M // 4) + 5).astype(np.float32)
        S = st.stft(x)
        full = st.istft(S)
        P = S.shape[1]
        parts = []
        for r in range(3):
            s = sharding.inverse_shard(P, M, M // 4, st.p_min, r, 3)
            a = s.p_a - st.p_min - s.halo_cols
            parts.append(st.istft(np.ascontiguousarray(S[:, a:s.p_b - st.p_min]), s.k0_local, s.k1_local))
        assert np.array_equal(np.concatenate(parts), full)


@pytest.mark.parametrize("mode", ["twosided", "centered"])
def test_wide_inverse_two_sided_forced(monkeypatch, mode):
    """the wide inverse of the two-sided layouts (direct loads) is kept correct although the dispatcher
    prefers the packed family there"""
    monkeypatch.setenv("VEILS_WIDE_INV_ALL", "1")
    rng = np.random.default_rng(10)
    M = 1024
    st = StandaloneSTFT(hann(M), M // 4, 16000.0, fft_mode=mode, precision="f32")
    assert st.kernel_name(True) == "pow2_wide"
    o = OracleSTFT(hann(M), M // 4, 16000.0, fft_mode=mode)
    x = rng.uniform(-1, 1, (2, 61 * 256 + 9))
    S = np.stack([o.stft(r) for r in x]).astype(np.complex64)
    y = st.istft(S)
    assert rel_err(y, np.stack([o.istft(s.astype(np.complex128)) for s in S])) <= 1e-5


def test_wide_tf16_variant_for_2048(monkeypatch):
    monkeypatch.setenv("VEILS_WIDE_TF", "16")
    rng = np.random.default_rng(9)
    st = StandaloneSTFT(hann(2048), 512, 16000.0, precision="f32")
    o = OracleSTFT(hann(2048), 512, 16000.0)
    x = rng.uniform(-1, 1, (2, 41 * 512 + 3)).astype(np.float32)
    S = st.stft(x)
    assert rel_err(S, np.stack([o.stft(r.astype(np.float64)) for r in x])) <= 1e-5
    y = st.istft(S)
    assert np.max(np.abs(y[:, :x.shape[1]] - x)) <= 4e-6
