"""GPU parity of the mixed-radix kernel family (non-power-of-two mfft with prime factors <= 13) through the
C ABI, against the oracle: the reference's own 15-point golden geometry, the Hann 400 / hop 160 speech front
end, mfft > m_num, odd hops, phase shifts, all layouts, f32 and f64; lengths with a larger prime factor
still run the O(M^2) kernels."""
import numpy as np
import pytest

from golden_util import rel_err
from oracle.stft_oracle import OracleSTFT
from veils_b200 import StandaloneSTFT

pytestmark = pytest.mark.gpu

CASES = [(15, 8, 15, 0), (400, 160, 400, 0), (50, 10, 60, 3), (21, 4, 63, -5), (9, 3, 441, 2), (30, 6, 1000, 0),
         (600, 150, 600, 0), (33, 7, 33, 0)]


@pytest.mark.parametrize("m,hop,mfft,ps", CASES)
@pytest.mark.parametrize("mode", ["onesided", "twosided", "centered", "onesided2X"])
@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_mixed_radix_matches_oracle(m, hop, mfft, ps, mode, prec, tol):
    rng = np.random.default_rng(m * 11 + hop)
    win = 0.5 * (1 - np.cos(2 * np.pi * (np.arange(m) + 0.5) / m)) + 0.05
    o = OracleSTFT(win, hop, 1000.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
    st = StandaloneSTFT(win, hop, 1000.0, fft_mode=mode, mfft=mfft, phase_shift=ps, precision=prec)
    assert st.kernel_name() == "mixed_radix" and st.kernel_name(True) == "mixed_radix"
    n = 150 * hop + m + 5
    x = rng.uniform(-1, 1, (3, n)).astype(np.float32 if prec == "f32" else np.float64)
    S = st.stft(x)
    ref = np.stack([o.stft(r.astype(np.float64)) for r in x])
    assert S.shape == ref.shape and rel_err(S, ref) <= tol
    assert rel_err(st.stft(x[1], 2, 40, k_offset=3), o.stft(x[1].astype(np.float64), 2, 40, 3)) <= tol
    assert rel_err(st.spectrogram(x[2]), o.spectrogram(x[2].astype(np.float64))) <= 2 * tol
    y = st.istft(S)
    assert rel_err(y, np.stack([o.istft(s.astype(np.complex128)) for s in S])) <= tol
    assert np.max(np.abs(y[:, :n] - x)) <= 50 * tol
    y2 = st.istft(S[0], hop + 1, n - 3)
    assert np.array_equal(y2, y[0, hop + 1:n - 3])                  # same summation order -> same bits


def test_large_prime_factors_stay_on_the_generic_kernels():
    win = np.hanning(34)[1:-1] + 0.1
    for mfft in (34, 38, 62):                                       # 17, 19, 31
        st = StandaloneSTFT(win[:32], 8, 1.0, mfft=mfft, precision="f64")
        assert st.kernel_name() == "dft_generic"
        o = OracleSTFT(win[:32], 8, 1.0, mfft=mfft)
        x = np.random.default_rng(mfft).standard_normal(500)
        assert rel_err(st.stft(x), o.stft(x)) <= 1e-12


def test_device_tensors_and_determinism_400_160():
    import torch
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(400) / 400))
    st = StandaloneSTFT(win, 160, 16000.0, precision="f32")
    o = OracleSTFT(win, 160, 16000.0)
    x = torch.rand((64, 160000), device="cuda", dtype=torch.float32) * 2 - 1
    S = st.stft(x)
    assert torch.equal(S, st.stft(x))
    assert rel_err(S[63].cpu().numpy(), o.stft(x[63].double().cpu().numpy())) <= 2e-5
    y = st.istft(S)
    assert torch.equal(y, st.istft(S))
    assert float((y[:, :160000] - x).abs().max()) <= 1e-5


@pytest.mark.parametrize("prec,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_randomized_mixed_radix_geometry(prec, tol):
    """Randomized geometry for the mixed-radix family: smooth mfft (even and odd), window shorter than mfft, arbitrary
    hop, roll, k_offset, slice and sample sub-ranges, very short signals, all four layouts."""
    rng = np.random.default_rng(4242)
    smooth = [6, 10, 12, 15, 18, 20, 21, 24, 30, 36, 45, 48, 60, 63, 70, 72, 90, 96, 100, 120, 126, 150, 160, 200, 240,
              243, 250, 320, 360, 400, 441, 480, 486, 500, 600, 625, 640, 720, 800, 882, 960, 1000, 1200, 1215]
    seen = 0
    for it in range(36):
        M = int(rng.choice(smooth))
        N = int(M if rng.random() < 0.5 else rng.integers(max(2, M // 2), M + 1))
        H = int(rng.choice([max(1, N // 4), max(1, N // 2), max(1, N // 3), int(rng.integers(1, N + 1))]))
        mode = str(rng.choice(["onesided", "twosided", "centered", "onesided2X"]))
        ps = int(rng.integers(-M + 1, M)) if rng.random() < 0.4 else 0
        scale = rng.choice([None, None, "magnitude", "psd"])
        win = rng.uniform(0.2, 1.0, N)
        n = int(rng.integers(N, N + 60 * H + 50)) if rng.random() < 0.8 else int(N - N // 2 + rng.integers(0, 3))
        x = rng.uniform(-1, 1, (2, n)).astype(np.float32 if prec == "f32" else np.float64)
        o = OracleSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale)
        s = StandaloneSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale, precision=prec)
        if s.kernel_name() != "mixed_radix":
            continue
        seen += 1
        koff = int(rng.integers(-3, 4))
        xr = x.astype(np.float64)
        Sref = np.stack([o.stft(r, k_offset=koff) for r in xr])
        assert rel_err(s.stft(x, k_offset=koff), Sref) <= tol, (M, N, H, mode, ps, n)
        assert rel_err(s.spectrogram(x, k_offset=koff), np.abs(Sref) ** 2) <= 4 * tol, (M, N, H, mode, ps, n)
        if o.invertible and s.kernel_name(True) == "mixed_radix":
            S0 = np.stack([o.stft(r) for r in xr])
            S0 = S0.astype(np.complex64 if prec == "f32" else np.complex128)
            yref = np.stack([o.istft(q.astype(np.complex128)) for q in S0])
            assert rel_err(s.istft(S0), yref) <= tol, (M, N, H, mode, ps, n)
            if n > 2 * N:
                k0 = int(rng.integers(0, n // 3))
                k1 = int(rng.integers(k0 + N, n))
                yr2 = np.stack([o.istft(q.astype(np.complex128), k0, k1) for q in S0])
                assert rel_err(s.istft(S0, k0, k1), yr2) <= tol, (M, N, H, mode, ps, n, k0, k1)
    assert seen >= 20


def test_plans_sharing_a_kernel_with_different_shared_memory():
    """Two plans that run the SAME kernel instantiation with different shared-memory sizes, interleaved: the per-function
    shared-memory attribute must never be lowered under a plan that needs more (regression: 'invalid argument')."""
    rng = np.random.default_rng(99)
    plans = []
    for m, hop in ((600, 150), (400, 160), (320, 80)):
        win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(m) / m))
        plans.append((StandaloneSTFT(win, hop, 16000.0, precision="f32"), OracleSTFT(win, hop, 16000.0)))
    x = rng.uniform(-1, 1, (2, 20000)).astype(np.float32)
    for _ in range(2):
        for st, o in plans + plans[::-1]:
            S = st.stft(x)
            assert rel_err(S[1], o.stft(x[1].astype(np.float64))) <= 2e-5
            assert np.max(np.abs(st.istft(S)[:, :20000] - x)) <= 1e-4
