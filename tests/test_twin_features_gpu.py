"""GPU parity of the twin-only features of SURVEY 8(f3) that round 2 added: complex input and complex
ISTFT in the two-sided layouts (python/standalone_stft.py:349, 404, 424), scipy's phase_shift=None,
and the reference's NaN screen (src/lib.rs:696-698) on the device-pointer flavour.  The reference for
complex signals is live scipy.signal.ShortTimeFFT, which the twin restates bit for bit."""
import numpy as np
import pytest
from scipy.signal import ShortTimeFFT

from golden_util import rel_err
from veils_b200 import StandaloneSTFT
from veils_b200.stft import VeilsError

pytestmark = pytest.mark.gpu


def hann(m):
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * (np.arange(m) + 0.5) / m))


@pytest.mark.parametrize("mode", ["twosided", "centered"])
@pytest.mark.parametrize("m,hop,prec,tol", [(64, 16, "f64", 1e-12), (1024, 256, "f32", 1e-5), (15, 4, "f64", 1e-12)])
def test_complex_input_and_complex_istft(mode, m, hop, prec, tol):
    rng = np.random.default_rng(m)
    x = rng.standard_normal(40 * hop + 7) + 1j * rng.standard_normal(40 * hop + 7)
    ref = ShortTimeFFT(hann(m), hop, 1000.0, fft_mode=mode)
    st = StandaloneSTFT(hann(m), hop, 1000.0, fft_mode=mode, precision=prec)
    S = st.stft(x)
    Sref = ref.stft(x)
    assert S.shape == Sref.shape
    assert rel_err(S, Sref) <= tol
    y = st.istft(Sref, complex_output=True)
    yref = ref.istft(Sref)
    assert y.shape == yref.shape and np.iscomplexobj(y)
    assert rel_err(y, yref) <= tol
    assert rel_err(y[:len(x)], x) <= 10 * tol                       # perfect reconstruction of a complex signal
    # the Rust crate's behaviour stays the default: real part only (src/lib.rs:903-907)
    assert rel_err(st.istft(Sref), yref.real) <= tol
    with pytest.raises(VeilsError, match="not allowed"):
        StandaloneSTFT(hann(m), hop, 1000.0, fft_mode="onesided", precision=prec).stft(x)


@pytest.mark.parametrize("mode", ["onesided", "twosided"])
def test_phase_shift_none_matches_scipy(mode):
    rng = np.random.default_rng(3)
    x = rng.standard_normal(3000)
    for m, hop, mfft in ((64, 16, 64), (50, 10, 128), (512, 128, 512)):
        ref = ShortTimeFFT(hann(m), hop, 8000.0, fft_mode=mode, mfft=mfft, phase_shift=None)
        st = StandaloneSTFT(hann(m), hop, 8000.0, fft_mode=mode, mfft=mfft, phase_shift=None, precision="f64")
        assert st.phase_shift is None
        S = st.stft(x)
        assert rel_err(S, ref.stft(x)) <= 1e-12
        assert rel_err(st.istft(S), ref.istft(ref.stft(x)).real) <= 1e-12


def test_nan_screen_on_device_tensors():
    import torch
    m, hop = 512, 128
    x = torch.rand((6, 20000), device="cuda", dtype=torch.float32)
    bad = x.clone()
    bad[4, 12345] = float("nan")
    inf = x.clone()
    inf[2, 777] = float("inf")
    win = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(m) / m))
    # 'sync': raises like the reference (src/lib.rs:696-698)
    st = StandaloneSTFT(win, hop, 16000.0, precision="f32", nan_check="sync")
    st.stft(x)
    with pytest.raises(VeilsError, match="Complex-valued input not allowed for one-sided FFT modes"):
        st.stft(bad)
    st.stft(inf)                                                  # Inf is not NaN: the reference lets it through
    st.spectrogram(x)
    with pytest.raises(VeilsError, match="one-sided"):
        st.spectrogram(bad)
    # NaN at a sample where the periodic Hann window is exactly zero is still found
    edge = x.clone()
    edge[0, 128 * 10 - 256] = float("nan")                       # first sample of slice 10: w[0] = 0
    with pytest.raises(VeilsError):
        st.stft(edge)
    # 'deferred' (default): reported by the next call on the object, or by check_input()
    sd = StandaloneSTFT(win, hop, 16000.0, precision="f32")
    sd.stft(bad)
    with pytest.raises(VeilsError, match="one-sided"):
        sd.check_input()
    sd.stft(x)
    sd.check_input()
    sd.stft(bad)
    torch.cuda.synchronize()
    with pytest.raises(VeilsError):
        sd.stft(x)
    # two-sided layouts are not screened (the reference only checks the one-sided modes)
    s2 = StandaloneSTFT(win, hop, 16000.0, fft_mode="twosided", precision="f32", nan_check="sync")
    assert bool(torch.isnan(s2.stft(bad).real).any())
    # host arrays: the error is returned by the call itself, f64 and the wide family included
    for prec, mm in (("f64", 512), ("f32", 4096)):
        wh = 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(mm) / mm))
        sh = StandaloneSTFT(wh, mm // 4, 16000.0, precision=prec)
        xh = np.random.default_rng(0).standard_normal((2, 30 * mm))
        sh.stft(xh)
        xh[1, 17 * mm + 3] = np.nan
        with pytest.raises(VeilsError, match="one-sided"):
            sh.stft(xh)
