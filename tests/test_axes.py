"""Axis / border helpers of the host mirror against live scipy.signal.ShortTimeFFT."""
import numpy as np
import pytest
from scipy.signal import ShortTimeFFT

from veils_b200 import StandaloneSTFT


def _wins():
    rng = np.random.default_rng(3)
    yield np.hanning(16), 4, None
    yield 0.5 * (1 - np.cos(2 * np.pi * np.arange(15) / 15)), 5, 20
    w = rng.uniform(0.1, 1, 33)
    w[:3] = 0
    w[-2:] = 0
    yield w, 7, 64
    yield np.ones(8), 8, None
    yield np.hanning(64), 1, 64


@pytest.mark.parametrize("mode", ["onesided", "onesided2X", "centered", "twosided"])
def test_border_and_extent_helpers(mode):
    for win, hop, mfft in _wins():
        kw = dict(scale_to="psd") if mode == "onesided2X" else {}
        ref = ShortTimeFFT(win, hop, fs=100.0, fft_mode=mode, mfft=mfft, **kw)
        st = StandaloneSTFT(win, hop, 100.0, fft_mode=mode, mfft=mfft, **kw)
        assert st.lower_border_end == ref.lower_border_end
        for n in (len(win), 3 * len(win) + 1, 1000):
            assert st.upper_border_begin(n) == ref.upper_border_begin(n)
            assert np.allclose(st.t(n), ref.t(n), rtol=1e-14, atol=0)     # Rust divides by fs (src/lib.rs:916-928), scipy multiplies by T
            if mode != "twosided":
                for seq in ("tf", "ft"):
                    for cb in (False, True):
                        assert st.extent(n, seq, cb) == pytest.approx(ref.extent(n, seq, cb), rel=1e-15)
        for k in (-9, 0, 5, 64, 1001):
            assert st.nearest_k_p(k) == ref.nearest_k_p(k)
            assert st.nearest_k_p(k, left=False) == ref.nearest_k_p(k, left=False)
        assert (st.p_min, st.k_min, st.delta_t, st.delta_f) == (ref.p_min, ref.k_min, ref.delta_t, ref.delta_f)
