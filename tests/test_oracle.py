"""Pins oracle/stft_oracle.py to the reference: golden vectors made by the reference's own
Python twin (oracle/make_golden.py), the scipy coefficients printed in
examples/basic_usage.rs:13-40, the thresholds of tests/integration_tests.rs, and scipy."""
import numpy as np
import pytest

from golden_util import load, opt, rel_err
from oracle.stft_oracle import OracleSTFT

TOL = 1e-13


def _mk(r, **kw):
    return OracleSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                      mfft=int(r["mfft"]), phase_shift=int(r.get("phase_shift", 0)), **kw)


def _check(r):
    o = _mk(r)
    n = len(r["x"])
    assert o.p_min == r["p_min"] and o.p_max(n) == r["p_max"]
    assert o.k_min == r["k_min"] and o.k_max(n) == r["k_max"] and o.f_pts == r["f_pts"]
    S = o.stft(r["x"], opt(r["p0"]), opt(r["p1"]), int(r["k_offset"]))
    assert S.shape == r["S"].shape
    assert rel_err(S, r["S"]) <= TOL
    if "y" in r:
        assert rel_err(o.dual_win, r["dual"]) <= TOL
        y = o.istft(r["S"], int(r["k0"]), opt(r["k1"]))
        assert y.shape == r["y"].shape
        assert rel_err(y, r["y"]) <= TOL


@pytest.mark.parametrize("name", ["basic_usage", "integration", "final_comparison", "sweep",
                                  "baseline_shapes"])
def test_oracle_matches_twin_fixtures(name):
    recs = load(name)
    assert recs
    for r in recs:
        _check(r)


def test_basic_usage_scipy_coefficients():
    """examples/basic_usage.rs:81-139: MSE < 1e-12 against the printed scipy values."""
    r = load("basic_usage")[0]
    S = _mk(r).stft(r["x"])
    assert S.shape == (8, 5)
    mse = np.mean(np.abs(S[:3, :5] - r["scipy_3x5"]) ** 2)
    assert mse < 1e-12


def test_integration_thresholds():
    """tests/integration_tests.rs:112-318: reconstruction error per signal / mode."""
    for r in load("integration"):
        o = _mk(r)
        S = o.stft(r["x"])
        assert S.shape == ((9 if o.onesided_fft else 16), 19)
        y = o.istft(S)
        err = np.max(np.abs(y[:64] - r["x"]))
        assert err <= (0.0 if r["name"] == "impulse" else 1e-15) + 1e-15
        assert len(y) == 76
    o = _mk(load("integration")[0])
    assert (o.m_num, o.m_num_mid, o.f_pts) == (16, 8, 9)          # :321-339
    assert o.f()[0] == 0.0 and abs(o.f()[8] - 500.0) < 1e-10
    assert o.t(64)[0] == o.p_min * o.delta_t                        # :342-351


def test_scipy_scaling_and_spectrogram():
    for r in load("scipy_scaling"):
        o = OracleSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                       mfft=int(r["mfft"]), scale_to=str(r["scale"]))
        assert rel_err(o.win, r["win_scaled"]) <= TOL
        assert rel_err(o.dual_win, r["dual"]) <= TOL
        assert rel_err(o.stft(r["x"]), r["S"]) <= TOL
        assert rel_err(o.spectrogram(r["x"]), r["spec"]) <= TOL
        assert rel_err(o.istft(r["S"]), r["y"]) <= TOL


def test_oracle_matches_live_scipy():
    """Second oracle: scipy.signal.ShortTimeFFT on fresh random cases (scipy is in the image)."""
    from scipy.signal import ShortTimeFFT
    rng = np.random.default_rng(3)
    for _ in range(60):
        m = int(rng.integers(4, 50))
        hop = int(rng.integers(1, m + 1))
        mfft = m + int(rng.integers(0, 6))
        mode = str(rng.choice(["onesided", "twosided", "centered"]))
        ps = int(rng.integers(-m + 1, m))
        win = rng.uniform(0.1, 1.0, m)
        x = rng.standard_normal(int(rng.integers(m, 5 * m + 9)))
        st = ShortTimeFFT(win, hop, fs=3.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
        o = OracleSTFT(win, hop, 3.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
        assert (o.p_min, o.p_max(len(x)), o.k_min) == (st.p_min, st.p_max(len(x)), st.k_min)
        S = st.stft(x)
        assert rel_err(o.stft(x), S) <= TOL
        assert rel_err(o.istft(S), np.real(st.istft(S))) <= TOL


def test_errors():
    w = np.hanning(16)
    for bad, msg in [(dict(win=[], hop=4), "non-empty"), (dict(win=[np.nan, 1], hop=1), "finite"),
                     (dict(win=w, hop=0), "not >= 1"), (dict(win=w, hop=4, fft_mode="x"), "Unknown")]:
        with pytest.raises(ValueError, match=msg):
            OracleSTFT(bad["win"], bad["hop"], 1.0, **{k: v for k, v in bad.items()
                                                        if k not in ("win", "hop")})
    o = OracleSTFT(np.ones(8), 9, 1.0)
    assert not o.invertible
    o = OracleSTFT(w, 4, 1.0)
    with pytest.raises(ValueError, match="Invalid slice range"):
        o.stft(np.zeros(64), p0=5, p1=5)
    with pytest.raises(ValueError, match="must be >= ceil"):
        o.stft(np.zeros(3))
