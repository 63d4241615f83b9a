"""Pins oracle/stft_ref.c (the C restatement used as the timed CPU baseline) to the numpy oracle
and to the reference-derived golden vectors (<= 1e-12 relative, the reference's own
cross-implementation threshold, python/final_comparison.py:53,86)."""
import numpy as np
import pytest

from golden_util import load, opt, rel_err
from oracle.ref_c import RefC, num_threads
from oracle.stft_oracle import OracleSTFT


@pytest.mark.parametrize("name", ["basic_usage", "integration", "final_comparison", "sweep",
                                  "baseline_shapes"])
def test_c_port_matches_golden(name):
    for r in load(name):
        o = OracleSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                       mfft=int(r["mfft"]), phase_shift=int(r.get("phase_shift", 0)))
        c = RefC(o)
        S = c.stft(r["x"], opt(r["p0"]), opt(r["p1"]), int(r["k_offset"]))[0]
        assert rel_err(S, r["S"]) <= 1e-12
        if "y" in r:
            y = c.istft(r["S"], int(r["k0"]), opt(r["k1"]))[0]
            assert rel_err(y, r["y"]) <= 1e-12


def test_c_port_batched_threads():
    assert num_threads() >= 1
    rng = np.random.default_rng(1)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(512) / 512))
    o = OracleSTFT(win, 128, 16000.0)
    x = rng.uniform(-1, 1, (6, 5000))
    c = RefC(o)
    S = c.stft(x)
    assert rel_err(S, np.stack([o.stft(r) for r in x])) <= 1e-12
    y = c.istft(S)
    assert np.max(np.abs(y[:, :5000] - x)) < 1e-12
