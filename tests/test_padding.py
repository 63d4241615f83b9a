"""The twin's border modes 'edge' / 'even' / 'odd' (python/standalone_stft.py:309-336):
oracle vs twin-generated golden vectors on the CPU, CUDA path vs the same vectors on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util as G  # noqa: E402
from oracle.stft_oracle import OracleSTFT  # noqa: E402

CASES = G.load("padding")


def _mk(cls, rec, **kw):
    return cls(rec["win"], int(rec["hop"]), float(rec["fs"]), fft_mode=str(rec["mode"]), mfft=int(rec["mfft"]), **kw)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_twin(i):
    rec = CASES[i]
    o = _mk(OracleSTFT, rec)
    S = o.stft(rec["x"], k_offset=int(rec["k_offset"]), padding=str(rec["padding"]))
    assert S.shape == rec["S"].shape
    assert G.rel_err(S, rec["S"]) <= 1e-13


def test_oracle_rejects_unknown_padding():
    rec = CASES[0]
    with pytest.raises(ValueError):
        _mk(OracleSTFT, rec).stft(rec["x"], padding="wrap")


def test_mirror_rejects_unknown_padding():
    from veils_b200 import StandaloneSTFT, VeilsError
    rec = CASES[0]
    with pytest.raises(VeilsError, match="padding"):
        _mk(StandaloneSTFT, rec).stft(rec["x"], padding="wrap")


def test_pad_extent_is_host_only():
    """veils_pad_extent needs no GPU: left = samples the first slice reaches before x[0]."""
    import ctypes as C
    from veils_b200 import StandaloneSTFT, lib
    rec = CASES[0]
    st = _mk(StandaloneSTFT, rec)
    n = len(rec["x"])
    left, n_ext = C.c_int64(), C.c_int64()
    assert lib().veils_pad_extent(st._h, n, None, None, 0, C.byref(left), C.byref(n_ext)) == 0
    lo = st.p_min * st.hop - st.m_num_mid
    hi = (st.p_max(n) - 1) * st.hop - st.m_num_mid + st.m_num
    assert left.value == max(0, -lo) and n_ext.value == left.value + n + max(0, hi - n)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("f64", 1e-12), ("f32", 1e-5)])
def test_gpu_matches_twin(precision, tol):
    import torch
    from veils_b200 import StandaloneSTFT
    for rec in CASES:
        st = _mk(StandaloneSTFT, rec, precision=precision)
        S = st.stft(rec["x"], k_offset=int(rec["k_offset"]), padding=str(rec["padding"]))
        assert S.shape == rec["S"].shape
        assert G.rel_err(S, rec["S"]) <= tol, (str(rec["padding"]), str(rec["mode"]))
        # device-tensor flavour, batched, and the spectrogram epilogue
        dt = torch.float32 if precision == "f32" else torch.float64
        xb = torch.from_numpy(np.stack([rec["x"], -2.0 * rec["x"]])).to("cuda", dt)
        Sb = st.stft(xb, k_offset=int(rec["k_offset"]), padding=str(rec["padding"])).cpu().numpy()
        assert G.rel_err(Sb[0], rec["S"]) <= tol and G.rel_err(Sb[1], -2.0 * rec["S"]) <= tol
        sp = st.spectrogram(xb, k_offset=int(rec["k_offset"]), padding=str(rec["padding"])).cpu().numpy()
        assert G.rel_err(sp[0], np.abs(rec["S"]) ** 2) <= 10 * tol
