"""The twin's stft_detrend ('constant' / 'linear', python/standalone_stft.py:38-50, 345-371):
oracle vs twin-generated golden vectors on the CPU, CUDA path (stft + rank-2 correction) on the GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util as G  # noqa: E402
from oracle.stft_oracle import OracleSTFT  # noqa: E402

CASES = G.load("detrend")


def _mk(cls, rec, **kw):
    return cls(rec["win"], int(rec["hop"]), float(rec["fs"]), fft_mode=str(rec["mode"]), mfft=int(rec["mfft"]), **kw)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_twin(i):
    rec = CASES[i]
    S = _mk(OracleSTFT, rec).stft_detrend(rec["x"], str(rec["detr"]), padding=str(rec["padding"]))
    assert S.shape == rec["S"].shape and G.rel_err(S, rec["S"]) <= 1e-13


def test_unknown_detrend_type():
    from veils_b200 import StandaloneSTFT, VeilsError
    rec = CASES[0]
    with pytest.raises(ValueError, match="Unknown detrend type: cubic"):
        _mk(OracleSTFT, rec).stft_detrend(rec["x"], "cubic")
    with pytest.raises(VeilsError, match="Unknown detrend type: cubic"):          # twin :50
        _mk(StandaloneSTFT, rec).stft_detrend(rec["x"], "cubic")


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("f64", 1e-12), ("f32", 2e-5)])
def test_gpu_matches_twin(precision, tol):
    import torch
    from veils_b200 import StandaloneSTFT
    for rec in CASES:
        st = _mk(StandaloneSTFT, rec, precision=precision)
        S = st.stft_detrend(rec["x"], str(rec["detr"]), padding=str(rec["padding"]))
        assert S.shape == rec["S"].shape
        # f32: the correction cancels the trend against S, so the error scales with the UN-detrended S
        ref_scale = np.max(np.abs(rec["S"])) if precision == "f64" else np.max(np.abs(_mk(OracleSTFT, rec).stft(
            rec["x"], padding=str(rec["padding"]))))
        assert np.max(np.abs(S - rec["S"])) <= tol * ref_scale, (str(rec["detr"]), str(rec["mode"]))
        dt = torch.float32 if precision == "f32" else torch.float64
        xb = torch.from_numpy(np.stack([rec["x"], 3.0 * rec["x"]])).to("cuda", dt)
        Sb = st.stft_detrend(xb, str(rec["detr"]), padding=str(rec["padding"])).cpu().numpy()
        assert np.max(np.abs(Sb[1] - 3.0 * rec["S"])) <= 3 * tol * ref_scale
    # detr=None is the plain stft
    rec = CASES[0]
    st = _mk(StandaloneSTFT, rec, precision="f64")
    assert np.array_equal(st.stft_detrend(rec["x"], None), st.stft(rec["x"]))
