"""GPU parity tests: the CUDA path, called through the C ABI, against (a) the golden vectors made
by the reference's own Python twin and (b) the oracle on seeded inputs.
Tolerances (BASELINE.json north_star): <= 1e-12 relative in f64, <= 1e-5 relative in f32,
relative = ||a-b||_inf / ||b||_inf over the whole array."""
import numpy as np
import pytest

from golden_util import load, opt, rel_err
from oracle.stft_oracle import OracleSTFT
from veils_b200 import StandaloneSTFT

pytestmark = pytest.mark.gpu
TOL = {"f64": 1e-12, "f32": 1e-5}


def _mk(r, precision, **kw):
    return StandaloneSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                          mfft=int(r["mfft"]), phase_shift=int(r.get("phase_shift", 0)),
                          precision=precision, **kw)


@pytest.mark.parametrize("precision", ["f64", "f32"])
@pytest.mark.parametrize("name", ["basic_usage", "integration", "final_comparison", "sweep",
                                  "baseline_shapes"])
def test_golden_vectors(name, precision):
    tol = TOL[precision]
    for r in load(name):
        s = _mk(r, precision)
        S = s.stft(r["x"], opt(r["p0"]), opt(r["p1"]), k_offset=int(r["k_offset"]))
        assert S.shape == r["S"].shape
        assert rel_err(S, r["S"]) <= tol, (name, r.get("name"), s.kernel_name())
        if "y" in r:
            y = s.istft(r["S"], int(r["k0"]), opt(r["k1"]))
            assert y.shape == r["y"].shape
            assert rel_err(y, r["y"]) <= tol, (name, r.get("name"))


def test_basic_usage_golden_coefficients():
    """examples/basic_usage.rs:81-139: MSE < 1e-12 vs the printed scipy coefficients."""
    r = load("basic_usage")[0]
    S = _mk(r, "f64").stft(r["x"])
    assert S.shape == (8, 5)
    assert np.mean(np.abs(S[:3, :5] - r["scipy_3x5"]) ** 2) < 1e-12


def test_integration_perfect_reconstruction():
    """tests/integration_tests.rs:112-318, :355-390 through the C ABI."""
    for r in load("integration"):
        s = _mk(r, "f64")
        S = s.stft(r["x"])
        assert S.shape == ((9 if s.onesided_fft else 16), 19)
        y = s.istft(S)
        assert len(y) == 76
        err = np.max(np.abs(y[:64] - r["x"]))
        assert err < 1e-14, (r["name"], r["mode"], err)
        S2 = s.stft(r["x"])
        assert np.array_equal(S, S2)                      # determinism (:355-390)
        assert np.array_equal(y, s.istft(S2))


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_scipy_scaling_and_spectrogram(precision):
    tol = TOL[precision]
    for r in load("scipy_scaling"):
        s = StandaloneSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                           mfft=int(r["mfft"]), scale_to=str(r["scale"]), precision=precision)
        assert rel_err(s.stft(r["x"]), r["S"]) <= tol
        assert rel_err(s.spectrogram(r["x"]), r["spec"]) <= 4 * tol
        assert rel_err(s.istft(r["S"]), r["y"]) <= tol


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_batched_against_oracle(precision):
    rng = np.random.default_rng(11)
    tol = TOL[precision]
    for (m, hop, mfft, mode) in [(64, 16, 64, "onesided"), (128, 32, 128, "twosided"),
                                 (256, 64, 256, "centered"), (512, 128, 512, "onesided"),
                                 (512, 128, 512, "onesided2X"), (100, 25, 128, "onesided"),
                                 (1024, 256, 1024, "twosided"), (48, 12, 48, "onesided")]:
        win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(m) / m))
        x = rng.uniform(-1, 1, (5, 4 * m + 37))
        o = OracleSTFT(win, hop, 16000.0, fft_mode=mode, mfft=mfft)
        s = StandaloneSTFT(win, hop, 16000.0, fft_mode=mode, mfft=mfft, precision=precision)
        S = s.stft(x)
        Sref = np.stack([o.stft(r) for r in x])
        assert rel_err(S, Sref) <= tol, (m, mode, s.kernel_name())
        y = s.istft(Sref)
        yref = np.stack([o.istft(r) for r in Sref])
        assert rel_err(y, yref) <= tol, (m, mode, s.kernel_name(True))
        # round trip through the device path only
        y2 = s.istft(S)
        assert np.max(np.abs(y2[:, :x.shape[1]] - x)) <= 50 * tol


def test_torch_device_pointer_path():
    import torch
    rng = np.random.default_rng(5)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(512) / 512))
    x = rng.uniform(-1, 1, (3, 4000)).astype(np.float32)
    s = StandaloneSTFT(win, 128, 16000.0, precision="f32")
    o = OracleSTFT(win, 128, 16000.0)
    xt = torch.from_numpy(x).cuda()
    S = s.stft(xt)
    torch.cuda.synchronize()
    Sref = np.stack([o.stft(r.astype(np.float64)) for r in x])
    assert rel_err(S.cpu().numpy(), Sref) <= 1e-5
    y = s.istft(S)
    torch.cuda.synchronize()
    assert np.max(np.abs(y.cpu().numpy()[:, :4000] - x)) <= 1e-4
    sp = s.spectrogram(xt)
    assert rel_err(sp.cpu().numpy(), np.abs(Sref) ** 2) <= 1e-5
