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


@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_randomized_pow2_geometry_against_oracle(precision):
    """Randomized geometry for the power-of-two kernel families (packed f32, scalar f32/f64):
    window length <= mfft, arbitrary hop, roll, k_offset, slice sub-ranges, all four layouts,
    long enough for several tiles / chunks of the streaming overlap-add."""
    rng = np.random.default_rng(2026)
    tol = TOL[precision]
    seen = set()
    for it in range(40):
        M = int(2 ** rng.integers(5, 12))
        N = int(M if rng.random() < 0.6 else rng.integers(M // 2 + 1, M + 1))
        H = int(rng.choice([N // 4, N // 2, max(2, N // 3), max(2, N // 8), int(rng.integers(2, N + 1))]))
        H = max(2, min(H, N))
        mode = str(rng.choice(["onesided", "twosided", "centered", "onesided2X"]))
        ps = int(rng.integers(-M + 1, M)) if rng.random() < 0.3 else 0
        scale = rng.choice([None, None, "magnitude", "psd"])
        win = rng.uniform(0.2, 1.0, N)
        n = int(rng.integers(6 * N, 6 * N + 40 * H + 100))
        x = rng.uniform(-1, 1, (3, n))
        if precision == "f32":
            x = x.astype(np.float32).astype(np.float64)
        o = OracleSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale)
        s = StandaloneSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale,
                           precision=precision)
        seen.add((s.kernel_name(), s.kernel_name(True)))
        koff = int(rng.integers(-3, 4))
        Sref = np.stack([o.stft(r, k_offset=koff) for r in x])
        S = s.stft(x, k_offset=koff)
        assert rel_err(S, Sref) <= tol, (M, N, H, mode, ps, s.kernel_name())
        sp = s.spectrogram(x, k_offset=koff)
        assert rel_err(sp, np.abs(Sref) ** 2) <= 4 * tol
        if o.invertible:
            S0 = np.stack([o.stft(r) for r in x])
            yref = np.stack([o.istft(q) for q in S0])
            y = s.istft(S0)
            assert rel_err(y, yref) <= tol, (M, N, H, mode, ps, s.kernel_name(True))
            k0 = int(rng.integers(0, n // 3))
            k1 = int(rng.integers(k0 + N, n))
            assert rel_err(s.istft(S0, k0, k1), np.stack([o.istft(q, k0, k1) for q in S0])) <= tol
    assert len(seen) >= 2


@pytest.mark.gpu
def test_cross_spectrogram_matches_scipy():
    """scipy ShortTimeFFT.spectrogram(x, y) = stft(x) * conj(stft(y)) (_short_time_fft.py:1401-1404)."""
    from scipy.signal import ShortTimeFFT
    rng = np.random.default_rng(11)
    for mode, m, hop, mfft, scale in (("onesided", 64, 16, 64, None), ("onesided2X", 32, 8, 64, "psd"),
                                      ("centered", 24, 6, 32, "magnitude"), ("twosided", 15, 5, 15, None)):
        win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(m) / m)) + 0.01
        x, y = rng.standard_normal(500), rng.standard_normal(500)
        ref = ShortTimeFFT(win, hop, fs=4.0, fft_mode=mode, mfft=mfft, scale_to=scale).spectrogram(x, y)
        for prec, tol in (("f64", 1e-12), ("f32", 1e-5)):
            s = StandaloneSTFT(win, hop, 4.0, fft_mode=mode, mfft=mfft, scale_to=scale, precision=prec)
            got = s.spectrogram(x, y)
            assert got.shape == ref.shape and np.iscomplexobj(got)
            assert rel_err(got, ref) <= 4 * tol, (mode, prec)
            assert rel_err(s.spectrogram(x, x), np.abs(s.stft(x)) ** 2) <= 4 * tol      # y = x: auto-spectrogram


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f32", "f64"])
def test_randomized_device_tensor_path_staged_kernels(precision):
    """The device-tensor flavour (128-byte row pitch) is what selects the TMA-staged forward kernel and
    the paired inverse pass 1 of the packed family.  Randomized geometry through torch tensors:
    one-sided layouts (staged / paired paths) and two-sided ones (plain-store paths), all mfft of the
    family, tiles cut by P, slice sub-ranges, batches, |S|^2, istft sub-ranges -- against the oracle."""
    import torch
    rng = np.random.default_rng(909)
    tol = TOL[precision]
    npdt = np.float32 if precision == "f32" else np.float64
    kinds = set()
    for it in range(36):
        M = int(2 ** rng.integers(5, 14))
        N = int(M if rng.random() < 0.7 else 2 * rng.integers(M // 4 + 1, M // 2 + 1))
        H = int(2 * max(1, rng.choice([N // 8, N // 4, N // 6, N // 2, N // 16]) // 1 // 2 * 1))
        H = max(2, min(H - H % 2, N))
        mode = str(rng.choice(["onesided", "onesided", "onesided2X", "twosided", "centered"]))
        ps = int(2 * rng.integers(-M // 4, M // 4)) if rng.random() < 0.3 else 0
        scale = rng.choice([None, None, "psd"])
        win = rng.uniform(0.2, 1.0, N)
        n_sl = int(rng.integers(40, 140))
        n = int(N + n_sl * H + rng.integers(0, H))
        B = int(rng.integers(1, 4))
        x = rng.uniform(-1, 1, (B, n)).astype(npdt)
        o = OracleSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale)
        s = StandaloneSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, phase_shift=ps, scale_to=scale, precision=precision)
        kinds.add((s.kernel_name(), s.kernel_name(True), mode in ("onesided", "onesided2X")))
        xt = torch.from_numpy(x).cuda()
        xd = x.astype(np.float64)
        Sref = np.stack([o.stft(r) for r in xd])
        S = s.stft(xt)
        assert S.stride(1) % (16 if precision == "f32" else 8) == 0 and S.shape == Sref.shape
        assert rel_err(S.cpu().numpy(), Sref) <= tol, (M, N, H, mode, ps, s.kernel_name())
        p0 = int(rng.integers(s.p_min, s.p_min + 5))
        p1 = int(rng.integers(p0 + 3, s.p_max(n) + 1))
        Ssub = s.stft(xt, p0, p1)
        assert rel_err(Ssub.cpu().numpy(), Sref[:, :, p0 - s.p_min:p1 - s.p_min]) <= tol
        assert rel_err(s.spectrogram(xt).cpu().numpy(), np.abs(Sref) ** 2) <= 4 * tol
        if o.invertible:
            yref = np.stack([o.istft(q) for q in Sref])
            y = s.istft(S)                                  # padded-pitch view straight back in
            assert rel_err(y.cpu().numpy(), yref) <= tol, (M, N, H, mode, ps, s.kernel_name(True))
            k0 = int(rng.integers(0, n // 3))
            k1 = int(rng.integers(k0 + N, n))
            assert rel_err(s.istft(S, k0, k1).cpu().numpy(), np.stack([o.istft(q, k0, k1) for q in Sref])) <= tol
            assert torch.equal(y, s.istft(S))               # deterministic overlap-add
    assert any(k[2] for k in kinds) and any(not k[2] for k in kinds)
    if precision == "f32":
        assert any(k[0] == "pow2_packed" and k[2] for k in kinds) and any(k[0] == "pow2_tile" and k[2] for k in kinds)


@pytest.mark.gpu
def test_unaligned_output_views_take_the_plain_store_path():
    """S rows that are not 16-byte aligned (odd pitch, or a view starting at an odd column) cannot be
    described by a tensor map: the same call must silently run the plain-store kernels and give
    bit-identical results (the arithmetic per slice is the same; only the way the tile leaves the SM
    differs for the packed family's paired / two-slices-per-lane last passes -> compare to tolerance)."""
    import torch
    from veils_b200 import lib
    L = lib()
    rng = np.random.default_rng(5)
    win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(512) / 512))
    for prec, dt, cdt, tol in (("f32", torch.float32, torch.complex64, 2e-6), ("f64", torch.float64, torch.complex128, 1e-14)):
        s = StandaloneSTFT(win, 128, 16000.0, precision=prec)
        x = torch.from_numpy(rng.uniform(-1, 1, (3, 20000))).to("cuda", dt)
        n = x.shape[1]
        P, F = s.p_num(n), s.f_pts
        ref = s.stft(x)                                               # aligned pitch: staged kernels
        st = torch.cuda.current_stream().cuda_stream
        for ldp, off in ((P, 0), (P + 3, 1), (P + 16, 1)):            # odd pitch / odd column offset
            big = torch.zeros((3, F, ldp + 2), dtype=cdt, device="cuda")
            view = big[:, :, off:off + P]
            rc = L.veils_stft_dev(s._h, x.data_ptr(), n, 3, n, None, None, 0, view.data_ptr(), ldp + 2, st)
            assert rc == 0, L.veils_last_error()
            torch.cuda.synchronize()
            assert rel_err(view.cpu().numpy(), ref.cpu().numpy()) <= tol
            assert float(big[:, :, :off].abs().max()) == 0.0 if off else True      # nothing written outside
            assert float(big[:, :, off + P:].abs().max()) == 0.0
            y = torch.empty((3, s.istft_range(P)[1]), dtype=dt, device="cuda")
            rc = L.veils_istft_dev(s._h, view.data_ptr(), P, 3, ldp + 2, None, None, y.data_ptr(), y.shape[1], st)
            assert rc == 0, L.veils_last_error()
            assert float((y[:, :n] - x).abs().max()) <= 10 * tol
