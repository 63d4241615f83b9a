"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/veils_cuda.h declares, and its host planner (geometry, dual window, validation,
error texts) agrees with the reference-derived golden vectors.  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from golden_util import load, opt, rel_err
from veils_b200 import StandaloneSTFT, VeilsError, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "veils_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(veils_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    syms = _header_symbols()
    assert len(syms) >= 45
    h = C.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(h, s), f"{s} declared in veils_cuda.h but not exported"
    assert set(_lib.PROTOTYPES) == set(syms)
    assert _lib.lib().veils_abi_version() == 1


def test_rust_ffi_mirrors_the_header():
    """rust/src/ffi.rs is generated from include/veils_cuda.h (tools/gen_rust_ffi.py): every declaration of the
    header -- and nothing else -- has its `extern "C"` item, and the crate shell keeps the reference's accessors."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_rust_ffi", os.path.join(root, "tools", "gen_rust_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text = open(os.path.join(root, "rust", "src", "ffi.rs")).read()
    assert text == gen.generate(), "rust/src/ffi.rs is stale: run python tools/gen_rust_ffi.py"
    fns = set(re.findall(r"pub fn (veils_[a-z0-9_]+)\(", text))
    assert fns == set(_header_symbols())
    shell = open(os.path.join(root, "rust", "src", "lib.rs")).read()
    for item in ("pub fn fft_mode(&self) -> &FftMode", "pub fn debug_ifft_func(", "pub enum FftMode", "fn is_onesided",
                 "pub fn stft(", "pub fn istft(", "pub fn spectrogram(", "pub fn p_range(", "pub fn dual_win("):
        assert item in shell, item


def test_debug_ifft_func_matches_the_oracle():
    """src/lib.rs:952-954 -- host-only, no GPU needed"""
    from oracle.stft_oracle import OracleSTFT
    rng = np.random.default_rng(2)
    for mode in ("onesided", "onesided2X", "twosided", "centered"):
        for m, mfft, ps in ((16, 16, 0), (15, 20, 3), (8, 8, -2)):
            win = rng.uniform(0.2, 1.0, m)
            o = OracleSTFT(win, 4, 1.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
            s = StandaloneSTFT(win, 4, 1.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
            X = rng.standard_normal(o.f_pts) + 1j * rng.standard_normal(o.f_pts)
            if mode.startswith("onesided"):
                X[0] = X[0].real
                if mfft % 2 == 0:
                    X[-1] = X[-1].real
            assert rel_err(s.debug_ifft_func(X), o._ifft_func(X[None, :])[0]) <= 1e-13


def _mk(r, **kw):
    return StandaloneSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                          mfft=int(r["mfft"]), phase_shift=int(r.get("phase_shift", 0)), **kw)


@pytest.mark.parametrize("name", ["basic_usage", "integration", "final_comparison", "sweep",
                                  "baseline_shapes"])
def test_planner_geometry_matches_reference(name):
    for r in load(name):
        s = _mk(r)
        n = len(r["x"])
        assert (s.p_min, s.p_max(n), s.k_min, s.k_max(n), s.f_pts) == \
            (r["p_min"], r["p_max"], r["k_min"], r["k_max"], r["f_pts"])
        assert s.p_num(n) == r["p_max"] - r["p_min"]
        if "y" in r:
            assert rel_err(s.dual_win, r["dual"]) <= 1e-14
            a, b = s.istft_range(r["S"].shape[1], int(r["k0"]), opt(r["k1"]))
            assert b - a == len(r["y"])
        p0, p1 = s.p_range(n, opt(r["p0"]), opt(r["p1"]))
        assert p1 - p0 == r["S"].shape[1]


def test_scaling_matches_scipy():
    for r in load("scipy_scaling"):
        s = StandaloneSTFT(r["win"], int(r["hop"]), float(r["fs"]), fft_mode=str(r["mode"]),
                           mfft=int(r["mfft"]), scale_to=str(r["scale"]))
        assert s.scaling == r["scale"]
        assert rel_err(s.win, r["win_scaled"]) <= 1e-14
        assert rel_err(s.dual_win, r["dual"]) <= 1e-14
        assert (s.p_min, s.p_max(len(r["x"]))) == (r["p_min"], r["p_max"])


def test_properties_and_axes():
    """tests/integration_tests.rs:55-73, 321-351."""
    w = 0.5 * (1 - np.cos(2 * np.pi * np.arange(16) / 15))
    s = StandaloneSTFT(w, 4, 1000.0)
    assert (s.m_num, s.m_num_mid, s.f_pts, s.hop, s.mfft, s.fs) == (16, 8, 9, 4, 16, 1000.0)
    assert s.fft_mode == "onesided" and s.onesided_fft and s.invertible
    f = s.f()
    assert f[0] == 0.0 and abs(f[8] - 500.0) < 1e-10
    t = s.t(64)
    assert len(t) == 19 and t[0] == s.p_min * s.delta_t
    assert abs(s.delta_f - 62.5) < 1e-12 and abs(s.delta_t - 0.004) < 1e-15
    for mode, fp in [("twosided", 16), ("centered", 16), ("onesided2X", 9)]:
        assert StandaloneSTFT(w, 4, 1000.0, fft_mode=mode).f_pts == fp
    c = StandaloneSTFT(w, 4, 1000.0, fft_mode="centered").f()
    assert np.allclose(c, np.fft.fftshift(np.fft.fftfreq(16, 1e-3)))


def test_reference_error_texts():
    """src/lib.rs:176-207, :75, :289-317, :605 -- same messages through the ABI."""
    w = np.hanning(16)
    cases = [(dict(win=[], hop=4), "Parameter win must be non-empty!"),
             (dict(win=[1.0, np.inf], hop=1), "Parameter win must have finite entries!"),
             (dict(win=w, hop=0), "Parameter hop=0 is not >= 1!"),
             (dict(win=w, hop=4, fft_mode="invalid_mode"), "Unknown fft_mode: invalid_mode"),
             (dict(win=w, hop=4, dual_win=np.ones(3)), "dual_win.len() must equal win.len()!"),
             (dict(win=w, hop=4, dual_win=np.full(16, np.nan)), "Parameter dual_win must have finite entries!")]
    for kw, msg in cases:
        with pytest.raises(VeilsError) as ei:
            StandaloneSTFT(kw.pop("win"), kw.pop("hop"), 1.0, **kw)
        assert str(ei.value) == msg
    s = StandaloneSTFT(np.ones(8), 9, 1.0)
    assert not s.invertible
    with pytest.raises(VeilsError, match="hop=9 is larger than window length of 8 => STFT not invertible!"):
        _ = s.dual_win
    z = StandaloneSTFT(np.r_[np.zeros(4), np.ones(4)] * 1.0, 8, 1.0)      # zero gap -> DD == 0
    z2 = StandaloneSTFT(np.r_[0.0, np.ones(7)], 8, 1.0)
    assert not z2.invertible
    with pytest.raises(VeilsError, match="Short-time Fourier Transform not invertible!"):
        _ = z2.dual_win
    assert not z.invertible
    s = StandaloneSTFT(w, 4, 1.0)
    with pytest.raises(VeilsError, match="Invalid slice range: p0=5, p1=5"):
        s.p_range(64, 5, 5)
    with pytest.raises(VeilsError, match=r"Parameter n must be >= ceil\(m_num/2\) = 8!"):
        s.p_max(3)                       # the reference panics here (src/lib.rs:562-564)
    with pytest.raises(VeilsError, match="needs to have at least"):
        s.istft_range(1)
    with pytest.raises(VeilsError, match="k0=10 must be < k1=5"):
        s.istft_range(19, 10, 5)
    with pytest.raises(VeilsError, match="Output length 3 has to be at least 8"):
        s.istft_range(19, 0, 3)


def test_compute_without_gpu_fails_loudly():
    """No CPU fallback: on a box without CUDA the compute entry points raise."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    s = StandaloneSTFT(np.hanning(16), 4, 1.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        s.stft(np.zeros(64))


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under veils_b200/ (Python or C++/CUDA sources) may
    import, include or link anything from oracle/ -- there is no CPU fallback to fall back to."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bad = []
    for dirpath, _, files in os.walk(os.path.join(root, "veils_b200")):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            if re.search(r"(from|import)\s+oracle|#include\s+[\"<][^\">]*oracle|stft_ref|libstft_ref", text):
                bad.append(os.path.join(dirpath, f))
    assert not bad, bad
    import subprocess
    import sys
    code = "import sys, veils_b200; veils_b200.lib(); assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=root)


def test_rust_build_script_lists_every_source():
    """rust/build.rs compiles the same translation units as veils_b200/build.py."""
    import re
    from veils_b200 import build as vb
    txt = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rust", "build.rs")).read()
    listed = set(re.findall(r'"([\w]+\.(?:cu|cpp))"', txt))
    assert listed == set(vb.SOURCES)
