"""Small all-families run for compute-sanitizer (memcheck / racecheck): packed f32, scalar f64,
generic DFT; forward, spectrogram and inverse with clipped ranges."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from veils_b200 import StandaloneSTFT

rng = np.random.default_rng(0)
for (M, N, H, mode, prec) in [(512, 512, 128, "onesided", "f32"), (512, 400, 160, "twosided", "f32"),
                              (1024, 1024, 256, "centered", "f32"), (4096, 4096, 1024, "onesided", "f32"),
                              (512, 512, 128, "onesided2X", "f64"), (2048, 2048, 512, "onesided", "f64"),
                              (15, 15, 8, "onesided", "f64"), (100, 60, 7, "twosided", "f32")]:
    win = rng.uniform(0.2, 1.0, N)
    x = rng.uniform(-1, 1, (3, 9 * N + 77))
    s = StandaloneSTFT(win, H, 8000.0, fft_mode=mode, mfft=M, precision=prec)
    S = s.stft(x)
    sp = s.spectrogram(x, k_offset=2)
    y = s.istft(S)
    y2 = s.istft(S, H + 1, 5 * N)
    err = np.max(np.abs(y[:, :x.shape[1]] - x))
    print(M, N, H, mode, prec, s.kernel_name(), s.kernel_name(True), f"recon {err:.2e}", flush=True)
    assert err < (1e-4 if prec == "f32" else 1e-11)
print("sanity ok")
