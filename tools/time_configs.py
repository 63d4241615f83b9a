"""Times forward / inverse / spectrogram kernels on BASELINE-shaped sub-batches and prints the
achieved fraction of the HBM roofline (algorithmic bytes)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from veils_b200 import StandaloneSTFT, lib

PEAK = 6551.7
def hann(m): return 0.5 * (1 - np.cos(2 * np.pi * np.arange(m) / m))
CASES = [
    ("c1 f64 2048/512 onesided B=64x44100", hann(2048), 512, "onesided", None, "f64", 64, 44100, False),
    ("c2 f32 512/128 onesided B=1024x160000", hann(512), 128, "onesided", None, "f32", 1024, 160000, False),
    ("c2 f64 512/128 onesided B=256x160000", hann(512), 128, "onesided", None, "f64", 256, 160000, False),
    ("c3 f32 4096/1024 onesided B=128x1440000", hann(4096), 1024, "onesided", None, "f32", 128, 1440000, False),
    ("c4 f32 gauss1024/256 twosided B=1x2^27", np.exp(-0.5 * ((np.arange(1024) - 512) / 128.0) ** 2), 256, "twosided", None, "f32", 1, 1 << 27, False),
    ("c5 f64 8192/64 onesided2X psd spectrogram B=16x160000", hann(8192), 64, "onesided2X", "psd", "f64", 16, 160000, True),
    ("f32 1024/256 onesided B=512x160000", hann(1024), 256, "onesided", None, "f32", 512, 160000, False),
    ("f32 2048/512 onesided B=256x480000", hann(2048), 512, "onesided", None, "f32", 256, 480000, False),
    ("f32 256/64 onesided B=1024x160000", hann(256), 64, "onesided", None, "f32", 1024, 160000, False),
    ("f32 400/160 onesided B=1024x160000", hann(400), 160, "onesided", None, "f32", 1024, 160000, False),
    ("f64 400/160 onesided B=256x160000", hann(400), 160, "onesided", None, "f64", 256, 160000, False),
    ("f32 1000/250 onesided B=256x160000", hann(1000), 250, "onesided", None, "f32", 256, 160000, False),
]
ALIGN = "--align" in sys.argv      # pad S rows to 128-byte lines (what veils_b200.stft does for device tensors)
sel = [a for a in sys.argv[1:] if not a.startswith("--")]
L = lib()
dev = torch.device("cuda", 0)
for name, win, hop, mode, scale, prec, B, n, spec in CASES:
    if sel and not any(s in name for s in sel): continue
    st = StandaloneSTFT(win, hop, 16000.0, fft_mode=mode, scale_to=scale, precision=prec)
    rdt = torch.float32 if prec == "f32" else torch.float64
    cdt = torch.complex64 if prec == "f32" else torch.complex128
    sz = 4 if prec == "f32" else 8
    x = torch.rand((B, n), device=dev, dtype=rdt) * 2 - 1
    P, F = st.p_num(n), st.f_pts
    q = (128 // (sz if spec else 2 * sz)) if ALIGN else 1
    ldp = (P + q - 1) // q * q
    S = torch.empty((B, F, ldp), dtype=rdt if spec else cdt, device=dev)
    def t(fn, reps=5):
        assert fn() == 0, "launch failed: " + (L.veils_last_error() or b"").decode()
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    if spec:
        ms = t(lambda: L.veils_spectrogram_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, S.data_ptr(), ldp, None))
        by = B * (n * sz + F * P * sz)
        print(f"{name}: [{st.kernel_name()}] spectrogram {ms:.3f} ms  {by/ms/1e6:.0f} GB/s  {by/ms/1e6/PEAK*100:.1f}% ; {B*n/ms/1e3:.0f} Msmp/s", flush=True)
        continue
    ms_f = t(lambda: L.veils_stft_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, S.data_ptr(), ldp, None))
    k0, k1 = st.istft_range(P)
    y = torch.empty((B, k1 - k0), dtype=rdt, device=dev)
    ms_i = t(lambda: L.veils_istft_dev(st._h, S.data_ptr(), P, B, ldp, None, None, y.data_ptr(), k1 - k0, None))
    bf = B * (n * sz + F * P * 2 * sz); bi = B * (F * P * 2 * sz + (k1 - k0) * sz)
    err = float((y[:, :n] - x).abs().max())
    print(f"{name}: [{st.kernel_name()}/{st.kernel_name(True)}] fwd {ms_f:.3f} ms {bf/ms_f/1e6/PEAK*100:.1f}%  inv {ms_i:.3f} ms {bi/ms_i/1e6/PEAK*100:.1f}%  "
          f"round trip {B*n/(ms_f+ms_i)/1e3:.0f} Msmp/s ({(bf+bi)/(ms_f+ms_i)/1e6/PEAK*100:.1f}%)  recon {err:.1e}", flush=True)
    del x, S, y
