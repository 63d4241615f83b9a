#!/usr/bin/env python
"""Generate tests/golden/*.npz from the REFERENCE's own Python twin (and scipy for the two
features the reference lacks).  TEST INFRASTRUCTURE -- runs only in the build container,
where /root/reference exists; the fixtures it writes are committed and travel to the GPU box.

    python oracle/make_golden.py            # rewrites tests/golden/
    python oracle/make_golden.py padding    # rewrites only the named fixture(s)

Sources of truth used here (nothing is copied from the reference, it is imported/executed):
  * /root/reference/python/standalone_stft.py  -- StandaloneSTFT (twin of src/lib.rs)
  * /root/reference/python/final_comparison.py -- create_test_data() + the 8 configs (:250-260)
  * examples/basic_usage.rs:13-40              -- golden scipy coefficients (typed in below as the
                                                  8-digit known answers they are)
  * scipy.signal.ShortTimeFFT                  -- scale_to / spectrogram (absent from the reference)
"""
import os
import sys

import numpy as np

REF = "/root/reference/python"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def hann_periodic(n):
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(n) / n))


def hann_symmetric(n):
    return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(n) / (n - 1)))


def gaussian_periodic(n, std):
    k = np.arange(n) - n / 2.0
    return np.exp(-0.5 * (k / std) ** 2)


def run_twin(Twin, win, hop, fs, mode, x, mfft=None, phase_shift=0, dual_win=None,
             p0=None, p1=None, k_offset=0, k0=0, k1=None, want_istft=True):
    tw = Twin(win=np.asarray(win, float), hop=hop, fs=fs, fft_mode=mode, mfft=mfft,
              dual_win=dual_win, phase_shift=phase_shift)
    S = tw.stft(np.asarray(x, float), p0, p1, k_offset=k_offset)
    rec = dict(win=np.asarray(win, float), hop=hop, fs=fs, mode=mode, x=np.asarray(x, float),
               mfft=tw.mfft, phase_shift=phase_shift, k_offset=k_offset,
               p0=-(10 ** 9) if p0 is None else p0, p1=-(10 ** 9) if p1 is None else p1,
               S=S, p_min=tw.p_min, p_max=tw.p_max(len(x)), k_min=tw.k_min,
               k_max=tw.k_max(len(x)), f_pts=tw.f_pts)
    if dual_win is not None:
        rec["dual_in"] = np.asarray(dual_win, float)
    if want_istft and tw.invertible and p0 is None and p1 is None and k_offset == 0:
        y = tw.istft(S, k0, k1)
        rec.update(k0=k0, k1=-(10 ** 9) if k1 is None else k1, y=np.real(y),
                   dual=np.asarray(tw.dual_win, float))
    return rec


ONLY = set(sys.argv[1:])


def save(name, recs):
    if ONLY and name not in ONLY:
        return
    flat = {}
    for i, r in enumerate(recs):
        for k, v in r.items():
            flat[f"{i}:{k}"] = np.asarray(v)
    flat["count"] = np.asarray(len(recs))
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **flat)
    print(f"{name}: {len(recs)} cases, {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    sys.path.insert(0, REF)
    from standalone_stft import StandaloneSTFT as Twin
    import final_comparison as fc
    from scipy.signal import ShortTimeFFT

    os.makedirs(OUT, exist_ok=True)

    # 1. examples/basic_usage.rs golden case (:13-40, :43-78, :150-160)
    np.random.seed(42)
    x30 = np.round(np.random.randn(30), 8)      # the example hard-codes the 8-digit roundings
    rec = run_twin(Twin, hann_periodic(15), 8, 1.0, "onesided", x30)
    rec["scipy_3x5"] = np.array([
        [2.16686938 + 0j, 0.41507935 + 0j, -4.76289176 + 0j, -2.7497881 + 0j, -0.35620841 + 0j],
        [0.92149425 - 1.42790582j, 1.45518835 + 2.23154074j, -3.19990722 - 1.40842161j,
         -2.04897287 + 0.55131399j, -0.01891355 - 0.33500474j],
        [-0.30760012 - 0.98739154j, 0.71763657 + 0.78718444j, -0.05623971 - 0.92249341j,
         -1.36664598 - 0.21423326j, 0.28685536 - 0.05804994j]])
    save("basic_usage", [rec])

    # 2. tests/integration_tests.rs: symmetric Hann(16), hop 4, fs 1000, n 64, 3 signals x 4 modes
    i = np.arange(64)
    tt = i / 1000.0
    sigs = {"impulse": (i == 32).astype(float),
            "sine_wave": np.sin(2 * np.pi * 5.0 * i / 1000.0),
            "chirp": np.sin(2 * np.pi * (5.0 + 10.0 * tt) * tt)}
    recs = []
    for name, x in sigs.items():
        for mode in ("onesided", "twosided", "centered", "onesided2X"):
            r = run_twin(Twin, hann_symmetric(16), 4, 1000.0, mode, x)
            r["name"] = name
            recs.append(r)
    save("integration", recs)

    # 3. python/final_comparison.py: the 8 cross-implementation configs (:250-260), fs = 1.0
    signals, windows = fc.create_test_data()
    cfgs = [("reference", "hann_15", 8, "onesided"), ("sine_5hz", "hann_16", 4, "onesided"),
            ("sine_5hz", "hann_16", 4, "twosided"), ("sine_5hz", "hann_16", 4, "centered"),
            ("chirp", "hann_32", 8, "onesided"), ("impulse", "hann_16", 4, "onesided"),
            ("noise", "hamming_16", 8, "onesided"), ("multi_freq", "hann_16", 4, "onesided")]
    recs = []
    for s, w, hop, mode in cfgs:
        r = run_twin(Twin, windows[w], hop, 1.0, mode, signals[s])
        r["name"] = f"{s}/{w}/{hop}/{mode}"
        recs.append(r)
    save("final_comparison", recs)

    # 4. randomized sweep: odd sizes, mfft > m_num, phase shifts, k_offset, p0/p1, k0/k1
    rng = np.random.default_rng(20260119)
    recs = []
    while len(recs) < 160:
        m = int(rng.integers(3, 41))
        hop = int(rng.integers(1, m + 1))
        mfft = int(rng.choice([m, m, m + int(rng.integers(1, 9)), 1 << int(np.ceil(np.log2(m)))]))
        mfft = max(mfft, m)
        mode = str(rng.choice(["onesided", "twosided", "centered", "onesided2X"]))
        ps = int(rng.integers(-mfft + 1, mfft)) if rng.random() < 0.4 else 0
        kind = rng.integers(0, 4)
        if kind == 0:
            win = hann_periodic(m) + 0.05
        elif kind == 1:
            win = hann_symmetric(max(m, 2))[:m] if m > 1 else np.ones(1)
        elif kind == 2:
            win = rng.uniform(0.2, 1.0, m)
        else:
            win = hann_periodic(m)
        n = int(rng.integers(m - m // 2 + 1, 6 * m + 20))
        x = rng.standard_normal(n)
        kw = {}
        variant = rng.integers(0, 4)
        if variant == 1:
            kw["k_offset"] = int(rng.integers(-5, 6))
        tw0 = Twin(win=win, hop=hop, fs=2.0, fft_mode=mode, mfft=mfft, phase_shift=ps)
        if variant == 2:
            pmin, pmax = tw0.p_min, tw0.p_max(n)
            kw["p0"] = int(rng.integers(pmin, pmax))
            kw["p1"] = int(rng.integers(kw["p0"] + 1, pmax + 3))
        r = run_twin(Twin, win, hop, 2.0, mode, x, mfft=mfft, phase_shift=ps, **kw)
        if variant == 3 and tw0.invertible:
            # non-default (k0, k1); k0 >= 0 (the twin mis-indexes S for q < p_min, see oracle)
            S = r["S"]
            kmax_d = (S.shape[1] + tw0.p_min - 1) * hop + m - m // 2
            k0 = int(rng.integers(0, max(1, kmax_d // 3)))
            k1 = int(rng.integers(k0 + (m - m // 2), kmax_d + 1))
            r.update(k0=k0, k1=k1, y=np.real(tw0.istft(S, k0, k1)))
        recs.append(r)
    save("sweep", recs)

    # 5. BASELINE shapes, down-scaled in length only (window / hop / mode as in BASELINE.json)
    rng = np.random.default_rng(7)
    recs = []
    n1 = 6000
    x = np.sin(2 * np.pi * 440 * np.arange(n1) / 44100.0) + 0.1 * rng.standard_normal(n1)
    recs.append(dict(run_twin(Twin, hann_periodic(2048), 512, 44100.0, "onesided", x), name="c1"))
    recs.append(dict(run_twin(Twin, hann_periodic(512), 128, 16000.0, "onesided",
                              rng.uniform(-1, 1, 4000)), name="c2"))
    recs.append(dict(run_twin(Twin, gaussian_periodic(1024, 128.0), 256, 48000.0, "twosided",
                              rng.uniform(-1, 1, 3000)), name="c4"))
    save("baseline_shapes", recs)

    # 6. scipy-only features: scale_to + spectrogram + onesided2X(psd)  (scipy :993, :1261)
    recs = []
    for (m, hop, mfft, mode, scale) in [(16, 4, 16, "onesided", "magnitude"),
                                       (16, 4, 16, "onesided", "psd"),
                                       (16, 4, 32, "onesided2X", "psd"),
                                       (15, 5, 15, "onesided2X", "magnitude"),
                                       (24, 6, 24, "twosided", "psd"),
                                       (24, 8, 32, "centered", "magnitude"),
                                       (64, 4, 64, "onesided2X", "psd")]:
        win = hann_periodic(m) + 0.01
        x = rng.standard_normal(5 * m + 3)
        st = ShortTimeFFT(win, hop, fs=8.0, fft_mode=mode, mfft=mfft, scale_to=scale)
        S = st.stft(x)
        recs.append(dict(win=win, hop=hop, fs=8.0, mode=mode, mfft=mfft, scale=scale, x=x, S=S,
                         spec=st.spectrogram(x), y=np.real(st.istft(S)), k1=st.k_max(len(x)),
                         win_scaled=st.win, dual=st.dual_win, p_min=st.p_min,
                         p_max=st.p_max(len(x))))
    save("scipy_scaling", recs)

    # 7. the twin's border modes (python/standalone_stft.py:309-336; not in the Rust crate)
    recs = []
    prng = np.random.default_rng(7)
    for (m, hop, mfft, mode, n, koff) in [(16, 4, 16, "onesided", 64, 0), (15, 5, 20, "twosided", 47, 0),
                                          (32, 8, 32, "centered", 100, 3), (24, 6, 32, "onesided2X", 90, -2),
                                          (64, 16, 64, "onesided", 300, 0), (12, 12, 12, "onesided", 40, 0)]:
        win = hann_periodic(m) + 0.05
        x = prng.standard_normal(n)
        for padding in ("zeros", "edge", "even", "odd"):
            tw = Twin(win=win, hop=hop, fs=2.0, fft_mode=mode, mfft=mfft)
            S = tw.stft(x, k_offset=koff, padding=padding)
            recs.append(dict(win=win, hop=hop, fs=2.0, mode=mode, mfft=mfft, x=x, k_offset=koff,
                             padding=padding, S=S, p_min=tw.p_min, p_max=tw.p_max(n)))
    save("padding", recs)

    # 8. per-slice detrending (python/standalone_stft.py:38-50, 345-371; not in the Rust crate)
    recs = []
    prng = np.random.default_rng(8)
    for (m, hop, mfft, mode, n) in [(16, 4, 16, "onesided", 64), (15, 5, 20, "twosided", 47),
                                    (32, 8, 32, "centered", 100), (24, 6, 32, "onesided2X", 90),
                                    (64, 16, 64, "onesided", 300)]:
        win = hann_periodic(m) + 0.05
        x = prng.standard_normal(n) + 0.02 * np.arange(n) + 1.5
        for detr in ("constant", "linear"):
            for padding in ("zeros", "even"):
                tw = Twin(win=win, hop=hop, fs=2.0, fft_mode=mode, mfft=mfft)
                S = tw.stft_detrend(x, detr, padding=padding)
                recs.append(dict(win=win, hop=hop, fs=2.0, mode=mode, mfft=mfft, x=x, detr=detr,
                                 padding=padding, S=S))
    save("detrend", recs)


if __name__ == "__main__":
    main()
