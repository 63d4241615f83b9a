"""The host-pointer flavour of the C ABI (veils_stft_host / veils_istft_host / veils_spectrogram_host,
the drop-in for the reference's Vec API, src/lib.rs:689-749 and :792-914): chunked two-stream
staging must give exactly the device path's results, reject NaN input like the reference
(src/lib.rs:696-698) and survive concurrent callers on one plan."""
import threading

import numpy as np
import pytest

from veils_b200 import StandaloneSTFT

pytestmark = pytest.mark.gpu


def _hann(n):
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)


@pytest.mark.parametrize("precision,batch,n,m,hop", [("f32", 40, 160000, 512, 128),    # several 32 MB chunks
                                                      ("f64", 5, 300000, 512, 128),
                                                      ("f32", 3, 5000, 96, 17),          # generic DFT path
                                                      ("f64", 1, 2000, 64, 16)])
def test_host_matches_device(precision, batch, n, m, hop):
    import torch
    s = StandaloneSTFT(_hann(m), hop, 16000.0, fft_mode="onesided", precision=precision)
    rng = np.random.default_rng(5)
    x = rng.standard_normal((batch, n)).astype(np.float32 if precision == "f32" else np.float64)
    S_h = s.stft(x)                                             # numpy in -> veils_stft_host
    xd = torch.from_numpy(x).cuda()
    S_d = s.stft(xd)                                            # device tensor -> veils_stft_dev
    assert np.array_equal(S_h, S_d.cpu().numpy())
    y_h = s.istft(S_h)
    y_d = s.istft(S_d)
    assert np.array_equal(y_h, y_d.cpu().numpy())
    P_h = s.spectrogram(x)
    assert np.array_equal(P_h, s.spectrogram(xd).cpu().numpy())
    assert np.max(np.abs(y_h[:, :n] - x)) < (1e-5 if precision == "f32" else 1e-12)


def test_host_rejects_nan_then_recovers():
    s = StandaloneSTFT(_hann(256), 64, 1.0, fft_mode="onesided", precision="f32")
    x = np.random.default_rng(1).standard_normal((300, 40000)).astype(np.float32)
    bad = x.copy()
    bad[271, 39999] = np.nan                                    # last chunk, last sample
    with pytest.raises(ValueError, match="Complex-valued input not allowed for one-sided FFT modes"):
        s.stft(bad)
    S = s.stft(x)                                               # the flag is reset per call
    assert np.isfinite(S).all()
    two = StandaloneSTFT(_hann(256), 64, 1.0, fft_mode="twosided", precision="f32")
    assert two.stft(bad[271:272]).shape[0] == 1                 # two-sided modes do not screen (src/lib.rs:696)


def test_host_calls_from_threads():
    s = StandaloneSTFT(_hann(512), 128, 1.0, fft_mode="onesided", precision="f32")
    rng = np.random.default_rng(2)
    xs = [rng.standard_normal((24, 160000)).astype(np.float32) for _ in range(4)]
    want = [s.stft(x) for x in xs]
    got = [None] * 4

    def work(i):
        got[i] = s.stft(xs[i])

    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
