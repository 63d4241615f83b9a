"""CPU ORACLE for the veils STFT hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This module is a numpy restatement of the algorithm of ultrasaw/veils'
``StandaloneSTFT`` (Rust ``src/lib.rs``; the crate's own Python twin is
``python/standalone_stft.py``).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only
as the checker or the CPU baseline.  The product (``veils_b200``) never imports it.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the reference's Python twin
from ``/root/reference/python`` (in the build container) and commits its outputs under
``tests/golden/``; ``tests/test_oracle.py`` holds this restatement to those vectors
(<= 1e-13 relative), to the golden scipy coefficients printed in
``examples/basic_usage.rs:13-40`` and to the perfect-reconstruction thresholds of
``tests/integration_tests.rs``.  ``scale_to`` / ``spectrogram`` do not exist in the
reference; their spec is scipy.signal.ShortTimeFFT (1.18.1), which the tests use as
the second oracle.

The FFT arithmetic of the reference lives in a third-party crate that is NOT vendored:
rustfft 6.4.1 (``Cargo.lock:77-89``; call sites ``src/lib.rs:209-211,356,362,372,390,
418,425,450,484``).  Its published contract is the unnormalised DFT
X[k] = sum_j x[j] exp(-/+ 2 pi i jk / M) for any length M; ``numpy.fft`` (pocketfft)
computes the same sums and stands in for it here.

Where the Rust code and its twin disagree this file follows:
  * pre_padding early-exit   -> twin / scipy (Rust mis-translates a negative slice)
  * p1 upper bound           -> none (Rust and twin both allow p1 > p_max)
  * istft q0 for negative k0 -> floor division (twin); ``k_min <= k0`` is enforced
  * istft output             -> real part only (Rust returns Vec<f64>, lib.rs:903-907)
  * invertibility epsilon    -> f64::EPSILON * max(DD) (Rust, lib.rs:313-318)
  * onesided2X w/o scaling   -> legal, factor 2 (Rust lib.rs:387-408); sqrt(2) iff psd
  * f()                      -> Rust for onesided (+fs/2 at Nyquist), scipy for centered
"""
from __future__ import annotations

import numpy as np

FFT_MODES = ("twosided", "centered", "onesided", "onesided2X")


def calc_dual_canonical_window(win: np.ndarray, hop: int) -> np.ndarray:
    """Canonical dual window.  Follows src/lib.rs:284-322 (twin :13-35).

    dual[i] = win[i] / sum_j win[i + j*hop]^2 over all in-range j.
    The accumulation order mirrors the reference loops (k = hop, 2hop, ...; first the
    "shift right" then the "shift left" contribution).
    """
    n = len(win)
    if hop > n:
        raise ValueError(
            f"hop={hop} is larger than window length of {n} => STFT not invertible!")
    w2 = win.real ** 2 + win.imag ** 2
    dd = w2.copy()
    for k in range(hop, n, hop):
        dd[k:] += w2[:-k]
        dd[:-k] += w2[k:]
    if not np.all(dd >= np.finfo(np.float64).eps * dd.max()):
        raise ValueError("Short-time Fourier Transform not invertible!")
    return win / dd


class OracleSTFT:
    """numpy restatement of veils::StandaloneSTFT (src/lib.rs:116-224)."""

    def __init__(self, win, hop, fs, fft_mode="onesided", mfft=None, dual_win=None,
                 scale_to=None, phase_shift=0):
        win = np.asarray(win, dtype=np.float64)
        # src/lib.rs:176-186
        if win.ndim != 1 or win.size == 0:
            raise ValueError("Parameter win must be non-empty!")
        if not np.all(np.isfinite(win)):
            raise ValueError("Parameter win must have finite entries!")
        if not hop >= 1:
            raise ValueError(f"Parameter hop={hop} is not >= 1!")
        if fft_mode not in FFT_MODES:                      # src/lib.rs:66-78
            raise ValueError(f"Unknown fft_mode: {fft_mode}")
        self.win = win.copy()
        self.hop = int(hop)
        self.fs = float(fs)
        self.fft_mode = fft_mode
        self.mfft = len(win) if mfft is None else int(mfft)   # src/lib.rs:193
        self.phase_shift = 0 if phase_shift is None else int(phase_shift)
        # scipy _short_time_fft.py:957-965 and :1080-1082 (the reference omits both checks)
        if self.mfft < len(win):
            raise ValueError(f"Attribute mfft={self.mfft} needs to be at least the "
                             f"window length m_num={len(win)}!")
        if not -self.mfft < self.phase_shift < self.mfft:
            raise ValueError(f"-mfft < phase_shift < mfft does not hold for "
                             f"mfft={self.mfft}, phase_shift={self.phase_shift}!")
        self._dual = None
        if dual_win is not None:                            # src/lib.rs:197-207
            dual_win = np.asarray(dual_win, dtype=np.float64)
            if dual_win.shape != win.shape:
                raise ValueError("dual_win.len() must equal win.len()!")
            if not np.all(np.isfinite(dual_win)):
                raise ValueError("Parameter dual_win must have finite entries!")
            self._dual = dual_win.copy()
        self.scaling = None
        if scale_to is not None:
            self.scale_to(scale_to)

    # -- scipy-only: scale_to (scipy _short_time_fft.py:993-1039, fac_* :1553-1595) ------
    def scale_to(self, scaling):
        if scaling not in ("magnitude", "psd"):
            raise ValueError(f"scaling={scaling!r} not in ['magnitude', 'psd']!")
        if self.scaling == scaling:
            return
        s_fac = self.fac_psd if scaling == "psd" else self.fac_magnitude
        self.win = self.win * s_fac
        if self._dual is not None:
            self._dual = self._dual / s_fac
        self.scaling = scaling

    @property
    def fac_magnitude(self):
        return 1.0 / abs(np.sum(self.win))

    @property
    def fac_psd(self):
        return 1.0 / np.sqrt(np.sum(self.win.real ** 2 + self.win.imag ** 2) / (1.0 / self.fs))

    # -- accessors (src/lib.rs:227-281) ---------------------------------------------------
    @property
    def m_num(self):
        return len(self.win)

    @property
    def m_num_mid(self):
        return self.m_num // 2

    @property
    def onesided_fft(self):
        return self.fft_mode in ("onesided", "onesided2X")

    @property
    def f_pts(self):
        return self.mfft // 2 + 1 if self.onesided_fft else self.mfft

    @property
    def T(self):
        return 1.0 / self.fs

    @property
    def delta_t(self):
        return self.hop * self.T

    @property
    def delta_f(self):
        return 1.0 / (self.mfft * self.T)

    @property
    def dual_win(self):
        """src/lib.rs:324-329 (lazy canonical dual)."""
        if self._dual is None:
            self._dual = calc_dual_canonical_window(self.win, self.hop)
        return self._dual

    @property
    def invertible(self):
        try:
            _ = self.dual_win
            return True
        except ValueError:
            return False

    @property
    def _p_s(self):
        """Roll amount; src/lib.rs:343 with a true (non-negative) modulus (twin :236)."""
        return (self.phase_shift + self.m_num_mid) % self.m_num

    @property
    def _fac2x(self):
        """onesided2X factor: 2, or sqrt(2) under psd scaling (scipy :2169, twin :249)."""
        return np.sqrt(2.0) if self.scaling == "psd" else 2.0

    # -- padding / slice ranges (src/lib.rs:524-609, corrected by twin :153-208) ---------
    def _pre_padding(self):
        w2 = self.win.real ** 2 + self.win.imag ** 2
        n0 = -self.m_num_mid
        for p_, n_ in enumerate(range(n0, n0 - self.m_num - 1, -self.hop)):
            n_next = n_ - self.hop
            if n_next + self.m_num <= 0 or np.all(w2[n_next:] == 0):
                return n_, -p_
        return n0 - self.m_num, -(self.m_num // self.hop + 1)

    def _post_padding(self, n):
        m2p = self.m_num - self.m_num_mid
        if not n >= m2p:
            raise ValueError(f"Parameter n must be >= ceil(m_num/2) = {m2p}!")
        w2 = self.win.real ** 2 + self.win.imag ** 2
        q1 = n // self.hop
        k1 = q1 * self.hop - self.m_num_mid
        for q_, k_ in enumerate(range(k1, n + self.m_num, self.hop), start=q1):
            n_next = k_ + self.hop
            if n_next >= n or np.all(w2[:n - n_next] == 0):
                return k_ + self.m_num, q_ + 1
        return n + self.m_num, (n + self.m_num_mid) // self.hop + 1

    @property
    def p_min(self):
        return self._pre_padding()[1]

    @property
    def k_min(self):
        return self._pre_padding()[0]

    def p_max(self, n):
        return self._post_padding(n)[1]

    def k_max(self, n):
        return self._post_padding(n)[0]

    def p_num(self, n):
        return self.p_max(n) - self.p_min

    def p_range(self, n, p0=None, p1=None):
        """src/lib.rs:595-609 -- note: no upper bound on p1."""
        p0 = self.p_min if p0 is None else p0
        p1 = self.p_max(n) if p1 is None else p1
        if not self.p_min <= p0 < p1:
            raise ValueError(f"Invalid slice range: p0={p0}, p1={p1}")
        return p0, p1

    # -- forward: x_slices + window + fft_func + transpose (src/lib.rs:612-749) -----------
    def _frames(self, x, p0, p1, k_offset, padding="zeros"):
        """All frames p0..p1 as a (P, m_num) array, zeros outside [0, n) (lib.rs:612-652).
        padding 'edge' | 'even' | 'odd': the twin's border values (python/standalone_stft.py:309-336;
        absent from the Rust crate): edge x[0] / x[n-1]; even x[min(-s, n-1)] left of 0 and
        x[clamp(2n-2-s)] right of n-1; odd the negative of even."""
        n = len(x)
        starts = np.arange(p0, p1, dtype=np.int64) * self.hop + k_offset - self.m_num_mid
        idx = starts[:, None] + np.arange(self.m_num, dtype=np.int64)[None, :]
        ok = (idx >= 0) & (idx < n)
        inside = x[np.clip(idx, 0, n - 1)]
        if padding == "zeros":
            return np.where(ok, inside, 0.0)
        if padding == "edge":
            return inside
        if padding not in ("even", "odd"):
            raise ValueError(f"Parameter padding='{padding}' not in ['zeros', 'edge', 'even', 'odd']!")
        mirror = np.where(idx < 0, np.minimum(-idx, n - 1), np.clip(2 * n - 2 - idx, 0, n - 1))
        pad = x[np.clip(mirror, 0, n - 1)]
        return np.where(ok, inside, pad if padding == "even" else -pad)

    def _fft_func(self, y):
        """Rows of y (P, m_num) -> (P, f_pts).  src/lib.rs:336-411."""
        M = self.mfft
        if y.shape[-1] < M:
            y = np.concatenate([y, np.zeros(y.shape[:-1] + (M - y.shape[-1],), y.dtype)], -1)
        y = np.roll(y, -self._p_s, axis=-1)
        if self.fft_mode == "twosided":
            return np.fft.fft(y, axis=-1)
        if self.fft_mode == "centered":
            return np.fft.fftshift(np.fft.fft(y, axis=-1), axes=-1)
        X = np.fft.rfft(y.real, axis=-1)          # lib.rs:367-386: Im(DC)=Im(Nyq)=0
        if self.fft_mode == "onesided2X":
            X = X.copy()
            X[..., 1:(-1 if M % 2 == 0 else None)] *= self._fac2x
        return X

    @staticmethod
    def _detrend(frames, detr):
        """Per-slice detrending of the twin (python/standalone_stft.py:38-50): 'constant' removes the
        mean, 'linear' the least-squares line a*t + b, t = 0..m_num-1."""
        if detr is None:
            return frames
        if detr == "constant":
            return frames - frames.mean(axis=1, keepdims=True)
        if detr != "linear":
            raise ValueError(f"Unknown detrend type: {detr}")
        t = np.arange(frames.shape[1], dtype=np.float64)
        tc = t - t.mean()
        m = frames.mean(axis=1, keepdims=True)
        a = ((frames - m) * tc[None, :]).sum(axis=1, keepdims=True) / np.sum(tc ** 2)
        b = m - a * t.mean()
        return frames - (a * t[None, :] + b)

    def stft_detrend(self, x, detr, p0=None, p1=None, k_offset=0, padding="zeros"):
        """python/standalone_stft.py:345-371."""
        return self.stft(x, p0, p1, k_offset, padding, detr)

    def stft(self, x, p0=None, p1=None, k_offset=0, padding="zeros", detr=None):
        """src/lib.rs:689-749.  Returns S[f_pts][P] complex128."""
        x = np.asarray(x, dtype=np.float64)
        if self.onesided_fft and np.any(np.isnan(x)):
            raise ValueError("Complex-valued input not allowed for one-sided FFT modes")
        n = len(x)
        m2p = self.m_num - self.m_num_mid
        if n < m2p:
            raise ValueError(f"Input length {n} must be >= ceil(m_num/2) = {m2p}")
        p0, p1 = self.p_range(n, p0, p1)
        S = np.empty((self.f_pts, p1 - p0), dtype=np.complex128)
        chunk = max(1, (1 << 22) // max(self.mfft, 1))        # bound temporaries
        for a in range(p0, p1, chunk):
            b = min(p1, a + chunk)
            y = self._detrend(self._frames(x, a, b, k_offset, padding), detr) * self.win[None, :]
            S[:, a - p0:b - p0] = self._fft_func(y).T
        return S

    def spectrogram(self, x, p0=None, p1=None, k_offset=0):
        """scipy _short_time_fft.py:1398-1400: Sx.real**2 + Sx.imag**2."""
        S = self.stft(x, p0, p1, k_offset)
        return S.real ** 2 + S.imag ** 2

    # -- inverse: ifft_func + dual window + overlap-add (src/lib.rs:414-505, 792-914) -----
    def _ifft_func(self, X):
        """Rows of X (Q, f_pts) -> (Q, m_num) complex.  src/lib.rs:414-505."""
        M = self.mfft
        if self.fft_mode == "twosided":
            x = np.fft.ifft(X, axis=-1)
        elif self.fft_mode == "centered":
            x = np.fft.ifft(np.fft.ifftshift(X, axes=-1), axis=-1)
        else:
            Xc = X
            if self.fft_mode == "onesided2X":
                Xc = X.copy()
                Xc[..., 1:(-1 if M % 2 == 0 else None)] /= self._fac2x
            x = np.fft.irfft(Xc, n=M, axis=-1)
        return np.roll(x, self._p_s, axis=-1)[..., :self.m_num]

    def istft_range(self, P, k0=None, k1=None):
        """Validation + (k0, k1, q0, q1) of src/lib.rs:798-847 for P time slices."""
        n_min = self.m_num - self.m_num_mid
        q_num = self.p_num(n_min)
        if P < q_num:
            raise ValueError(f"STFT time dimension {P} needs to have at least {q_num} slices")
        q_max = P + self.p_min
        k_max = (q_max - 1) * self.hop + self.m_num - self.m_num_mid
        k0 = 0 if k0 is None else k0
        k1 = k_max if k1 is None else k1
        if not self.k_min <= k0 < k1:
            raise ValueError(f"k0={k0} must be < k1={k1} and >= k_min={self.k_min}")
        if k1 - k0 < n_min:
            raise ValueError(f"Output length {k1 - k0} has to be at least {n_min}")
        q0 = k0 // self.hop + self.p_min if k0 >= 0 else k0 // self.hop
        q1 = min(self.p_max(k1), q_max)
        return k0, k1, q0, q1

    def istft(self, S, k0=None, k1=None):
        """src/lib.rs:792-914.  S is [f_pts][P]; returns real float64 of length k1-k0."""
        S = np.asarray(S)
        if S.ndim != 2 or S.shape[0] == 0:
            raise ValueError("STFT data cannot be empty")
        if S.shape[0] != self.f_pts:
            raise ValueError(f"STFT frequency dimension {S.shape[0]} must equal {self.f_pts}")
        P = S.shape[1]
        k0, k1, q0, q1 = self.istft_range(P, k0, k1)
        dual = self.dual_win
        num_pts = k1 - k0
        x = np.zeros(num_pts, dtype=np.float64)
        p_min = self.p_min
        chunk = max(1, (1 << 22) // max(self.mfft, 1))
        for qa in range(q0, q1, chunk):
            qb = min(q1, qa + chunk)
            cols = np.arange(qa, qb) - p_min
            valid = (cols >= 0) & (cols < P)
            xs_all = np.zeros((qb - qa, self.m_num))
            if valid.any():
                xs_all[valid] = (self._ifft_func(S[:, cols[valid]].T) * dual[None, :]).real
            for q in range(qa, qb):                         # increasing q: lib.rs:853
                if not valid[q - qa]:
                    continue
                i0 = q * self.hop - self.m_num_mid
                i1 = min(i0 + self.m_num, num_pts + k0)
                j0, j1 = 0, i1 - i0
                if i0 < k0:
                    j0 += k0 - i0
                    i0 = k0
                if i0 < i1 and j0 < j1:
                    x[i0 - k0:i1 - k0] += xs_all[q - qa, j0:j1]
        return x

    # -- axes (src/lib.rs:916-949) --------------------------------------------------------
    def t(self, n, p0=None, p1=None, k_offset=0):
        p0, p1 = self.p_range(n, p0, p1)
        return (np.arange(p0, p1) * self.hop + k_offset) * self.T

    def f(self):
        M = self.mfft
        if self.onesided_fft:
            return np.arange(self.f_pts) / (M * self.T)
        i = np.arange(M)
        fr = i / (M * self.T)
        fr = np.where(i > M // 2, fr - 1.0 / self.T, fr)    # lib.rs:936-945
        if self.fft_mode == "centered":
            # scipy fftshift(fftfreq): differs from the Rust list only in ordering and in
            # the sign convention of the Nyquist bin for even M.
            fr = np.fft.fftshift(np.fft.fftfreq(M, self.T))
        return fr
