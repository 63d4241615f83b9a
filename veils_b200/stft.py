"""Host-side mirror of the reference's `StandaloneSTFT` (Rust src/lib.rs:116-955; Python twin
python/standalone_stft.py:53-444) on top of the C ABI of libveils_cuda.so.

Same constructor arguments, methods, defaults and error texts as the reference, plus the two
scipy.signal.ShortTimeFFT features BASELINE.json's north_star asks for (`scale_to`,
`spectrogram`).  All arithmetic runs in the CUDA kernels; this file only marshals arguments.
numpy inputs go through the host-pointer entry points, torch CUDA tensors through the
device-pointer entry points (no copies, current torch stream)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, opt_i64

_MODES = ("twosided", "centered", "onesided", "onesided2X")
_SCALES = (None, "magnitude", "psd")


class VeilsError(ValueError):
    """An `Err(String)` of the reference API (status code in `.code`)."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _check(rc):
    if rc != _lib.OK:
        msg = _lib.last_error()
        if rc == _lib.ERR_CUDA:
            raise RuntimeError(f"veils_b200 CUDA failure: {msg}")
        raise VeilsError(rc, msg)


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class StandaloneSTFT:
    """Drop-in for veils::StandaloneSTFT backed by the sm_100a kernels."""

    NAN_TEXT = "Complex-valued input not allowed for one-sided FFT modes"      # src/lib.rs:697

    def __init__(self, win, hop, fs, *, fft_mode="onesided", mfft=None, dual_win=None,
                 scale_to=None, phase_shift=0, precision="f64", device=0, nan_check="deferred"):
        """phase_shift=None (scipy): no roll of the zero-padded slice, i.e. phase_shift = -m_num_mid.
        nan_check (device tensors, one-sided modes; src/lib.rs:696-698): 'deferred' -- the screen runs behind
        the transform and the error is raised by the next call on this object or by check_input();
        'sync' -- stft() waits for it and raises like the reference; 'off'.  Host arrays always raise."""
        L = lib()
        win = np.ascontiguousarray(np.asarray(win, dtype=np.float64).ravel())
        dual = None if dual_win is None else np.ascontiguousarray(
            np.asarray(dual_win, dtype=np.float64).ravel())
        if dual is not None and dual.shape != win.shape:
            raise VeilsError(_lib.ERR_INVALID, "dual_win.len() must equal win.len()!")
        if not isinstance(hop, (int, np.integer)):
            raise VeilsError(_lib.ERR_INVALID, f"Parameter hop={hop} is not >= 1!")
        if precision not in ("f32", "f64"):
            raise ValueError("precision must be 'f32' or 'f64'")
        d = _lib.PlanDesc()
        d.win = _dptr(win) if win.size else None
        d.m_num = win.size
        d.hop = int(hop)
        d.fs = float(fs)
        d.fft_mode = None if fft_mode is None else str(fft_mode).encode()
        d.mfft = 0 if mfft is None else int(mfft)
        d.dual_win = None if dual is None else _dptr(dual)
        d.scale_to = None if scale_to is None else str(scale_to).encode()
        self._phase_none = phase_shift is None
        d.phase_shift = -(win.size // 2) if phase_shift is None else int(phase_shift)
        if nan_check not in ("deferred", "sync", "off"):
            raise ValueError("nan_check must be 'deferred', 'sync' or 'off'")
        self.nan_check = nan_check
        self._nanflag = None
        d.precision = _lib.F32 if precision == "f32" else _lib.F64
        d.device = int(device)
        if mfft is not None and int(mfft) <= 0:
            raise VeilsError(_lib.ERR_INVALID, f"Parameter mfft={mfft} is not >= 1!")
        h = C.c_void_p()
        self._h = None
        _check(L.veils_plan_create(C.byref(d), C.byref(h)))
        self._h = h
        self._L = L
        self.precision = precision
        self.pitch_align = 16 if precision == "f32" else 8     # complex elements per 128-byte line
        self._rdt = np.float32 if precision == "f32" else np.float64
        self._cdt = np.complex64 if precision == "f32" else np.complex128

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.veils_plan_destroy(h)

    # ---- accessors (src/lib.rs:227-281) --------------------------------------------------
    @property
    def win(self):
        out = np.empty(self.m_num)
        _check(self._L.veils_win(self._h, _dptr(out)))
        return out

    @property
    def dual_win(self):
        out = np.empty(self.m_num)
        _check(self._L.veils_dual_win(self._h, _dptr(out)))
        return out

    @property
    def invertible(self):
        return bool(self._L.veils_invertible(self._h))

    hop = property(lambda s: s._L.veils_hop(s._h))
    fs = property(lambda s: s._L.veils_fs(s._h))
    mfft = property(lambda s: s._L.veils_mfft(s._h))
    m_num = property(lambda s: s._L.veils_m_num(s._h))
    m_num_mid = property(lambda s: s._L.veils_m_num_mid(s._h))
    f_pts = property(lambda s: s._L.veils_f_pts(s._h))
    phase_shift = property(lambda s: None if s._phase_none else s._L.veils_phase_shift(s._h))
    delta_t = property(lambda s: s._L.veils_delta_t(s._h))
    delta_f = property(lambda s: s._L.veils_delta_f(s._h))
    T = property(lambda s: 1.0 / s.fs)
    sampling_period = T
    fft_mode = property(lambda s: _MODES[s._L.veils_fft_mode(s._h)])
    onesided_fft = property(lambda s: bool(s._L.veils_onesided_fft(s._h)))
    scaling = property(lambda s: _SCALES[s._L.veils_scaling(s._h)])
    p_min = property(lambda s: s._L.veils_p_min(s._h))
    k_min = property(lambda s: s._L.veils_k_min(s._h))
    device = property(lambda s: s._L.veils_device(s._h))

    def debug_ifft_func(self, X):
        """src/lib.rs:952-954: ifft_func of one slice (f_pts complex -> m_num complex), host-side f64."""
        X = np.ascontiguousarray(np.asarray(X, dtype=np.complex128).ravel())
        if X.size != self.f_pts:
            raise VeilsError(_lib.ERR_INVALID, f"expected {self.f_pts} bins")
        out = np.empty(self.m_num, dtype=np.complex128)
        _check(self._L.veils_debug_ifft_func(self._h, X.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)),
                                             out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def kernel_name(self, inverse=False):
        return self._L.veils_plan_kernel_name(self._h, int(inverse)).decode()

    def _i64(self, fn, *args):
        out = C.c_int64()
        _check(fn(self._h, *args, C.byref(out)))
        return out.value

    def p_max(self, n):
        return self._i64(self._L.veils_p_max, int(n))

    def k_max(self, n):
        return self._i64(self._L.veils_k_max, int(n))

    def p_num(self, n):
        return self._i64(self._L.veils_p_num, int(n))

    def p_range(self, n, p0=None, p1=None):
        a, b = C.c_int64(), C.c_int64()
        _check(self._L.veils_p_range(self._h, int(n), opt_i64(p0), opt_i64(p1), C.byref(a), C.byref(b)))
        return a.value, b.value

    def istft_range(self, P, k0=None, k1=None):
        a, b = C.c_int64(), C.c_int64()
        _check(self._L.veils_istft_range(self._h, int(P), opt_i64(k0), opt_i64(k1), C.byref(a), C.byref(b)))
        return a.value, b.value

    def t(self, n, p0=None, p1=None, k_offset=0):
        a, b = self.p_range(n, p0, p1)
        out = np.empty(b - a)
        _check(self._L.veils_t(self._h, int(n), opt_i64(p0), opt_i64(p1), int(k_offset), _dptr(out)))
        return out

    def f(self):
        out = np.empty(self.f_pts)
        _check(self._L.veils_f(self._h, _dptr(out)))
        return out

    # ---- scipy.signal.ShortTimeFFT axis helpers: host integer logic behind the C ABI -------
    @property
    def lower_border_end(self):
        """(k, p): first sample / slice not affected by the left border (scipy `lower_border_end`)."""
        k, p = C.c_int64(), C.c_int64()
        _check(self._L.veils_lower_border_end(self._h, C.byref(k), C.byref(p)))
        return k.value, p.value

    def upper_border_begin(self, n):
        """(k, p): first sample / slice affected by the right border of an n-sample signal
        (scipy `upper_border_begin`)."""
        k, p = C.c_int64(), C.c_int64()
        _check(self._L.veils_upper_border_begin(self._h, int(n), C.byref(k), C.byref(p)))
        return k.value, p.value

    def extent(self, n, axes_seq="tf", center_bins=False):
        """(t0, t1, f0, f1) of stft(x) for plotting (scipy `extent`)."""
        if axes_seq not in ("tf", "ft"):
            raise VeilsError(_lib.ERR_INVALID, f"Parameter axes_seq='{axes_seq}' not in ['tf', 'ft']!")
        out = np.empty(4)
        _check(self._L.veils_extent(self._h, int(n), int(axes_seq == "ft"), int(bool(center_bins)), _dptr(out)))
        return tuple(float(v) for v in out)

    def nearest_k_p(self, k, left=True):
        """Nearest sample index that is a slice position t[p] (scipy `nearest_k_p`)."""
        return int(self._L.veils_nearest_k_p(self._h, int(k), int(bool(left))))

    # ---- compute ---------------------------------------------------------------------------
    @staticmethod
    def _is_tensor(x):
        return type(x).__module__.startswith("torch")

    PADDINGS = ("zeros", "edge", "even", "odd")

    @classmethod
    def _is_complex(cls, x):
        if cls._is_tensor(x):
            return x.is_complex()
        return np.iscomplexobj(x)

    def _fwd_complex(self, x, p0, p1, k_offset, spectrogram, padding, detr):
        """The twin accepts complex x in the two-sided layouts (python/standalone_stft.py:349; the Rust crate
        is real-only).  The transform is linear: stft(x) = stft(Re x) + i stft(Im x) -- two real transforms
        and one A += i B pass (veils_complex_op_dev)."""
        if self.onesided_fft:
            raise VeilsError(_lib.ERR_INVALID, f"Complex-valued `x` not allowed for fft_mode='{self.fft_mode}'! "
                                               "Set fft_mode to 'twosided' or 'centered'.")
        if spectrogram or detr is not None:
            raise VeilsError(_lib.ERR_INVALID, "complex input: only stft() is supported")
        import torch
        tdt = torch.float32 if self.precision == "f32" else torch.float64
        was_numpy = not self._is_tensor(x)
        xt = torch.from_numpy(np.ascontiguousarray(x)).cuda(self.device) if was_numpy else x
        Sr = self._fwd_tensor(xt.real.to(tdt).contiguous(), p0, p1, k_offset, False, padding)
        Si = self._fwd_tensor(xt.imag.to(tdt).contiguous(), p0, p1, k_offset, False, padding)
        S3, T3 = (Sr[None], Si[None]) if Sr.dim() == 2 else (Sr, Si)
        batch, F, P = S3.shape
        st = torch.cuda.current_stream(Sr.device).cuda_stream
        _check(self._L.veils_complex_op_dev(self._h, 1, S3.data_ptr(), T3.data_ptr(), P, batch,
                                            S3.stride(1) if F > 1 else P, T3.stride(1) if F > 1 else P, st))
        return Sr.cpu().numpy() if was_numpy else Sr

    # ---- NaN screen of device tensors (the flag lives in page-locked host memory the kernel writes to)
    def _flag(self):
        if self._nanflag is None:
            import torch
            self._nanflag = torch.zeros(1, dtype=torch.int32, pin_memory=True)
        return self._nanflag

    def _raise_if_nan(self):
        if self._nanflag is not None and int(self._nanflag[0]) != 0:
            self._nanflag[0] = 0
            raise VeilsError(_lib.ERR_INVALID, self.NAN_TEXT)

    def check_input(self):
        """Waits for the transforms issued on this device and raises if a NaN went into a one-sided stft."""
        if self._nanflag is not None:
            import torch
            torch.cuda.synchronize(self.device)
        self._raise_if_nan()

    def _fwd(self, x, p0, p1, k_offset, spectrogram, padding="zeros", detr=None):
        if padding not in self.PADDINGS:
            raise VeilsError(_lib.ERR_INVALID, f"Parameter padding='{padding}' not in {list(self.PADDINGS)}!")
        self._raise_if_nan()
        if self._is_complex(x):
            return self._fwd_complex(x, p0, p1, k_offset, spectrogram, padding, detr)
        if self._is_tensor(x):
            return self._fwd_tensor(x, p0, p1, k_offset, spectrogram, padding, detr)
        x = np.asarray(x)
        if padding != "zeros" or detr is not None:
            # the border extension runs on the device (veils_stft_dev_padded): stage through torch
            import torch
            if np.iscomplexobj(x):
                raise VeilsError(_lib.ERR_INVALID, "Complex-valued input not allowed (real input only)")
            xt = torch.from_numpy(np.ascontiguousarray(x, dtype=self._rdt)).cuda(self.device)
            return self._fwd_tensor(xt, p0, p1, k_offset, spectrogram, padding, detr).cpu().numpy()
        if np.iscomplexobj(x):
            raise VeilsError(_lib.ERR_INVALID, "Complex-valued input not allowed (real input only)")
        squeeze = x.ndim == 1
        x2 = np.ascontiguousarray(np.atleast_2d(x), dtype=self._rdt)
        batch, n = x2.shape
        # validation of n / p-range happens in the library, but the output must be sized first
        if n < self.m_num - self.m_num_mid:
            raise VeilsError(_lib.ERR_INVALID,
                             f"Input length {n} must be >= ceil(m_num/2) = {self.m_num - self.m_num_mid}")
        a, b = self.p_range(n, p0, p1)
        out = np.empty((batch, self.f_pts, b - a), dtype=self._rdt if spectrogram else self._cdt)
        fn = self._L.veils_spectrogram_host if spectrogram else self._L.veils_stft_host
        _check(fn(self._h, x2.ctypes.data, n, batch, n, opt_i64(a), opt_i64(b), int(k_offset),
                  out.ctypes.data))
        return out[0] if squeeze else out

    def _fwd_tensor(self, x, p0, p1, k_offset, spectrogram, padding="zeros", detr=None):
        import torch
        tdt = torch.float32 if self.precision == "f32" else torch.float64
        if not x.is_cuda or x.dtype != tdt:
            raise ValueError(f"expected a CUDA tensor of dtype {tdt}")
        if x.device.index != self.device:
            raise ValueError(f"tensor on cuda:{x.device.index}, plan on cuda:{self.device}")
        squeeze = x.dim() == 1
        x2 = x.reshape(1, -1) if squeeze else x
        if x2.stride(-1) != 1:
            x2 = x2.contiguous()
        batch, n = x2.shape
        if n < self.m_num - self.m_num_mid:
            raise VeilsError(_lib.ERR_INVALID,
                             f"Input length {n} must be >= ceil(m_num/2) = {self.m_num - self.m_num_mid}")
        a, b = self.p_range(n, p0, p1)
        P = b - a
        # Row pitch: the reference keeps every frequency row in its own Vec, so any pitch >= P is
        # equally faithful.  Rows that start on a 128-byte line boundary let the tile stores write
        # whole lines (measured on B200, store-only kernel of this shape: 3.65 TB/s with 8-byte
        # aligned rows, 4.8 TB/s with 32-byte, 6.0 TB/s with 128-byte aligned rows); the result is
        # returned as the [.., :P] view of the padded buffer.
        ldp = -(-P // self.pitch_align) * self.pitch_align
        if spectrogram:
            out = torch.empty((batch, self.f_pts, ldp), dtype=tdt, device=x.device)
        else:
            out = torch.empty((batch, self.f_pts, ldp), device=x.device,
                              dtype=torch.complex64 if self.precision == "f32" else torch.complex128)
        st = torch.cuda.current_stream(x.device).cuda_stream
        if padding != "zeros" or detr is not None:
            left, n_ext = C.c_int64(0), C.c_int64(0)
            _check(self._L.veils_pad_extent(self._h, n, opt_i64(a), opt_i64(b), int(k_offset),
                                            C.byref(left), C.byref(n_ext)))
            xe = torch.empty((batch, n_ext.value), dtype=tdt, device=x.device)
            if detr is not None:
                coef = torch.empty((batch, P, 2), dtype=torch.float64, device=x.device)
                _check(self._L.veils_stft_detrend_dev(self._h, x2.data_ptr(), n, batch,
                                                      x2.stride(0) if batch > 1 else n, opt_i64(a), opt_i64(b),
                                                      int(k_offset), padding.encode(), detr.encode(), xe.data_ptr(),
                                                      n_ext.value, coef.data_ptr(), out.data_ptr(), ldp, st))
                out = out[:, :, :P]
                return out[0] if squeeze else out
            _check(self._L.veils_stft_dev_padded(self._h, x2.data_ptr(), n, batch,
                                                 x2.stride(0) if batch > 1 else n, opt_i64(a), opt_i64(b),
                                                 int(k_offset), padding.encode(), xe.data_ptr(), n_ext.value,
                                                 out.data_ptr(), ldp, st, 1 if spectrogram else 0))
            out = out[:, :, :P]
            return out[0] if squeeze else out
        if self.onesided_fft and self.nan_check != "off":
            _check(self._L.veils_stft_dev_checked(self._h, x2.data_ptr(), n, batch, x2.stride(0) if batch > 1 else n,
                                                  opt_i64(a), opt_i64(b), int(k_offset), out.data_ptr(), ldp, st,
                                                  1 if spectrogram else 0, self._flag().data_ptr()))
            if self.nan_check == "sync":
                torch.cuda.current_stream(x.device).synchronize()
                self._raise_if_nan()
        else:
            fn = self._L.veils_spectrogram_dev if spectrogram else self._L.veils_stft_dev
            _check(fn(self._h, x2.data_ptr(), n, batch, x2.stride(0) if batch > 1 else n, opt_i64(a),
                      opt_i64(b), int(k_offset), out.data_ptr(), ldp, st))
        out = out[:, :, :P]
        return out[0] if squeeze else out

    def stft(self, x, p0=None, p1=None, *, k_offset=0, padding="zeros"):
        """src/lib.rs:689-749 -> S[f_pts][P] (or [batch][f_pts][P] for 2-d input).
        padding: 'zeros' (the Rust crate's only mode) or the twin's 'edge' | 'even' | 'odd'
        (python/standalone_stft.py:309-336)."""
        return self._fwd(x, p0, p1, k_offset, False, padding)

    def stft_detrend(self, x, detr, p0=None, p1=None, *, k_offset=0, padding="zeros"):
        """The twin's stft_detrend (python/standalone_stft.py:345-371): detr = None | 'constant' |
        'linear' is removed from every slice before windowing (callables are not supported on the
        device).  Runs on the GPU as stft + a rank-2 correction (csrc/detrend.cu)."""
        if detr is None:
            return self.stft(x, p0, p1, k_offset=k_offset, padding=padding)
        if detr not in ("constant", "linear"):
            if isinstance(detr, str):
                raise VeilsError(_lib.ERR_INVALID, f"Unknown detrend type: {detr}")
            raise VeilsError(_lib.ERR_INVALID, f"Parameter detr={detr!r} is not 'constant', 'linear' or None!")
        return self._fwd(x, p0, p1, k_offset, False, padding, detr)

    def spectrogram(self, x, y=None, p0=None, p1=None, *, k_offset=0, padding="zeros"):
        """scipy ShortTimeFFT.spectrogram: |stft(x)|^2 (real), or the cross-spectrogram
        stft(x) * conj(stft(y)) (complex) when a second signal y is given (scipy :1398-1404)."""
        if y is None or y is x:
            return self._fwd(x, p0, p1, k_offset, True, padding)
        import torch
        was_numpy = not self._is_tensor(x)
        tdt = torch.float32 if self.precision == "f32" else torch.float64
        if was_numpy:
            x = torch.from_numpy(np.ascontiguousarray(x, dtype=self._rdt)).cuda(self.device)
        if not self._is_tensor(y):
            y = torch.from_numpy(np.ascontiguousarray(y, dtype=self._rdt)).to(x.device)
        if tuple(x.shape) != tuple(y.shape):
            raise VeilsError(_lib.ERR_INVALID, f"x.shape={tuple(x.shape)} must equal y.shape={tuple(y.shape)}")
        Sx = self._fwd(x.to(tdt), p0, p1, k_offset, False, padding)
        Sy = self._fwd(y.to(tdt), p0, p1, k_offset, False, padding)
        squeeze = Sx.dim() == 2
        Sx3, Sy3 = (Sx[None], Sy[None]) if squeeze else (Sx, Sy)
        batch, F, P = Sx3.shape
        st = torch.cuda.current_stream(Sx.device).cuda_stream
        _check(self._L.veils_cross_multiply_dev(self._h, Sx3.data_ptr(), Sy3.data_ptr(), P, batch,
                                                Sx3.stride(1) if F > 1 else P, Sy3.stride(1) if F > 1 else P, st))
        return Sx.cpu().numpy() if was_numpy else Sx

    def istft(self, S, k0=None, k1=None, *, complex_output=False):
        """src/lib.rs:792-914 -> real signal of length k1-k0 (default k0=0, k1=k_max of S).
        complex_output=True (two-sided layouts): the twin's complex result (python/standalone_stft.py:404, 424;
        the Rust crate keeps the real part, src/lib.rs:903-907): Re from S, Im = istft_real(-i S)."""
        self._raise_if_nan()
        if complex_output:
            return self._istft_complex(S, k0, k1)
        if self._is_tensor(S):
            return self._istft_tensor(S, k0, k1)
        S = np.asarray(S)
        if S.ndim not in (2, 3) or S.size == 0:
            raise VeilsError(_lib.ERR_INVALID, "STFT data cannot be empty")
        squeeze = S.ndim == 2
        S3 = np.ascontiguousarray(S[None] if squeeze else S, dtype=self._cdt)
        batch, F, P = S3.shape
        if F != self.f_pts:
            raise VeilsError(_lib.ERR_INVALID,
                             f"STFT frequency dimension {F} must equal {self.f_pts}")
        a, b = self.istft_range(P, k0, k1)
        out = np.empty((batch, b - a), dtype=self._rdt)
        _check(self._L.veils_istft_host(self._h, S3.ctypes.data, P, batch, opt_i64(a), opt_i64(b),
                                        out.ctypes.data))
        return out[0] if squeeze else out

    def _istft_complex(self, S, k0, k1):
        if self.onesided_fft:
            raise VeilsError(_lib.ERR_INVALID, "complex_output needs fft_mode 'twosided' or 'centered'")
        import torch
        was_numpy = not self._is_tensor(S)
        cdt = torch.complex64 if self.precision == "f32" else torch.complex128
        St = torch.from_numpy(np.ascontiguousarray(S)).to(cdt).cuda(self.device) if was_numpy else S
        S3 = (St[None] if St.dim() == 2 else St).contiguous()
        batch, F, P = S3.shape
        T3 = torch.empty_like(S3)
        st = torch.cuda.current_stream(S3.device).cuda_stream
        _check(self._L.veils_complex_op_dev(self._h, 2, T3.data_ptr(), S3.data_ptr(), P, batch, P, P, st))
        y = torch.complex(self._istft_tensor(S3, k0, k1), self._istft_tensor(T3, k0, k1))   # packaging of the two results
        y = y[0] if St.dim() == 2 else y
        return y.cpu().numpy() if was_numpy else y

    def _istft_tensor(self, S, k0, k1):
        import torch
        cdt = torch.complex64 if self.precision == "f32" else torch.complex128
        if not S.is_cuda or S.dtype != cdt:
            raise ValueError(f"expected a CUDA tensor of dtype {cdt}")
        if S.device.index != self.device:
            raise ValueError(f"tensor on cuda:{S.device.index}, plan on cuda:{self.device}")
        squeeze = S.dim() == 2
        S3 = S[None] if squeeze else S
        batch, F, P = S3.shape
        # accept padded-pitch views as produced by stft() without copying
        if not (S3.stride(2) == 1 and S3.stride(1) >= P and (batch == 1 or S3.stride(0) == F * S3.stride(1))):
            S3 = S3.contiguous()
        ldp = S3.stride(1) if F > 1 else P
        if F != self.f_pts:
            raise VeilsError(_lib.ERR_INVALID,
                             f"STFT frequency dimension {F} must equal {self.f_pts}")
        a, b = self.istft_range(P, k0, k1)
        out = torch.empty((batch, b - a), device=S.device,
                          dtype=torch.float32 if self.precision == "f32" else torch.float64)
        st = torch.cuda.current_stream(S.device).cuda_stream
        _check(self._L.veils_istft_dev(self._h, S3.data_ptr(), P, batch, ldp, opt_i64(a), opt_i64(b),
                                       out.data_ptr(), b - a, st))
        return out[0] if squeeze else out
