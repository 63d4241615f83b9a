//! `veils::StandaloneSTFT` with the public API of ultrasaw/veils 0.1.2 (reference src/lib.rs:116-955),
//! backed by the sm_100a kernels in libveils_cuda.so.  Same constructor, methods, `Option`
//! defaults, `Result<_, String>` error texts and `[freq][time]` layout; `scale_to` and
//! `spectrogram` are the two additions BASELINE.json asks for (scipy semantics).
mod ffi;
use num_complex::Complex;
use std::ffi::{CStr, CString};
use std::os::raw::c_void;
use std::ptr;

/// reference src/lib.rs:54-84
#[derive(Debug, Clone, PartialEq)]
pub enum FftMode {
    TwoSided,
    Centered,
    OneSided,
    OneSided2X,
}

impl std::str::FromStr for FftMode {
    type Err = String;
    fn from_str(s: &str) -> Result<Self, Self::Err> {
        match s {
            "twosided" => Ok(FftMode::TwoSided),
            "centered" => Ok(FftMode::Centered),
            "onesided" => Ok(FftMode::OneSided),
            "onesided2X" => Ok(FftMode::OneSided2X),
            _ => Err(format!("Unknown fft_mode: {}", s)),
        }
    }
}

impl FftMode {
    pub fn is_onesided(&self) -> bool {
        matches!(self, FftMode::OneSided | FftMode::OneSided2X)
    }
    fn from_code(c: i32) -> FftMode {
        match c {
            ffi::VEILS_TWOSIDED => FftMode::TwoSided,
            ffi::VEILS_CENTERED => FftMode::Centered,
            ffi::VEILS_ONESIDED2X => FftMode::OneSided2X,
            _ => FftMode::OneSided,
        }
    }
}

pub struct StandaloneSTFT {
    plan: *mut ffi::veils_plan,
    fft_mode: FftMode,
}

unsafe impl Send for StandaloneSTFT {}

fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::veils_last_error()).to_string_lossy().into_owned() }
}
fn check(rc: i32) -> Result<(), String> {
    if rc == ffi::VEILS_OK { Ok(()) } else { Err(last_error()) }
}
fn opt(v: &Option<i64>) -> *const i64 {
    v.as_ref().map_or(ptr::null(), |r| r as *const i64)
}

impl StandaloneSTFT {
    /// reference src/lib.rs:166-174, plus `scale_to` (scipy keyword order) -- f64 plan
    pub fn new(win: Vec<f64>, hop: usize, fs: f64, fft_mode: Option<&str>, mfft: Option<usize>,
               dual_win: Option<Vec<f64>>, phase_shift: Option<i32>) -> Result<Self, String> {
        Self::with_scaling(win, hop, fs, fft_mode, mfft, dual_win, None, phase_shift)
    }

    #[allow(clippy::too_many_arguments)]
    pub fn with_scaling(win: Vec<f64>, hop: usize, fs: f64, fft_mode: Option<&str>,
                        mfft: Option<usize>, dual_win: Option<Vec<f64>>, scale_to: Option<&str>,
                        phase_shift: Option<i32>) -> Result<Self, String> {
        if let Some(dw) = &dual_win {
            if dw.len() != win.len() {
                return Err("dual_win.len() must equal win.len()!".to_string());
            }
        }
        let mode = fft_mode.map(|m| CString::new(m).unwrap());
        let scale = scale_to.map(|m| CString::new(m).unwrap());
        let desc = ffi::veils_plan_desc {
            win: win.as_ptr(),
            m_num: win.len() as i64,
            hop: hop as i64,
            fs,
            fft_mode: mode.as_ref().map_or(ptr::null(), |m| m.as_ptr()),
            mfft: mfft.unwrap_or(0) as i64,
            dual_win: dual_win.as_ref().map_or(ptr::null(), |d| d.as_ptr()),
            scale_to: scale.as_ref().map_or(ptr::null(), |m| m.as_ptr()),
            phase_shift: phase_shift.unwrap_or(0) as i64,
            precision: ffi::VEILS_F64,
            device: 0,
        };
        let mut plan = ptr::null_mut();
        check(unsafe { ffi::veils_plan_create(&desc, &mut plan) })?;
        let fft_mode = FftMode::from_code(unsafe { ffi::veils_fft_mode(plan) });
        Ok(StandaloneSTFT { plan, fft_mode })
    }

    /// reference src/lib.rs:239
    pub fn fft_mode(&self) -> &FftMode { &self.fft_mode }
    pub fn hop(&self) -> usize { unsafe { ffi::veils_hop(self.plan) as usize } }
    pub fn fs(&self) -> f64 { unsafe { ffi::veils_fs(self.plan) } }
    pub fn mfft(&self) -> usize { unsafe { ffi::veils_mfft(self.plan) as usize } }
    pub fn phase_shift(&self) -> i32 { unsafe { ffi::veils_phase_shift(self.plan) as i32 } }
    pub fn sampling_period(&self) -> f64 { 1.0 / self.fs() }
    pub fn delta_t(&self) -> f64 { unsafe { ffi::veils_delta_t(self.plan) } }
    pub fn delta_f(&self) -> f64 { unsafe { ffi::veils_delta_f(self.plan) } }
    pub fn m_num(&self) -> usize { unsafe { ffi::veils_m_num(self.plan) as usize } }
    pub fn m_num_mid(&self) -> usize { unsafe { ffi::veils_m_num_mid(self.plan) as usize } }
    pub fn f_pts(&self) -> usize { unsafe { ffi::veils_f_pts(self.plan) as usize } }
    pub fn onesided_fft(&self) -> bool { unsafe { ffi::veils_onesided_fft(self.plan) != 0 } }
    pub fn p_min(&self) -> i32 { unsafe { ffi::veils_p_min(self.plan) as i32 } }

    pub fn win(&self) -> Vec<Complex<f64>> {
        let mut w = vec![0.0f64; self.m_num()];
        unsafe { ffi::veils_win(self.plan, w.as_mut_ptr()) };
        w.into_iter().map(|x| Complex::new(x, 0.0)).collect()
    }
    pub fn dual_win(&mut self) -> Result<Vec<Complex<f64>>, String> {
        let mut w = vec![0.0f64; self.m_num()];
        check(unsafe { ffi::veils_dual_win(self.plan, w.as_mut_ptr()) })?;
        Ok(w.into_iter().map(|x| Complex::new(x, 0.0)).collect())
    }
    pub fn invertible(&mut self) -> bool { unsafe { ffi::veils_invertible(self.plan) != 0 } }

    /// reference panics for n < ceil(m_num/2) (src/lib.rs:562-564); here an Err
    pub fn p_max(&self, n: usize) -> Result<i32, String> {
        let mut out = 0i64;
        check(unsafe { ffi::veils_p_max(self.plan, n as i64, &mut out) })?;
        Ok(out as i32)
    }
    pub fn p_range(&self, n: usize, p0: Option<i32>, p1: Option<i32>) -> Result<(i32, i32), String> {
        let (a, b) = (p0.map(i64::from), p1.map(i64::from));
        let (mut o0, mut o1) = (0i64, 0i64);
        check(unsafe { ffi::veils_p_range(self.plan, n as i64, opt(&a), opt(&b), &mut o0, &mut o1) })?;
        Ok((o0 as i32, o1 as i32))
    }

    /// reference src/lib.rs:689-749: returns S[freq][time]
    pub fn stft(&self, x: &[f64], p0: Option<i32>, p1: Option<i32>, k_offset: Option<i32>)
                -> Result<Vec<Vec<Complex<f64>>>, String> {
        let (a, b) = self.p_range(x.len(), p0, p1)?;
        let (f, p) = (self.f_pts(), (b - a) as usize);
        let mut flat = vec![Complex::new(0.0f64, 0.0); f * p];
        let (a64, b64) = (Some(a as i64), Some(b as i64));
        check(unsafe {
            ffi::veils_stft_host(self.plan, x.as_ptr() as *const c_void, x.len() as i64, 1,
                                 x.len() as i64, opt(&a64), opt(&b64), k_offset.unwrap_or(0) as i64,
                                 flat.as_mut_ptr() as *mut c_void)
        })?;
        Ok(flat.chunks(p).map(|r| r.to_vec()).collect())
    }

    /// scipy ShortTimeFFT.spectrogram: |stft|^2, Sxx[freq][time]
    pub fn spectrogram(&self, x: &[f64], p0: Option<i32>, p1: Option<i32>, k_offset: Option<i32>)
                       -> Result<Vec<Vec<f64>>, String> {
        let (a, b) = self.p_range(x.len(), p0, p1)?;
        let (f, p) = (self.f_pts(), (b - a) as usize);
        let mut flat = vec![0.0f64; f * p];
        let (a64, b64) = (Some(a as i64), Some(b as i64));
        check(unsafe {
            ffi::veils_spectrogram_host(self.plan, x.as_ptr() as *const c_void, x.len() as i64, 1,
                                        x.len() as i64, opt(&a64), opt(&b64),
                                        k_offset.unwrap_or(0) as i64, flat.as_mut_ptr() as *mut c_void)
        })?;
        Ok(flat.chunks(p).map(|r| r.to_vec()).collect())
    }

    /// reference src/lib.rs:792-914
    pub fn istft(&mut self, stft_data: &[Vec<Complex<f64>>], k0: Option<i32>, k1: Option<i32>)
                 -> Result<Vec<f64>, String> {
        if stft_data.is_empty() { return Err("STFT data cannot be empty".to_string()); }
        let f = self.f_pts();
        if stft_data.len() != f {
            return Err(format!("STFT frequency dimension {} must equal {}", stft_data.len(), f));
        }
        let p = stft_data[0].len();
        let flat: Vec<Complex<f64>> = stft_data.iter().flat_map(|r| r.iter().copied()).collect();
        let (a, b) = (k0.map(i64::from), k1.map(i64::from));
        let (mut o0, mut o1) = (0i64, 0i64);
        check(unsafe { ffi::veils_istft_range(self.plan, p as i64, opt(&a), opt(&b), &mut o0, &mut o1) })?;
        let mut out = vec![0.0f64; (o1 - o0) as usize];
        let (s0, s1) = (Some(o0), Some(o1));
        check(unsafe {
            ffi::veils_istft_host(self.plan, flat.as_ptr() as *const c_void, p as i64, 1, opt(&s0),
                                  opt(&s1), out.as_mut_ptr() as *mut c_void)
        })?;
        Ok(out)
    }

    pub fn t(&self, n: usize, p0: Option<i32>, p1: Option<i32>, k_offset: Option<i32>)
             -> Result<Vec<f64>, String> {
        let (a, b) = self.p_range(n, p0, p1)?;
        let mut out = vec![0.0f64; (b - a) as usize];
        let (a64, b64) = (Some(a as i64), Some(b as i64));
        check(unsafe { ffi::veils_t(self.plan, n as i64, opt(&a64), opt(&b64),
                                    k_offset.unwrap_or(0) as i64, out.as_mut_ptr()) })?;
        Ok(out)
    }
    /// reference src/lib.rs:952-954: ifft_func of one slice (f_pts bins -> m_num samples)
    pub fn debug_ifft_func(&self, x: Vec<Complex<f64>>) -> Vec<Complex<f64>> {
        let mut out = vec![Complex::new(0.0f64, 0.0); self.m_num()];
        if x.len() == self.f_pts() {
            unsafe { ffi::veils_debug_ifft_func(self.plan, x.as_ptr() as *const f64, out.as_mut_ptr() as *mut f64) };
        }
        out
    }
    pub fn f(&self) -> Vec<f64> {
        let mut out = vec![0.0f64; self.f_pts()];
        unsafe { ffi::veils_f(self.plan, out.as_mut_ptr()) };
        out
    }
}

impl Drop for StandaloneSTFT {
    fn drop(&mut self) { unsafe { ffi::veils_plan_destroy(self.plan) } }
}
