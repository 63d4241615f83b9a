//! Raw bindings of include/veils_cuda.h (hand-written; one item per C declaration).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct veils_plan { _private: [u8; 0] }

#[repr(C)]
pub struct veils_plan_desc {
    pub win: *const f64,
    pub m_num: i64,
    pub hop: i64,
    pub fs: f64,
    pub fft_mode: *const c_char,
    pub mfft: i64,
    pub dual_win: *const f64,
    pub scale_to: *const c_char,
    pub phase_shift: i64,
    pub precision: i32,
    pub device: i32,
}

pub const VEILS_OK: i32 = 0;
pub const VEILS_F32: i32 = 0;
pub const VEILS_F64: i32 = 1;

extern "C" {
    pub fn veils_plan_create(desc: *const veils_plan_desc, out: *mut *mut veils_plan) -> i32;
    pub fn veils_plan_destroy(plan: *mut veils_plan);
    pub fn veils_last_error() -> *const c_char;
    pub fn veils_m_num(p: *const veils_plan) -> i64;
    pub fn veils_m_num_mid(p: *const veils_plan) -> i64;
    pub fn veils_hop(p: *const veils_plan) -> i64;
    pub fn veils_mfft(p: *const veils_plan) -> i64;
    pub fn veils_f_pts(p: *const veils_plan) -> i64;
    pub fn veils_phase_shift(p: *const veils_plan) -> i64;
    pub fn veils_fs(p: *const veils_plan) -> f64;
    pub fn veils_delta_t(p: *const veils_plan) -> f64;
    pub fn veils_delta_f(p: *const veils_plan) -> f64;
    pub fn veils_fft_mode(p: *const veils_plan) -> i32;
    pub fn veils_onesided_fft(p: *const veils_plan) -> i32;
    pub fn veils_win(p: *const veils_plan, out: *mut f64) -> i32;
    pub fn veils_dual_win(p: *mut veils_plan, out: *mut f64) -> i32;
    pub fn veils_invertible(p: *mut veils_plan) -> i32;
    pub fn veils_p_min(p: *const veils_plan) -> i64;
    pub fn veils_k_min(p: *const veils_plan) -> i64;
    pub fn veils_p_max(p: *const veils_plan, n: i64, out: *mut i64) -> i32;
    pub fn veils_k_max(p: *const veils_plan, n: i64, out: *mut i64) -> i32;
    pub fn veils_p_range(p: *const veils_plan, n: i64, p0: *const i64, p1: *const i64,
                         out_p0: *mut i64, out_p1: *mut i64) -> i32;
    pub fn veils_istft_range(p: *mut veils_plan, cols: i64, k0: *const i64, k1: *const i64,
                             out_k0: *mut i64, out_k1: *mut i64) -> i32;
    pub fn veils_t(p: *const veils_plan, n: i64, p0: *const i64, p1: *const i64, k_offset: i64,
                   out: *mut f64) -> i32;
    pub fn veils_f(p: *const veils_plan, out: *mut f64) -> i32;
    pub fn veils_stft_host(p: *mut veils_plan, x: *const c_void, n: i64, batch: i64, x_pitch: i64,
                           p0: *const i64, p1: *const i64, k_offset: i64, s: *mut c_void) -> i32;
    pub fn veils_spectrogram_host(p: *mut veils_plan, x: *const c_void, n: i64, batch: i64,
                                  x_pitch: i64, p0: *const i64, p1: *const i64, k_offset: i64,
                                  sxx: *mut c_void) -> i32;
    pub fn veils_istft_host(p: *mut veils_plan, s: *const c_void, cols: i64, batch: i64,
                            k0: *const i64, k1: *const i64, x_out: *mut c_void) -> i32;
    pub fn veils_stft_dev(p: *mut veils_plan, x: *const c_void, n: i64, batch: i64, x_pitch: i64,
                          p0: *const i64, p1: *const i64, k_offset: i64, s: *mut c_void, ldp: i64,
                          stream: *mut c_void) -> i32;
    pub fn veils_pad_extent(p: *const veils_plan, n: i64, p0: *const i64, p1: *const i64, k_offset: i64,
                            left: *mut i64, n_ext: *mut i64) -> i32;
    pub fn veils_stft_dev_padded(p: *mut veils_plan, x: *const c_void, n: i64, batch: i64, x_pitch: i64,
                                 p0: *const i64, p1: *const i64, k_offset: i64, padding: *const c_char,
                                 x_ext: *mut c_void, ext_pitch: i64, s: *mut c_void, ldp: i64,
                                 stream: *mut c_void, spectrogram: i32) -> i32;
    pub fn veils_stft_detrend_dev(p: *mut veils_plan, x: *const c_void, n: i64, batch: i64, x_pitch: i64,
                                  p0: *const i64, p1: *const i64, k_offset: i64, padding: *const c_char,
                                  detr: *const c_char, x_ext: *mut c_void, ext_pitch: i64,
                                  coef_ws: *mut c_void, s: *mut c_void, ldp: i64, stream: *mut c_void) -> i32;
    pub fn veils_istft_dev(p: *mut veils_plan, s: *const c_void, cols: i64, batch: i64, ldp: i64,
                           k0: *const i64, k1: *const i64, x_out: *mut c_void, out_pitch: i64,
                           stream: *mut c_void) -> i32;
    pub fn veils_dev_alloc(device: i32, bytes: u64, out: *mut *mut c_void) -> i32;
    pub fn veils_dev_free(device: i32, ptr: *mut c_void) -> i32;
    pub fn veils_dev_upload(dst: *mut c_void, src: *const c_void, bytes: u64, stream: *mut c_void) -> i32;
    pub fn veils_dev_download(dst: *mut c_void, src: *const c_void, bytes: u64, stream: *mut c_void) -> i32;
    pub fn veils_stream_create(device: i32, out: *mut *mut c_void) -> i32;
    pub fn veils_stream_destroy(stream: *mut c_void) -> i32;
    pub fn veils_stream_sync(stream: *mut c_void) -> i32;
}
