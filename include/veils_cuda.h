/* veils_cuda.h -- C ABI of the B200-native STFT engine that replaces the hot path of
 * ultrasaw/veils' `StandaloneSTFT` (Rust, src/lib.rs).
 *
 * The reference has no FFI of its own: its boundary is the crate's public Rust API
 * (src/lib.rs:166-281, 324-333, 587-609, 689-749, 792-949).  Every entry point below names
 * the reference item it replaces, so that a thin Rust `ffi` module can
 * re-export the identical `StandaloneSTFT` type on top of this library.
 *
 * Conventions
 *   - Every fallible call returns an int32 status (VEILS_OK == 0).  The message of the last
 *     failure on the calling thread is returned by veils_last_error(); messages carry the
 *     reference's own `Err(String)` texts.  Nothing panics, aborts or throws across the ABI.
 *   - All sample / slice indices are 64 bit (the reference uses i32, src/lib.rs:617-619).
 *   - `Option<T>` arguments of the Rust API are nullable pointers (`NULL` == `None`).
 *   - Outputs are caller-allocated; sizes come from the geometry calls.  Nothing allocated
 *     on one side of the boundary is freed on the other.
 *   - STFT layout is the reference's `[f_pts][time]` (src/lib.rs:735-746), batched:
 *     S[b][f][p] at element offset (b * f_pts + f) * ldp + p, interleaved complex {re, im},
 *     ldp >= P.  Signals: x[b][i] at b * x_pitch + i.
 *   - Element type of device and host buffers == the plan's precision (float / double).
 *   - A plan is immutable after creation: all *_dev compute entry points may be called
 *     concurrently from several threads on distinct streams.  The *_host flavours stage
 *     through device buffers owned by the plan (chunked, two streams: upload, transform and
 *     download overlap when the host buffers are page-locked, see veils_host_alloc_pinned);
 *     concurrent *_host calls on ONE plan serialise on an internal lock.
 */
#ifndef VEILS_CUDA_H
#define VEILS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VEILS_ABI_VERSION 1

/* status codes */
enum {
    VEILS_OK = 0,
    VEILS_ERR_INVALID = 1,        /* the reference's Err(String) validation failures          */
    VEILS_ERR_NOT_INVERTIBLE = 2, /* src/lib.rs:288-318                                       */
    VEILS_ERR_CUDA = 3,           /* CUDA runtime failure / no device / kernel image missing  */
    VEILS_ERR_UNSUPPORTED = 4,
    VEILS_ERR_NCCL = 5
};

/* enum FftMode (src/lib.rs:54-84) */
enum { VEILS_TWOSIDED = 0, VEILS_CENTERED = 1, VEILS_ONESIDED = 2, VEILS_ONESIDED2X = 3 };
/* scale_to (not in the reference; scipy.signal.ShortTimeFFT.scale_to) */
enum { VEILS_SCALE_NONE = 0, VEILS_SCALE_MAGNITUDE = 1, VEILS_SCALE_PSD = 2 };
/* arithmetic / storage precision of the device path */
enum { VEILS_F32 = 0, VEILS_F64 = 1 };

typedef struct veils_plan veils_plan;

/* Arguments of StandaloneSTFT::new (src/lib.rs:166-174), plus scale_to (scipy keyword),
 * the device precision and the CUDA device ordinal. */
typedef struct veils_plan_desc {
    const double *win;      /* win: Vec<f64>, copied                                        */
    int64_t m_num;          /* win.len()                                                    */
    int64_t hop;            /* hop: usize                                                   */
    double fs;              /* fs: f64                                                      */
    const char *fft_mode;   /* Option<&str>; NULL -> "onesided" (src/lib.rs:192)            */
    int64_t mfft;           /* Option<usize>; 0 -> m_num (src/lib.rs:193)                   */
    const double *dual_win; /* Option<Vec<f64>>; NULL -> canonical dual, computed lazily    */
    const char *scale_to;   /* NULL | "magnitude" | "psd"                                   */
    int64_t phase_shift;    /* Option<i32>; 0 == None (src/lib.rs:194)                      */
    int32_t precision;      /* VEILS_F32 | VEILS_F64                                        */
    int32_t device;         /* CUDA ordinal used by the compute calls                       */
} veils_plan_desc;

/* ---- lifetime ------------------------------------------------------------------------ */
int32_t veils_abi_version(void);
/* StandaloneSTFT::new (src/lib.rs:166-224).  Host-only: needs no GPU. */
int32_t veils_plan_create(const veils_plan_desc *desc, veils_plan **out);
void veils_plan_destroy(veils_plan *plan);
/* thread-local message of the last failed call ("" if none) */
const char *veils_last_error(void);

/* ---- accessors (src/lib.rs:227-281) --------------------------------------------------- */
int64_t veils_m_num(const veils_plan *p);       /* m_num()      */
int64_t veils_m_num_mid(const veils_plan *p);   /* m_num_mid()  */
int64_t veils_hop(const veils_plan *p);         /* hop()        */
int64_t veils_mfft(const veils_plan *p);        /* mfft()       */
int64_t veils_f_pts(const veils_plan *p);       /* f_pts()      */
int64_t veils_phase_shift(const veils_plan *p); /* phase_shift()*/
double veils_fs(const veils_plan *p);           /* fs()         */
double veils_delta_t(const veils_plan *p);      /* delta_t()    */
double veils_delta_f(const veils_plan *p);      /* delta_f()    */
int32_t veils_fft_mode(const veils_plan *p);    /* fft_mode()   */
int32_t veils_onesided_fft(const veils_plan *p);/* onesided_fft()*/
int32_t veils_scaling(const veils_plan *p);     /* scipy .scaling */
int32_t veils_precision(const veils_plan *p);
int32_t veils_device(const veils_plan *p);
/* win() -- the (scaled) window, m_num doubles (src/lib.rs:228) */
int32_t veils_win(const veils_plan *p, double *out);
/* dual_win() (src/lib.rs:324-329): m_num doubles; VEILS_ERR_NOT_INVERTIBLE with the
 * reference's message when no dual window exists. */
int32_t veils_dual_win(veils_plan *p, double *out);
/* invertible() (src/lib.rs:331-333): 1 / 0 */
int32_t veils_invertible(veils_plan *p);

/* ---- slice / sample geometry (src/lib.rs:524-609; 64-bit) ----------------------------- */
int64_t veils_p_min(const veils_plan *p);                                  /* p_min()   */
int64_t veils_k_min(const veils_plan *p);                                  /* pre_padding().0 */
int32_t veils_p_max(const veils_plan *p, int64_t n, int64_t *out);         /* p_max(n): error instead of the panic at :562-564 */
int32_t veils_k_max(const veils_plan *p, int64_t n, int64_t *out);         /* post_padding(n).0 */
int32_t veils_p_num(const veils_plan *p, int64_t n, int64_t *out);         /* p_max(n) - p_min */
/* p_range(n, p0, p1) (src/lib.rs:595-609) */
int32_t veils_p_range(const veils_plan *p, int64_t n, const int64_t *p0, const int64_t *p1,
                      int64_t *out_p0, int64_t *out_p1);
/* Validation and default resolution of istft's (k0, k1) for an S with P slices
 * (src/lib.rs:798-847).  Output length is *out_k1 - *out_k0. */
int32_t veils_istft_range(veils_plan *p, int64_t P, const int64_t *k0, const int64_t *k1,
                          int64_t *out_k0, int64_t *out_k1);
/* t(n, p0, p1, k_offset) (src/lib.rs:916-928): writes p1-p0 doubles */
int32_t veils_t(const veils_plan *p, int64_t n, const int64_t *p0, const int64_t *p1,
                int64_t k_offset, double *out);
/* f() (src/lib.rs:930-949; centered follows scipy): writes f_pts doubles */
int32_t veils_f(const veils_plan *p, double *out);
/* ---- scipy.signal.ShortTimeFFT axis helpers (host integer logic; the reference crate has none -- SURVEY 8 f4) */
/* lower_border_end: first sample *k / slice *pp not touched by the left border (_short_time_fft.py lower_border_end) */
int32_t veils_lower_border_end(const veils_plan *p, int64_t *k, int64_t *pp);
/* upper_border_begin(n): first sample / slice touched by the right border of an n-sample signal */
int32_t veils_upper_border_begin(const veils_plan *p, int64_t n, int64_t *k, int64_t *pp);
/* extent(n, axes_seq, center_bins): out = (t0, t1, f0, f1), or (f0, f1, t0, t1) when ft_order != 0; two-sided layouts
 * other than "centered" are an error, as in scipy */
int32_t veils_extent(const veils_plan *p, int64_t n, int32_t ft_order, int32_t center_bins, double *out4);
/* nearest_k_p(k, left): nearest sample index that is a slice position */
int64_t veils_nearest_k_p(const veils_plan *p, int64_t k, int32_t left);
/* StandaloneSTFT::debug_ifft_func (src/lib.rs:952-954): ifft_func (:414-505) of ONE slice, evaluated on the
 * host in f64 (debug helper, O(mfft^2)).  X: f_pts complex {re, im}; out: m_num complex {re, im}. */
int32_t veils_debug_ifft_func(const veils_plan *p, const double *X, double *out);

/* ---- compute: device pointers + stream (the measured path) ---------------------------- */
/* stft (src/lib.rs:689-749) for `batch` signals of n samples.  x, S are device pointers of
 * the plan's precision; stream is a cudaStream_t (NULL == default stream).  The call is
 * asynchronous with respect to the host. */
int32_t veils_stft_dev(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                       const int64_t *p0, const int64_t *p1, int64_t k_offset,
                       void *S, int64_t ldp, void *stream);
/* spectrogram (scipy ShortTimeFFT.spectrogram, |S|^2): real output [batch][f_pts][ldp] */
int32_t veils_spectrogram_dev(veils_plan *p, const void *x, int64_t n, int64_t batch,
                              int64_t x_pitch, const int64_t *p0, const int64_t *p1,
                              int64_t k_offset, void *Sxx, int64_t ldp, void *stream);
/* The same with the reference's NaN screen (src/lib.rs:696-698: "Complex-valued input not allowed for
 * one-sided FFT modes"): *nan_flag -- device memory or page-locked host memory, zeroed by the caller --
 * is set to 1 when a NaN sample went into a one-sided transform (a < 1 % pass over the DC row of S behind
 * the transform; slices whose DC bin is not finite are re-read and tested exactly).  spectrogram != 0:
 * Sxx is the real |X|^2 array.  The host-pointer entry points always screen and return the error. */
int32_t veils_stft_dev_checked(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                               const int64_t *p0, const int64_t *p1, int64_t k_offset, void *S, int64_t ldp,
                               void *stream, int32_t spectrogram, int32_t *nan_flag);
/* istft (src/lib.rs:792-914): S[batch][f_pts][ldp] with P valid slices -> x_out[batch][k1-k0] */
int32_t veils_istft_dev(veils_plan *p, const void *S, int64_t P, int64_t batch, int64_t ldp,
                        const int64_t *k0, const int64_t *k1, void *x_out, int64_t out_pitch,
                        void *stream);

/* stft with the twin's border handling (python/standalone_stft.py:309-336; the Rust crate pads
 * with zeros only, src/lib.rs:631-645): padding = NULL | "zeros" | "edge" | "even" | "odd".
 * For the non-zero modes the signal is first extended on the device into the caller's workspace
 * x_ext (batch rows of ext_pitch >= n_ext elements; sizes from veils_pad_extent), then transformed
 * by the same kernels.  spectrogram != 0 writes |S|^2 (real) instead. */
int32_t veils_pad_extent(const veils_plan *p, int64_t n, const int64_t *p0, const int64_t *p1,
                         int64_t k_offset, int64_t *left, int64_t *n_ext);
int32_t veils_stft_dev_padded(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                              const int64_t *p0, const int64_t *p1, int64_t k_offset,
                              const char *padding, void *x_ext, int64_t ext_pitch, void *S,
                              int64_t ldp, void *stream, int32_t spectrogram);

/* stft_detrend of the twin (python/standalone_stft.py:38-50, 345-371; not in the Rust crate):
 * detr = NULL | "constant" (slice mean removed) | "linear" (least-squares line removed), applied
 * to every slice before windowing -- realised as the rank-2 correction S - a_p W1 - b_p W0 of the
 * plain stft.  coef_ws: device workspace of 16 bytes per slice and signal (batch * (p1-p0) * 16).
 * Accuracy: the correction is applied to S already rounded to the plan's precision, so in f32 its
 * absolute error is ~1e-7 of the UN-detrended spectrum's magnitude; for signals with a large offset or
 * ramp (where that dwarfs the detrended content) use an f64 plan, which meets the twin at 1e-12. */
int32_t veils_stft_detrend_dev(veils_plan *p, const void *x, int64_t n, int64_t batch,
                               int64_t x_pitch, const int64_t *p0, const int64_t *p1,
                               int64_t k_offset, const char *padding, const char *detr, void *x_ext,
                               int64_t ext_pitch, void *coef_ws, void *S, int64_t ldp, void *stream);

/* Cross-spectrogram of scipy's ShortTimeFFT.spectrogram(x, y) (_short_time_fft.py:1401-1404):
 * Sx <- Sx * conj(Sy), element-wise over [batch][f_pts][P] (row pitches ldx, ldy), after two stft calls. */
int32_t veils_cross_multiply_dev(veils_plan *p, void *Sx, const void *Sy, int64_t P, int64_t batch,
                                 int64_t ldx, int64_t ldy, void *stream);
/* Helpers for the twin's complex signals in the two-sided layouts (python/standalone_stft.py:349, 404, 424; the
 * Rust crate is real-only): over batch * f_pts rows of P complex elements, op 1: A += i B -- stft(x) of a
 * complex x is stft(Re x) + i stft(Im x); op 2: A = -i B -- Im(istft(S)) = istft_real(-i S). */
int32_t veils_complex_op_dev(veils_plan *p, int32_t op, void *A, const void *B, int64_t P, int64_t batch, int64_t lda,
                             int64_t ldb, void *stream);

/* ---- compute: host pointers (drop-in for the Vec API; copies inside) ------------------- */
/* Also performs the reference's NaN scan for one-sided modes (src/lib.rs:696-698).
 * S is dense: ldp == P.  Synchronous. */
int32_t veils_stft_host(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                        const int64_t *p0, const int64_t *p1, int64_t k_offset, void *S);
int32_t veils_spectrogram_host(veils_plan *p, const void *x, int64_t n, int64_t batch,
                               int64_t x_pitch, const int64_t *p0, const int64_t *p1,
                               int64_t k_offset, void *Sxx);
int32_t veils_istft_host(veils_plan *p, const void *S, int64_t P, int64_t batch,
                         const int64_t *k0, const int64_t *k1, void *x_out);

/* ---- device-buffer and stream handles (so a Rust caller needs no CUDA bindings) -------- */
int32_t veils_device_count(int32_t *out);
int32_t veils_dev_alloc(int32_t device, uint64_t bytes, void **out);
int32_t veils_dev_free(int32_t device, void *ptr);
int32_t veils_dev_upload(void *dst_dev, const void *src_host, uint64_t bytes, void *stream);
int32_t veils_dev_download(void *dst_host, const void *src_dev, uint64_t bytes, void *stream);
int32_t veils_host_alloc_pinned(uint64_t bytes, void **out);
int32_t veils_host_free_pinned(void *ptr);
int32_t veils_stream_create(int32_t device, void **out);
int32_t veils_stream_destroy(void *stream);
int32_t veils_stream_sync(void *stream);

/* ---- introspection used by bench.py / the tests ---------------------------------------- */
/* Number of kernels this library has launched in this process (all threads). */
/* host-side cost of one veils_stft_dev call in microseconds: `reps` back-to-back calls timed inside the library */
int32_t veils_debug_launch_cost_us(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch, void *S,
                                   int64_t ldp, void *stream, int32_t reps, double *us_per_call);
uint64_t veils_kernel_launch_count(void);
/* Name of the kernel family the plan dispatches to for forward / inverse: "pow2_wide" (f32, mfft 1024 / 2048 /
 * 4096, m_num = mfft = 4 hop) | "pow2_packed" (f32 powers of two) | "pow2_tile" (f32 / f64 powers of two) |
 * "mixed_radix" (any mfft = product of <= 4 radices 2 .. 16) | "dft_generic" (everything else); static storage. */
const char *veils_plan_kernel_name(const veils_plan *p, int32_t inverse);

#ifdef __cplusplus
}
#endif
#endif /* VEILS_CUDA_H */
