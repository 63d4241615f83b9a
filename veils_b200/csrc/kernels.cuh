// kernels.cuh -- launch interface between the C ABI (api.cu) and the kernel families.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>

namespace veils {

// Forward problem: frames p0 .. p0+P of `batch` signals (reference src/lib.rs:612-749).
struct FwdParams {
    int64_t n, x_pitch, batch;          // input signals x[b][0..n)
    int64_t N, H, M, F, mid, ps;        // m_num, hop, mfft, f_pts, m_num_mid, roll
    int64_t p0, P, k_offset;            // slice range and sample offset
    int64_t ldp;                        // row pitch of S (elements), >= P
    int mode;                           // Mode
    double fac2x;                       // onesided2X factor
    int spectrogram;                    // 1: write re^2 + im^2 (real) instead of complex
};

// Inverse problem (reference src/lib.rs:792-914).
struct InvParams {
    int64_t P, ldp, batch;              // S[b][F][ldp] with P valid slices
    int64_t N, H, M, F, mid, ps;
    int64_t p_min;                      // S column c holds slice q = c + p_min
    int64_t k0, k1, q0, q1;             // output samples [k0, k1), slices [q0, q1)
    int64_t out_pitch;
    int mode;
    double fac2x;
};

// Device tables of one plan (element type T = float / double).
template <typename T>
struct Tables {
    const T *win = nullptr;             // [N] window (scaled)
    const T *dual = nullptr;            // [N] dual window (nullptr until istft is used)
    const double2 *tw_full = nullptr;   // [M] exp(-2 pi i m / M) in f64 (generic path)
    const T *win_half = nullptr;        // [N] window / 2           (pow2 forward)
    const T *dual_scaled = nullptr;     // [N] dual / M or / (2M)   (pow2 inverse)
    const T *tw_pow2 = nullptr;         // pow2 twiddles (layout: pow2_cfg.hpp)
    const float *tw_pk = nullptr;       // packed-f32 forward twiddles (layout: pow2_pk_cfg.hpp)
    const float *tw_pk_inv = nullptr;   // packed-f32 inverse twiddles
    const void *ptab_pk = nullptr;      // packed-f32 tile table [M/2] x 16 bytes
    const void *itab_pk = nullptr;      // packed-f32 inverse table [M/2] x 16 bytes
    const void *smooth_fwd = nullptr;   // mixed-radix family: plan blobs (layout: smooth_impl.cuh)
    const void *smooth_inv = nullptr;
    const float *wide_fwd = nullptr;    // wide family: forward table (layout: wide_impl.cuh)
    const float *wide_inv = nullptr;    // wide family: inverse table
};

extern std::atomic<unsigned long long> g_launch_count;   // kernels launched by this library

// ---- generic O(M^2) DFT family: any mfft, any mode (dft_generic.cu) ----------------------
template <typename T>
cudaError_t dft_generic_forward(const T *x, void *out, const FwdParams &prm, const Tables<T> &tb,
                                cudaStream_t st);
template <typename T>
cudaError_t dft_generic_inverse(const void *S, T *x_out, const InvParams &prm, const Tables<T> &tb,
                                cudaStream_t st);

// ---- power-of-two tile family (pow2_tile.cu) ----------------------------------------------
bool pow2_supported(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse);
template <typename T>
cudaError_t pow2_forward(const T *x, void *out, const FwdParams &prm, const Tables<T> &tb,
                         cudaStream_t st);
template <typename T>
cudaError_t pow2_inverse(const void *S, T *x_out, const InvParams &prm, const Tables<T> &tb,
                         cudaStream_t st);
// size (in T elements) and contents of the fast-path twiddle table for mfft = M
int64_t pow2_twiddle_count(int64_t M);
void pow2_twiddle_fill(int64_t M, double *re_im_pairs);

// ---- packed f32 family: two slices per lane, FADD2/FFMA2 (pow2_pk.cu) ----------------------
bool pk_supported(int64_t M, int64_t N, int64_t H, int64_t ps, bool inverse);
int64_t pk_twiddle_count(int64_t M);
void pk_twiddle_fill(int64_t M, bool inverse, double *entries4);
// tile table (offset + window per sample pair), M/2 entries of 16 bytes, from the unscaled window
// inverse table (overlap-row offsets + dual / M per sample pair) from the scaled dual window
void pk_itab_fill_host(int64_t M, int64_t N, int64_t ps, const double *dual_scaled, void *out16);
void pk_ptab_fill_host(int64_t M, int64_t N, int64_t H, int64_t ps, const double *win, void *out16);
cudaError_t pk_forward(const float *x, void *out, const FwdParams &prm, const Tables<float> &tb, cudaStream_t st);
cudaError_t pk_inverse(const void *S, float *x_out, const InvParams &prm, const Tables<float> &tb, cudaStream_t st);

// ---- wide f32 family: mfft 1024 / 2048 / 4096 with m_num = mfft = 4 hop, phase_shift = 0 (wide_fwd.cu, wide_inv.cu)
bool wide_supported(int64_t M, int64_t N, int64_t H, int64_t ps);
int wide_tf(int64_t M);                                   // slices per tile (0: mfft not covered)
int64_t wide_table_floats(int64_t M);
// plan table of one direction: window * 0.5 (forward) or dual * scale (inverse) + twiddles, as doubles
void wide_table_fill_host(int64_t M, int tf, bool inverse, const double *vec_n, double scale, double *out);
cudaError_t wide_forward(const float *x, void *out, const FwdParams &prm, const float *tab, int tf, cudaStream_t st);
cudaError_t wide_inverse(const void *S, float *x_out, const InvParams &prm, const float *tab, int tf, cudaStream_t st);

// ---- mixed-radix family: any mfft = product of <= 4 radices 2 .. 16, f32 / f64 (smooth.cu) ------------
bool smooth_supported(int64_t M, int64_t N, int64_t H, int precision, bool inverse);
// plan blob of one direction (twiddles + load / split / row tables; layout: smooth_impl.cuh) in the plan's precision;
// vec_n = the window (forward) or the dual window scaled by 1/M resp. 1/(2M) (inverse)
int64_t smooth_table_bytes(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse);
void smooth_table_fill(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse, const double *vec_n, void *out);
template <typename T>
cudaError_t smooth_forward(const T *x, void *out, const FwdParams &prm, const void *tab, cudaStream_t st);
template <typename T>
cudaError_t smooth_inverse(const void *S, T *x_out, const InvParams &prm, const void *tab, cudaStream_t st);

// ---- border extension for the twin's padding modes (pad.cu): mode 0 zeros, 1 edge, 2 even, 3 odd
template <typename T>
cudaError_t pad_signal(const T *x, T *x_ext, int64_t n, int64_t batch, int64_t x_pitch, int64_t left, int64_t n_ext,
                       int64_t ext_pitch, int mode, cudaStream_t st);

// ---- per-slice detrending as a rank-2 correction of S (detrend.cu) ---------------------------
template <typename T>
cudaError_t detrend_correct(const T *x, int64_t n, int64_t batch, int64_t x_pitch, int64_t p0, int64_t P, int64_t H,
                            int64_t N, int64_t start0, int linear, const T *W, double2 *coef, T *S, int64_t F,
                            int64_t ldp, cudaStream_t st);

// Sx <- Sx * conj(Sy) over rows x P complex elements (cross-spectrogram, detrend.cu)
template <typename T>
cudaError_t cross_multiply(T *Sx, const T *Sy, int64_t rows, int64_t P, int64_t ldx, int64_t ldy, cudaStream_t st);

// A += i B (op 1) / A = -i B (op 2) over rows x P complex elements (complex input / complex istft of the
// two-sided layouts, python/standalone_stft.py:349, 404, 424; detrend.cu)
template <typename T>
cudaError_t complex_op(T *A, const T *B, int64_t rows, int64_t P, int64_t lda, int64_t ldb, int op, cudaStream_t st);

}  // namespace veils
