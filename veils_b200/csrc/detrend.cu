// detrend.cu -- per-slice detrending of the twin's stft_detrend (python/standalone_stft.py:38-50,
// 345-371: 'constant' removes the slice mean, 'linear' the least-squares line; not in the Rust crate).
// The transform is linear, so instead of touching the frames the correction is applied to S:
//     S_detr[:, p] = S[:, p] - a_p W1 - b_p W0,    W0 = fft_func(win * 1), W1 = fft_func(win * i)
// with (a_p, b_p) the line fitted to slice p.  Two small kernels: slice moments -> (a, b) per slice
// (f64 accumulation), and the rank-2 update of S.  W0 / W1 come from the plan's own forward kernel
// (api.cu), so they carry exactly the mode epilogue / roll / scaling of S.
#include "kernels.cuh"

namespace veils {

namespace {

// one warp per slice: s0 = sum x_i, s1 = sum i x_i over the m_num samples of the slice (zeros outside [0, n))
template <typename T>
__global__ void detrend_coef_kernel(const T *__restrict__ x, int64_t n, int64_t x_pitch, int64_t p0, int64_t P,
                                    int64_t H, int64_t N, int64_t start0, int linear, double2 *__restrict__ coef) {
    const int lane = threadIdx.x & 31;
    const int64_t p = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t b = blockIdx.y;
    if (p >= P) return;
    const T *xb = x + b * x_pitch;
    const int64_t s0i = (p0 + p) * H + start0;          // first sample of the slice
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = lane; i < N; i += 32) {
        const int64_t s = s0i + i;
        if (s >= 0 && s < n) {
            const double v = (double)xb[s];
            s0 += v;
            s1 += (double)i * v;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
        const double dn = (double)N, mean = s0 / dn, tbar = 0.5 * (dn - 1.0);
        double a = 0.0, c = mean;
        if (linear && N > 1) {
            const double den = dn * (dn * dn - 1.0) / 12.0;          // sum (i - tbar)^2
            a = (s1 - tbar * s0) / den;
            c = mean - a * tbar;
        }
        coef[b * P + p] = make_double2(a, c);
    }
}

// S[b][f][p] -= a_p W1[f] + c_p W0[f]      (W: [2][F] complex, W[0] = W0, W[1] = W1)
template <typename T>
__global__ void detrend_apply_kernel(T *__restrict__ S, int64_t F, int64_t P, int64_t ldp, const T *__restrict__ W,
                                     const double2 *__restrict__ coef) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = blockIdx.z;
    if (p >= P) return;
    const double2 ac = coef[b * P + p];
    for (int64_t f = blockIdx.y; f < F; f += gridDim.y) {
        const double w0r = (double)W[2 * f], w0i = (double)W[2 * f + 1];
        const double w1r = (double)W[2 * (F + f)], w1i = (double)W[2 * (F + f) + 1];
        T *q = S + 2 * ((b * F + f) * ldp + p);
        q[0] = (T)((double)q[0] - (ac.x * w1r + ac.y * w0r));
        q[1] = (T)((double)q[1] - (ac.x * w1i + ac.y * w0i));
    }
}

// Sxy[b][f][p] = Sx * conj(Sy), in place over Sx (scipy ShortTimeFFT.spectrogram(x, y), :1401-1404)
template <typename T>
__global__ void cross_kernel(T *__restrict__ Sx, const T *__restrict__ Sy, int64_t rows, int64_t P, int64_t ldx, int64_t ldy) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        T *a = Sx + 2 * (r * ldx + p);
        const T *b = Sy + 2 * (r * ldy + p);
        const T ar = a[0], ai = a[1], br = b[0], bi = b[1];
        a[0] = ar * br + ai * bi;
        a[1] = ai * br - ar * bi;
    }
}

// op 1: A += i B (complex-input stft of a two-sided layout = stft(Re x) + i stft(Im x), the transform is linear);
// op 2: A = -i B (Im(istft(S)) = Re(istft(-i S)): the imaginary part of the twin's complex istft)
template <typename T>
__global__ void complex_op_kernel(T *__restrict__ A, const T *__restrict__ B, int64_t rows, int64_t P, int64_t lda,
                                  int64_t ldb, int op) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        T *a = A + 2 * (r * lda + p);
        const T *b = B + 2 * (r * ldb + p);
        const T br = b[0], bi = b[1];
        if (op == 1) { a[0] -= bi; a[1] += br; }
        else { a[0] = bi; a[1] = -br; }
    }
}

}  // namespace

template <typename T>
cudaError_t complex_op(T *A, const T *B, int64_t rows, int64_t P, int64_t lda, int64_t ldb, int op, cudaStream_t st) {
    if (rows <= 0 || P <= 0) return cudaSuccess;
    const int threads = 128;
    dim3 grid((unsigned)((P + threads - 1) / threads), (unsigned)(rows < 4096 ? rows : 4096));
    complex_op_kernel<T><<<grid, threads, 0, st>>>(A, B, rows, P, lda, ldb, op);
    ++g_launch_count;
    return cudaGetLastError();
}
template cudaError_t complex_op<float>(float *, const float *, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);
template cudaError_t complex_op<double>(double *, const double *, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);

template <typename T>
cudaError_t cross_multiply(T *Sx, const T *Sy, int64_t rows, int64_t P, int64_t ldx, int64_t ldy, cudaStream_t st) {
    if (rows <= 0 || P <= 0) return cudaSuccess;
    const int threads = 128;
    dim3 grid((unsigned)((P + threads - 1) / threads), (unsigned)(rows < 4096 ? rows : 4096));
    cross_kernel<T><<<grid, threads, 0, st>>>(Sx, Sy, rows, P, ldx, ldy);
    ++g_launch_count;
    return cudaGetLastError();
}
template cudaError_t cross_multiply<float>(float *, const float *, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
template cudaError_t cross_multiply<double>(double *, const double *, int64_t, int64_t, int64_t, int64_t, cudaStream_t);

template <typename T>
cudaError_t detrend_correct(const T *x, int64_t n, int64_t batch, int64_t x_pitch, int64_t p0, int64_t P, int64_t H,
                            int64_t N, int64_t start0, int linear, const T *W, double2 *coef, T *S, int64_t F,
                            int64_t ldp, cudaStream_t st) {
    if (batch <= 0 || P <= 0) return cudaSuccess;
    if (batch > 65535) return cudaErrorInvalidValue;
    {
        const int wpb = 8;
        dim3 grid((unsigned)((P + wpb - 1) / wpb), (unsigned)batch);
        detrend_coef_kernel<T><<<grid, wpb * 32, 0, st>>>(x, n, x_pitch, p0, P, H, N, start0, linear, coef);
        ++g_launch_count;
    }
    {
        const int threads = 128;
        dim3 grid((unsigned)((P + threads - 1) / threads), (unsigned)(F < 64 ? F : 64), (unsigned)batch);
        detrend_apply_kernel<T><<<grid, threads, 0, st>>>(S, F, P, ldp, W, coef);
        ++g_launch_count;
    }
    return cudaGetLastError();
}
template cudaError_t detrend_correct<float>(const float *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                                            int64_t, int, const float *, double2 *, float *, int64_t, int64_t, cudaStream_t);
template cudaError_t detrend_correct<double>(const double *, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                                             int64_t, int, const double *, double2 *, double *, int64_t, int64_t, cudaStream_t);

}  // namespace veils
