// dft_generic.cu -- correctness-first kernel family for ANY mfft (odd, prime, mfft > m_num ...).
//
// The reference accepts arbitrary FFT lengths (rustfft plans any size; its own golden vector
// uses mfft = m_num = 15, examples/basic_usage.rs:150-160).  This family evaluates the DFT sums
// directly on the GPU -- O(M^2) per slice, f64 accumulation, twiddles from an exact f64 table --
// so that every configuration the reference accepts has a CUDA path (there is no CPU fallback).
// Performance-critical power-of-two sizes dispatch to pow2_tile.cu instead.
//
// Index algebra: SURVEY.md appendix A (derived from src/lib.rs:336-411, 414-505, 612-652, 853-911).
#include "kernels.cuh"

namespace veils {

std::atomic<unsigned long long> g_launch_count{0};

namespace {

constexpr int TP = 32;     // slices per block (x)   -> stores of 32 consecutive p are coalesced
constexpr int TK = 8;      // bins / samples per block (y)
constexpr int CH = 64;     // reduction chunk staged in shared memory

template <typename T> struct Cplx;
template <> struct Cplx<float> { using type = float2; };
template <> struct Cplx<double> { using type = double2; };

// Forward: S[b][kk][p] = epilogue( sum_i z'[i] * exp(-2 pi i * i*k / M) ),
// z'[i] = win[j] * x[p*H + k_offset - mid + j] (0 outside [0,n)),  j = (i + ps) mod M, j < N.
template <typename T>
__global__ void __launch_bounds__(TP *TK)
dft_forward_kernel(const T *__restrict__ x, void *__restrict__ out, FwdParams prm,
                   const T *__restrict__ win, const double2 *__restrict__ tw) {
    __shared__ double zt[CH][TP + 1];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TP + tx;
    const int64_t ptiles = (prm.P + TP - 1) / TP;
    const int64_t b = blockIdx.x / ptiles;
    const int64_t pt = blockIdx.x % ptiles;
    const int64_t M = prm.M, N = prm.N;
    const int64_t kk = (int64_t)blockIdx.y * TK + ty;          // output row
    // bin computed for this output row (fftshift for centered: src/lib.rs:508-513)
    int64_t k = kk;
    if (prm.mode == 1) k = (kk + (M + 1) / 2) % M;
    const bool krow = kk < prm.F;
    const int64_t pl = pt * TP + tx;                            // slice index relative to p0
    const T *xb = x + b * prm.x_pitch;

    double ar = 0.0, ai = 0.0;
    for (int64_t i0 = 0; i0 < M; i0 += CH) {
        // stage CH windowed, rolled samples of TP slices
        for (int e = tid; e < CH * TP; e += TP * TK) {
            const int ii = e % CH, pp = e / CH;
            const int64_t i = i0 + ii;
            const int64_t pq = pt * TP + pp;
            double v = 0.0;
            if (i < M && pq < prm.P) {
                int64_t j = i + prm.ps;
                if (j >= M) j -= M;
                if (j < N) {
                    const int64_t s = (prm.p0 + pq) * prm.H + prm.k_offset - prm.mid + j;
                    if (s >= 0 && s < prm.n) v = (double)xb[s] * (double)win[j];
                }
            }
            zt[ii][pp] = v;
        }
        __syncthreads();
        if (krow) {
            int64_t idx = (int64_t)(((unsigned long long)i0 * (unsigned long long)k) % (unsigned long long)M);
            const int lim = (int)min((int64_t)CH, M - i0);
            for (int ii = 0; ii < lim; ++ii) {
                const double2 w = tw[idx];
                const double v = zt[ii][tx];
                ar = fma(v, w.x, ar);
                ai = fma(v, w.y, ai);
                idx += k;
                if (idx >= M) idx -= M;
            }
        }
        __syncthreads();
    }
    if (!krow || pl >= prm.P) return;
    // mode epilogue (src/lib.rs:367-408)
    if (prm.mode >= 2) {
        if (k == 0 || (M % 2 == 0 && k == M / 2)) ai = 0.0;
        if (prm.mode == 3) {
            const int64_t last = (M % 2 == 0) ? prm.F - 2 : prm.F - 1;
            if (k >= 1 && k <= last) { ar *= prm.fac2x; ai *= prm.fac2x; }
        }
    }
    const int64_t o = (b * prm.F + kk) * prm.ldp + pl;
    if (prm.spectrogram) {
        // spectrogram must equal |round(S)|^2 of the stored-precision S (scipy: Sx.real**2+Sx.imag**2)
        const T r = (T)ar, im = (T)ai;
        ((T *)out)[o] = r * r + im * im;
    } else {
        typename Cplx<T>::type v;
        v.x = (T)ar;
        v.y = (T)ai;
        ((typename Cplx<T>::type *)out)[o] = v;
    }
}

// Inverse, stage A: xs[b][q-q0][j] = Re( IDFT_M(X)[(j - ps) mod M] ) / M * dual[j]
// (src/lib.rs:414-505 and :869-873).  Threads: x = slice q, y = sample j.
template <typename T>
__global__ void __launch_bounds__(TP *TK)
dft_inverse_frames_kernel(const void *__restrict__ Sv, double *__restrict__ ws, InvParams prm,
                          int64_t b0, const T *__restrict__ dual, const double2 *__restrict__ tw) {
    using C = typename Cplx<T>::type;
    __shared__ double sr[CH][TP + 1];
    __shared__ double si[CH][TP + 1];
    const C *S = (const C *)Sv;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TP + tx;
    const int64_t Q = prm.q1 - prm.q0;
    const int64_t qtiles = (Q + TP - 1) / TP;
    const int64_t bl = blockIdx.x / qtiles;                     // batch index within the chunk
    const int64_t b = b0 + bl;
    const int64_t qt = blockIdx.x % qtiles;
    const int64_t M = prm.M, F = prm.F;
    const int64_t j = (int64_t)blockIdx.y * TK + ty;
    const bool jrow = j < prm.N;
    int64_t i = jrow ? (j - prm.ps) % M : 0;                    // roll right by ps over M
    if (i < 0) i += M;
    const bool onesided = prm.mode >= 2;
    const int64_t last2x = (M % 2 == 0) ? F - 2 : F - 1;        // paired bins 1..last2x

    double acc = 0.0;
    for (int64_t c0 = 0; c0 < F; c0 += CH) {
        for (int e = tid; e < CH * TP; e += TP * TK) {
            const int qq = e % TP, kl = e / TP;
            const int64_t kk = c0 + kl;
            const int64_t q = prm.q0 + qt * TP + qq;
            const int64_t col = q - prm.p_min;
            double vr = 0.0, vi = 0.0;
            if (kk < F && q < prm.q1 && col >= 0 && col < prm.P) {
                const C v = S[(b * F + kk) * prm.ldp + col];
                vr = (double)v.x;
                vi = (double)v.y;
            }
            sr[kl][qq] = vr;
            si[kl][qq] = vi;
        }
        __syncthreads();
        if (jrow) {
            const int lim = (int)min((int64_t)CH, F - c0);
            for (int kl = 0; kl < lim; ++kl) {
                const int64_t kk = c0 + kl;
                int64_t k = kk;                                 // DFT bin held by row kk
                if (prm.mode == 1) { k = kk + (M + 1) / 2; if (k >= M) k -= M; }
                double wgt = 1.0;
                double vr = sr[kl][tx], vi = si[kl][tx];
                if (onesided) {
                    // irfft semantics: imag of DC / Nyquist ignored, paired bins counted twice
                    if (k == 0 || (M % 2 == 0 && k == M / 2)) vi = 0.0;
                    else wgt = 2.0;
                    if (prm.mode == 3 && k >= 1 && k <= last2x) wgt /= prm.fac2x;
                }
                const int64_t idx = (int64_t)(((unsigned long long)i * (unsigned long long)k) % (unsigned long long)M);
                const double2 w = tw[idx];                      // (cos, -sin)
                acc += wgt * (vr * w.x + vi * w.y);             // Re(X * exp(+i theta))
            }
        }
        __syncthreads();
    }
    const int64_t ql = qt * TP + tx;
    if (jrow && ql < Q) ws[(bl * Q + ql) * prm.N + j] = acc / (double)M * (double)dual[j];
}

// Inverse, stage B: deterministic overlap-add as a gather, increasing q (src/lib.rs:853-911).
template <typename T>
__global__ void dft_inverse_ola_kernel(const double *__restrict__ ws, T *__restrict__ xo,
                                       InvParams prm, int64_t b0, int64_t nb) {
    const int64_t len = prm.k1 - prm.k0;
    const int64_t Q = prm.q1 - prm.q0;
    const int64_t total = len * nb;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t bl = e / len, o = e % len;
        const int64_t k = prm.k0 + o;
        // q with 0 <= k - (q*H - mid) < N
        const int64_t a = k + prm.mid;
        int64_t qlo = a - prm.N + 1;                            // ceil(qlo / H)
        qlo = (qlo >= 0) ? (qlo + prm.H - 1) / prm.H : -((-qlo) / prm.H);
        int64_t qhi = (a >= 0) ? a / prm.H : -((-a + prm.H - 1) / prm.H);   // floor(a / H)
        if (qlo < prm.q0) qlo = prm.q0;
        if (qhi > prm.q1 - 1) qhi = prm.q1 - 1;
        double acc = 0.0;
        for (int64_t q = qlo; q <= qhi; ++q) {
            const int64_t col = q - prm.p_min;
            if (col < 0 || col >= prm.P) continue;
            acc += ws[(bl * Q + (q - prm.q0)) * prm.N + (a - q * prm.H)];
        }
        xo[(b0 + bl) * prm.out_pitch + o] = (T)acc;
    }
}

}  // namespace

template <typename T>
cudaError_t dft_generic_forward(const T *x, void *out, const FwdParams &prm, const Tables<T> &tb,
                                cudaStream_t st) {
    if (prm.P <= 0 || prm.batch <= 0) return cudaSuccess;
    const int64_t ptiles = (prm.P + TP - 1) / TP;
    const int64_t ktiles = (prm.F + TK - 1) / TK;
    if (ktiles > 65535 || ptiles * prm.batch > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    dim3 grid((unsigned)(ptiles * prm.batch), (unsigned)ktiles), block(TP, TK);
    dft_forward_kernel<T><<<grid, block, 0, st>>>(x, out, prm, tb.win, tb.tw_full);
    ++g_launch_count;
    return cudaGetLastError();
}

template <typename T>
cudaError_t dft_generic_inverse(const void *S, T *x_out, const InvParams &prm, const Tables<T> &tb,
                                cudaStream_t st) {
    if (prm.batch <= 0 || prm.k1 <= prm.k0) return cudaSuccess;
    const int64_t Q = prm.q1 - prm.q0;
    const int64_t len = prm.k1 - prm.k0;
    if (Q <= 0) {
        return cudaMemset2DAsync(x_out, prm.out_pitch * sizeof(T), 0, len * sizeof(T), prm.batch, st);
    }
    // workspace bounded to ~256 MiB: process the batch in chunks
    const int64_t per_row = Q * prm.N;
    int64_t nb = (int64_t)((256LL << 20) / 8) / (per_row > 0 ? per_row : 1);
    if (nb < 1) nb = 1;
    if (nb > prm.batch) nb = prm.batch;
    double *ws = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&ws, (size_t)(nb * per_row) * sizeof(double), st);
    if (e != cudaSuccess) return e;
    const int64_t qtiles = (Q + TP - 1) / TP;
    const int64_t jtiles = (prm.N + TK - 1) / TK;
    if (jtiles > 65535) { cudaFreeAsync(ws, st); return cudaErrorInvalidConfiguration; }
    for (int64_t b0 = 0; b0 < prm.batch; b0 += nb) {
        const int64_t cur = (prm.batch - b0 < nb) ? prm.batch - b0 : nb;
        dim3 grid((unsigned)(qtiles * cur), (unsigned)jtiles), block(TP, TK);
        dft_inverse_frames_kernel<T><<<grid, block, 0, st>>>(S, ws, prm, b0, tb.dual, tb.tw_full);
        ++g_launch_count;
        const int64_t total = len * cur;
        const int threads = 256;
        int64_t blocks = (total + threads - 1) / threads;
        if (blocks > 148 * 64) blocks = 148 * 64;
        dft_inverse_ola_kernel<T><<<(unsigned)blocks, threads, 0, st>>>(ws, x_out, prm, b0, cur);
        ++g_launch_count;
    }
    e = cudaGetLastError();
    cudaFreeAsync(ws, st);
    return e;
}

template cudaError_t dft_generic_forward<float>(const float *, void *, const FwdParams &,
                                                const Tables<float> &, cudaStream_t);
template cudaError_t dft_generic_forward<double>(const double *, void *, const FwdParams &,
                                                 const Tables<double> &, cudaStream_t);
template cudaError_t dft_generic_inverse<float>(const void *, float *, const InvParams &,
                                                const Tables<float> &, cudaStream_t);
template cudaError_t dft_generic_inverse<double>(const void *, double *, const InvParams &,
                                                 const Tables<double> &, cudaStream_t);

}  // namespace veils
