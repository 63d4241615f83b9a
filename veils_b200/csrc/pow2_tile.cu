// pow2_tile.cu -- CUDA kernels and launchers of the power-of-two "slice-lane" family.
// The per-phase device code lives in pow2_tile_impl.cuh (shared with the host emulation).
#include "pow2_cfg.hpp"

namespace veils {

namespace {

using namespace p2;

template <typename T, int LOGMC, int TF, int NSUB>
__global__ void __launch_bounds__(TF *NSUB)
pow2_fwd_kernel(const T *__restrict__ x, void *__restrict__ out, FwdParams prm,
                const T *__restrict__ win, const T *__restrict__ tw, TileLayout lay, int vec) {
    extern __shared__ __align__(16) unsigned char smem[];
    using K = Fwd<T, LOGMC, TF, NSUB>;
    FwdCtx<T> c;
    c.x = x; c.out = out; c.prm = prm; c.win = win; c.tw = tw;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.xin = reinterpret_cast<T *>(smem + sizeof(cx<T>) * K::MC * TF);
    const int64_t ptiles = (prm.P + TF - 1) / TF;
    c.b = blockIdx.x / ptiles;
    c.pt = blockIdx.x % ptiles;
    c.lay = lay;
    c.vec = vec;
    const int tid = threadIdx.x;
    K::fill(c, tid);
    __syncthreads();
    K::pass1(c, tid);
    __syncthreads();
    if (K::NP == 3) {
        K::pass_mid(c, tid);
        __syncthreads();
    }
    K::pass_last(c, tid);
}

template <typename T, int LOGMC, int TF, int NSUB>
__global__ void __launch_bounds__(TF *NSUB)
pow2_inv_kernel(const void *__restrict__ S, T *__restrict__ xo, InvParams prm,
                const T *__restrict__ dual, const T *__restrict__ tw, int64_t qa0, int64_t ntiles,
                int halo, int ostride) {
    extern __shared__ __align__(16) unsigned char smem[];
    using K = Inv<T, LOGMC, TF, NSUB>;
    InvCtx<T> c;
    c.S = S; c.xo = xo; c.prm = prm; c.dual = dual; c.tw = tw;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.ola = reinterpret_cast<T *>(smem);
    c.b = blockIdx.x / ntiles;
    c.qa = qa0 + (int64_t)(blockIdx.x % ntiles) * (TF - halo);
    c.halo = halo;
    c.ostride = ostride;
    const int tid = threadIdx.x;
    K::pass1(c, tid);
    __syncthreads();
    if (K::NP == 3) {
        K::pass_mid(c, tid);
        __syncthreads();
    }
    cx<T> r[K::GPT][K::RL];
    K::last_load(c, tid, r);
    __syncthreads();                       // all reads of `work` done before it becomes `ola`
    K::last_store(c, tid, r);
    __syncthreads();
    K::gather(c, tid);
}

struct FwdLaunch {
    const void *x; void *out; const FwdParams *prm; const void *win; const void *tw; cudaStream_t st;
    template <typename T, int LOGMC, int TF, int NSUB>
    cudaError_t run() {
        const FwdParams &p = *prm;
        const bool vec = fwd_vec(p.ps, p.H, p.N);
        const TileLayout lay = tile_layout(p.H, vec);
        const size_t smem = sizeof(cx<T>) * (size_t)(1 << LOGMC) * TF +
                            sizeof(T) * (size_t)fwd_xin_elems(p.N, p.H, TF, lay);
        auto kern = pow2_fwd_kernel<T, LOGMC, TF, NSUB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t ptiles = (p.P + TF - 1) / TF;
        const int64_t blocks = ptiles * p.batch;
        if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        kern<<<(unsigned)blocks, TF * NSUB, smem, st>>>((const T *)x, out, p, (const T *)win,
                                                        (const T *)tw, lay, vec ? 1 : 0);
        ++g_launch_count;
        return cudaGetLastError();
    }
};

struct InvLaunch {
    const void *S; void *xo; const InvParams *prm; const void *dual; const void *tw; cudaStream_t st;
    template <typename T, int LOGMC, int TF, int NSUB>
    cudaError_t run() {
        const InvParams &p = *prm;
        const int halo = inv_halo(p.N, p.H);
        const int own = TF - halo;
        if (own < 1) return cudaErrorInvalidConfiguration;
        const int64_t qh0 = floordiv_i64(p.k0 + p.mid, p.H);
        const int64_t qh1 = floordiv_i64(p.k1 - 1 + p.mid, p.H);
        const int64_t ntiles = (qh1 - qh0 + 1 + own - 1) / own;
        const int ostride = inv_ostride(p.N);
        size_t smem = sizeof(cx<T>) * (size_t)(1 << LOGMC) * TF;
        const size_t ola = sizeof(T) * (size_t)TF * ostride;
        if (ola > smem) smem = ola;
        auto kern = pow2_inv_kernel<T, LOGMC, TF, NSUB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t blocks = ntiles * p.batch;
        if (blocks > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        kern<<<(unsigned)blocks, TF * NSUB, smem, st>>>(S, (T *)xo, p, (const T *)dual, (const T *)tw,
                                                        qh0 - halo, ntiles, halo, ostride);
        ++g_launch_count;
        return cudaGetLastError();
    }
    static int64_t floordiv_i64(int64_t a, int64_t b) {
        int64_t q = a / b;
        return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
    }
};

}  // namespace

bool pow2_supported(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse) {
    return p2::supported(M, N, H, ps, precision, inverse);
}
int64_t pow2_twiddle_count(int64_t M) { return p2::twiddle_count(M); }
void pow2_twiddle_fill(int64_t M, double *t) { p2::twiddle_fill(M, t); }

template <typename T>
cudaError_t pow2_forward(const T *x, void *out, const FwdParams &prm, const Tables<T> &tb, cudaStream_t st) {
    if (prm.P <= 0 || prm.batch <= 0) return cudaSuccess;
    p2::Cfg c;
    const int prec = sizeof(T) == 4 ? 0 : 1;
    if (!p2::cfg_for(prm.M, prec, &c)) return cudaErrorNotSupported;
    FwdLaunch l{x, out, &prm, tb.win_half, tb.tw_pow2, st};
    return p2::dispatch(prec, c.logmc, l);
}

template <typename T>
cudaError_t pow2_inverse(const void *S, T *x_out, const InvParams &prm, const Tables<T> &tb, cudaStream_t st) {
    if (prm.batch <= 0 || prm.k1 <= prm.k0) return cudaSuccess;
    p2::Cfg c;
    const int prec = sizeof(T) == 4 ? 0 : 1;
    if (!p2::cfg_for(prm.M, prec, &c)) return cudaErrorNotSupported;
    InvLaunch l{S, x_out, &prm, tb.dual_scaled, tb.tw_pow2, st};
    return p2::dispatch(prec, c.logmc, l);
}

template cudaError_t pow2_forward<float>(const float *, void *, const FwdParams &, const Tables<float> &, cudaStream_t);
template cudaError_t pow2_forward<double>(const double *, void *, const FwdParams &, const Tables<double> &, cudaStream_t);
template cudaError_t pow2_inverse<float>(const void *, float *, const InvParams &, const Tables<float> &, cudaStream_t);
template cudaError_t pow2_inverse<double>(const void *, double *, const InvParams &, const Tables<double> &, cudaStream_t);

}  // namespace veils
