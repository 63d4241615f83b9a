// pow2_tile.cu -- CUDA kernels and launchers of the power-of-two "slice-lane" family.
// The per-phase device code lives in pow2_tile_impl.cuh (shared with the host emulation).
#include "pow2_cfg.hpp"

namespace veils {

namespace {

using namespace p2;

// two CTAs per SM whenever the work buffer (Mc * TF complex) allows it
template <typename T, int LOGMC, int TF>
constexpr int min_ctas() { return sizeof(T) * 2 * (1 << LOGMC) * TF <= 80 * 1024 ? 2 : 1; }

template <typename T, int LOGMC, int TF, int NSUB>
__global__ void __launch_bounds__(TF *NSUB, min_ctas<T, LOGMC, TF>())
pow2_fwd_kernel(const __grid_constant__ FwdArgs<T> args) {
    extern __shared__ __align__(16) unsigned char smem[];
    using K = Fwd<T, LOGMC, TF, NSUB>;
    FwdCtx<T> c;
    c.a = args;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.xin = reinterpret_cast<T *>(smem + sizeof(cx<T>) * K::MC * TF);
    c.b = blockIdx.y;
    c.pt = blockIdx.x;
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
__global__ void __launch_bounds__(TF *NSUB, min_ctas<T, LOGMC, TF>())
pow2_inv_kernel(const __grid_constant__ InvArgs<T> args) {
    extern __shared__ __align__(16) unsigned char smem[];
    using K = Inv<T, LOGMC, TF, NSUB>;
    InvCtx<T> c;
    c.a = args;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.ola = reinterpret_cast<T *>(smem);
    c.b = blockIdx.y;
    c.qa = args.qa0 + (int64_t)blockIdx.x * (TF - args.halo);
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

// batch is mapped to gridDim.y (<= 65535 per launch); larger batches are launched in slabs
struct FwdLaunch {
    const void *x; void *out; const FwdParams *prm; const void *win; const void *tw; cudaStream_t st;
    template <typename T, int LOGMC, int TF, int NSUB>
    cudaError_t run() {
        const FwdParams &p = *prm;
        FwdArgs<T> a = make_fwd_args<T>((const T *)x, out, p, (const T *)win, (const T *)tw, TF);
        const size_t smem = sizeof(cx<T>) * (size_t)(1 << LOGMC) * TF +
                            sizeof(T) * (size_t)fwd_xin_elems(p.N, p.H, TF, a.lay);
        auto kern = pow2_fwd_kernel<T, LOGMC, TF, NSUB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (a.ptiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const size_t osz = p.spectrogram ? sizeof(T) : sizeof(cx<T>);
        for (int64_t b0 = 0; b0 < p.batch; b0 += 65535) {
            const int64_t nb = p.batch - b0 < 65535 ? p.batch - b0 : 65535;
            FwdArgs<T> as = a;
            as.x = a.x + b0 * p.x_pitch;
            as.out = (char *)out + (size_t)b0 * p.F * p.ldp * osz;
            kern<<<dim3((unsigned)a.ptiles, (unsigned)nb), TF * NSUB, smem, st>>>(as);
            ++g_launch_count;
        }
        return cudaGetLastError();
    }
};

struct InvLaunch {
    const void *S; void *xo; const InvParams *prm; const void *dual; const void *tw; cudaStream_t st;
    template <typename T, int LOGMC, int TF, int NSUB>
    cudaError_t run() {
        const InvParams &p = *prm;
        if (TF - inv_halo(p.N, p.H) < 1) return cudaErrorInvalidConfiguration;
        InvArgs<T> a = make_inv_args<T>(S, (T *)xo, p, (const T *)dual, (const T *)tw, TF);
        size_t smem = sizeof(cx<T>) * (size_t)(1 << LOGMC) * TF;
        const size_t ola = sizeof(T) * (size_t)TF * a.ostride;
        if (ola > smem) smem = ola;
        auto kern = pow2_inv_kernel<T, LOGMC, TF, NSUB>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (a.ntiles > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        for (int64_t b0 = 0; b0 < p.batch; b0 += 65535) {
            const int64_t nb = p.batch - b0 < 65535 ? p.batch - b0 : 65535;
            InvArgs<T> as = a;
            as.S = (const char *)S + (size_t)b0 * p.F * p.ldp * sizeof(cx<T>);
            as.xo = a.xo + b0 * p.out_pitch;
            kern<<<dim3((unsigned)a.ntiles, (unsigned)nb), TF * NSUB, smem, st>>>(as);
            ++g_launch_count;
        }
        return cudaGetLastError();
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
