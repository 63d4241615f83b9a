// smooth.cu -- CUDA kernels and launchers of the mixed-radix family (phase code: smooth_impl.cuh).
#include "smooth_impl.cuh"
#include "launch_cache.cuh"

namespace veils {

namespace {
using namespace sm;

constexpr int SM_NT = 256;
constexpr size_t SM_SMEM_LIMIT = 227 * 1024;

template <typename T, int TF>
__global__ void __launch_bounds__(SM_NT) smooth_fwd_kernel(const __grid_constant__ FwdArgs<T> args, const int64_t total) {
    extern __shared__ __align__(16) unsigned char smem[];
    FwdCtx<T> c;
    c.a = &args;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.xin = reinterpret_cast<T *>(c.work + (size_t)args.s.M * TF);
    const int tid = threadIdx.x;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        c.b = w / args.ptiles;
        c.pt = w - c.b * args.ptiles;
        Fwd<T, TF>::stage(c, tid, SM_NT);
        __syncthreads();
        Fwd<T, TF>::load(c, tid, SM_NT);
        __syncthreads();
        for (int p = 0; p < args.s.np; ++p) {
            Fwd<T, TF>::pass(c, p, tid, SM_NT);
            __syncthreads();
        }
        Fwd<T, TF>::store(c, tid, SM_NT);
        __syncthreads();
    }
}

template <typename T, int TF>
__global__ void __launch_bounds__(SM_NT) smooth_inv_kernel(const __grid_constant__ InvArgs<T> args, const int64_t total) {
    extern __shared__ __align__(16) unsigned char smem[];
    InvCtx<T> c;
    c.a = &args;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.ola = reinterpret_cast<T *>(c.work + (size_t)args.s.M * TF);
    const int tid = threadIdx.x;
    const int own = TF - args.halo;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        c.b = w / args.ntiles;
        c.qa = args.qa0 + (w - c.b * args.ntiles) * own;
        Inv<T, TF>::load(c, tid, SM_NT);
        __syncthreads();
        for (int p = 0; p < args.s.np; ++p) {
            Inv<T, TF>::pass(c, p, tid, SM_NT);
            __syncthreads();
        }
        Inv<T, TF>::rows(c, tid, SM_NT);
        __syncthreads();
        Inv<T, TF>::gather(c, tid, SM_NT);
        __syncthreads();
    }
}

template <typename T> size_t fwd_smem(int64_t M, int64_t N, int64_t H, int tf) {
    const int64_t span = (int64_t)(tf - 1) * H + N;
    return sizeof(cx<T>) * (size_t)M * tf + sizeof(T) * (size_t)(span + span / H + 8);
}
template <typename T> size_t inv_smem(int64_t M, int64_t N, int tf) {
    return sizeof(cx<T>) * (size_t)M * tf + sizeof(T) * (size_t)tf * (size_t)((N | 1) + 1);
}
inline int halo_of(int64_t N, int64_t H) { return (int)((N + H - 1) / H) - 1; }
// slices per tile: the widest of {32, 16, 8, 4} that leaves room for two CTAs per SM, else whatever fits
template <typename T> int pick_tf(int64_t M, int64_t N, int64_t H, bool inverse) {
    const int cand[4] = {sizeof(T) == 4 ? 32 : 16, sizeof(T) == 4 ? 16 : 8, sizeof(T) == 4 ? 8 : 4, 4};
    int fits = 0;
    for (int i = 0; i < 4; ++i) {
        const int tf = cand[i];
        if (inverse && halo_of(N, H) >= tf) continue;
        const size_t b = inverse ? inv_smem<T>(M, N, tf) : fwd_smem<T>(M, N, H, tf);
        if (b <= 110 * 1024) return tf;
        if (!fits && b <= SM_SMEM_LIMIT) fits = tf;
    }
    return fits;
}

template <typename T, int TF>
cudaError_t launch_fwd(FwdArgs<T> &a, int64_t batch, size_t smem, cudaStream_t st) {
    auto kern = smooth_fwd_kernel<T, TF>;
    static KernelSlot slot;
    int per_sm = 0, dev = 0;
    cudaError_t e = prepare_kernel(slot, kern, SM_NT, smem, &per_sm, &dev);
    if (e != cudaSuccess) return e;
    a.ptiles = (a.P + TF - 1) / TF;
    const int64_t total = a.ptiles * batch, cap = (int64_t)per_sm * device_sms(dev);
    kern<<<(unsigned)(total < cap ? total : cap), SM_NT, smem, st>>>(a, total);
    ++g_launch_count;
    return cudaGetLastError();
}
template <typename T, int TF>
cudaError_t launch_inv(InvArgs<T> &a, int64_t batch, size_t smem, cudaStream_t st) {
    auto kern = smooth_inv_kernel<T, TF>;
    static KernelSlot slot;
    int per_sm = 0, dev = 0;
    cudaError_t e = prepare_kernel(slot, kern, SM_NT, smem, &per_sm, &dev);
    if (e != cudaSuccess) return e;
    const int own = TF - a.halo;
    const int64_t qh0 = p2::fdiv64(a.k0 + a.mid, a.s.H), qh1 = p2::fdiv64(a.k1 - 1 + a.mid, a.s.H);
    a.ntiles = (qh1 - qh0 + 1 + own - 1) / own;
    a.qa0 = qh0 - a.halo;
    const int64_t total = a.ntiles * batch, cap = (int64_t)per_sm * device_sms(dev);
    kern<<<(unsigned)(total < cap ? total : cap), SM_NT, smem, st>>>(a, total);
    ++g_launch_count;
    return cudaGetLastError();
}
}  // namespace

bool smooth_supported(int64_t M, int64_t N, int64_t H, int precision, bool inverse) {
    sm::Shape s;
    if (N > M || H < 1 || !sm::make_shape(M, N, H, M, 0, &s)) return false;
    if (inverse && H > N) return false;
    return (precision == 0 ? pick_tf<float>(M, N, H, inverse) : pick_tf<double>(M, N, H, inverse)) > 0;
}
int64_t smooth_twiddle_count(int64_t M) {
    sm::Shape s;
    if (!sm::make_shape(M, M, 1, M, 0, &s)) return 0;
    return 2 * sm::twiddle_entries(s);
}
void smooth_twiddle_fill(int64_t M, double *re_im) {
    sm::Shape s;
    if (sm::make_shape(M, M, 1, M, 0, &s)) sm::twiddle_fill(s, re_im);
}

template <typename T>
cudaError_t smooth_forward(const T *x, void *out, const FwdParams &p, const T *win, const T *tw, cudaStream_t st) {
    if (p.P <= 0 || p.batch <= 0) return cudaSuccess;
    sm::FwdArgs<T> a;
    if (!sm::make_shape(p.M, p.N, p.H, p.F, p.ps, &a.s)) return cudaErrorNotSupported;
    a.x = x; a.out = out; a.win = win; a.tw = reinterpret_cast<const sm::cx<T> *>(tw);
    a.n = p.n; a.x_pitch = p.x_pitch; a.ldp = p.ldp; a.p0 = p.p0; a.P = p.P; a.k_offset = p.k_offset; a.ptiles = 0;
    a.mid = (int)p.mid; a.mode = p.mode; a.spectrogram = p.spectrogram; a.xpad = (p.H % 2 == 0) ? 1 : 0;
    a.fac2x = p.mode == 3 ? (T)p.fac2x : (T)1;
    const int tf = pick_tf<T>(p.M, p.N, p.H, false);
    const size_t smem = fwd_smem<T>(p.M, p.N, p.H, tf);
    switch (tf) {
        case 32: return launch_fwd<T, 32>(a, p.batch, smem, st);
        case 16: return launch_fwd<T, 16>(a, p.batch, smem, st);
        case 8: return launch_fwd<T, 8>(a, p.batch, smem, st);
        case 4: return launch_fwd<T, 4>(a, p.batch, smem, st);
    }
    return cudaErrorNotSupported;
}
template <typename T>
cudaError_t smooth_inverse(const void *S, T *x_out, const InvParams &p, const T *dual_over_m, const T *tw, cudaStream_t st) {
    if (p.batch <= 0 || p.k1 <= p.k0) return cudaSuccess;
    sm::InvArgs<T> a;
    if (!sm::make_shape(p.M, p.N, p.H, p.F, p.ps, &a.s)) return cudaErrorNotSupported;
    a.S = S; a.xo = x_out; a.dual = dual_over_m; a.tw = reinterpret_cast<const sm::cx<T> *>(tw);
    a.P = p.P; a.ldp = p.ldp; a.out_pitch = p.out_pitch; a.p_min = p.p_min;
    a.k0 = p.k0; a.k1 = p.k1; a.q0 = p.q0; a.q1 = p.q1; a.qa0 = 0; a.ntiles = 0;
    a.mid = (int)p.mid; a.mode = p.mode; a.halo = halo_of(p.N, p.H); a.ostride = (int)((p.N | 1) + 0);
    a.rfac2x = p.mode == 3 ? (T)(1.0 / p.fac2x) : (T)1;
    const int tf = pick_tf<T>(p.M, p.N, p.H, true);
    const size_t smem = inv_smem<T>(p.M, p.N, tf);
    switch (tf) {
        case 32: return launch_inv<T, 32>(a, p.batch, smem, st);
        case 16: return launch_inv<T, 16>(a, p.batch, smem, st);
        case 8: return launch_inv<T, 8>(a, p.batch, smem, st);
        case 4: return launch_inv<T, 4>(a, p.batch, smem, st);
    }
    return cudaErrorNotSupported;
}
template cudaError_t smooth_forward<float>(const float *, void *, const FwdParams &, const float *, const float *, cudaStream_t);
template cudaError_t smooth_forward<double>(const double *, void *, const FwdParams &, const double *, const double *, cudaStream_t);
template cudaError_t smooth_inverse<float>(const void *, float *, const InvParams &, const float *, const float *, cudaStream_t);
template cudaError_t smooth_inverse<double>(const void *, double *, const InvParams &, const double *, const double *, cudaStream_t);

}  // namespace veils
