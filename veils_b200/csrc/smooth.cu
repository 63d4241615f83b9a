// smooth.cu -- CUDA kernels and launchers of the mixed-radix family (phase code: smooth_impl.cuh).
#include "smooth_impl.cuh"
#include "launch_cache.cuh"

namespace veils {

namespace {
using namespace sm;

constexpr int SM_NT = 256;
#ifndef VEILS_SMOOTH_MINB
#define VEILS_SMOOTH_MINB 2
#endif
constexpr size_t SM_SMEM_LIMIT = 227 * 1024;

// cooperative copy of the plan blob into shared memory (16-byte words; the blob is 16-byte padded)
__device__ __forceinline__ void stage_tab(unsigned char *dst, const unsigned char *src, int bytes, int tid) {
    const int4 *s = reinterpret_cast<const int4 *>(src);
    int4 *d = reinterpret_cast<int4 *>(dst);
    for (int i = tid; i < bytes / 16; i += SM_NT) d[i] = __ldg(s + i);
}

// Persistent forward kernel.  Shared memory: work[LC][TF] | staged samples | plan blob.
template <typename T, int TF>
__global__ void __launch_bounds__(SM_NT, VEILS_SMOOTH_MINB) smooth_fwd_kernel(const __grid_constant__ FwdArgs<T> args, const int64_t total,
                                                           const int xin_bytes, const int dbuf) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    FwdCtx<T> c;
    c.a = &args;
    c.work = reinterpret_cast<cx<vec_t<T>> *>(smem);
    unsigned char *q = smem + sizeof(cx<T>) * (size_t)args.s.LC * TF;
    // the two staging buffers are addressed by arithmetic on the shared-memory base (a pointer picked from an array
    // would be a generic pointer: LD instead of LDS in pass 1)
    T *const xin0 = reinterpret_cast<T *>(q);
    const int xin_elems = xin_bytes / (int)sizeof(T);
    // dbuf: two staging buffers (the next tile is staged while this one is transformed); else one buffer, re-filled
    // as soon as pass 1 has read it (the wider tile that then fits half an SM is worth more than the earlier prefetch)
    const int nbuf = dbuf ? 2 : 1;
    c.tab = q + nbuf * xin_bytes;
    stage_tab(q + nbuf * xin_bytes, args.tab, args.lay.bytes, tid);
    // (signal, tile) of work item w, advanced by gridDim.x per iteration without a 64-bit division
    c.b = blockIdx.x / args.ptiles;
    c.pt = blockIdx.x - c.b * args.ptiles;
    const int64_t db = gridDim.x / args.ptiles, dp = gridDim.x - db * args.ptiles;
    if ((int64_t)blockIdx.x < total) {
        c.xin = xin0;
        Fwd<T, TF>::stage(c, tid, SM_NT);
        p2::cp_async_commit();
    }
    int it = 0;
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x, it ^= 1) {
        p2::cp_async_wait_all();
        __syncthreads();                   // this tile's samples (and the tables) are in place
        int64_t nb = c.b + db, npt = c.pt + dp;
        if (npt >= args.ptiles) { npt -= args.ptiles; ++nb; }
        const bool more = w + gridDim.x < total;
        FwdCtx<T> cn = c;
        cn.b = nb; cn.pt = npt; cn.xin = xin0 + (dbuf ? (it ^ 1) * xin_elems : 0);
        if (more && dbuf) {                // stage the next tile into the other buffer while this one is transformed
            Fwd<T, TF>::stage(cn, tid, SM_NT);
            p2::cp_async_commit();
        }
        c.xin = xin0 + (dbuf ? it * xin_elems : 0);
        for (int p = 0; p < args.s.np; ++p) {
            Fwd<T, TF>::pass(c, p, tid, SM_NT);
            __syncthreads();
            if (p == 0 && more && !dbuf) { // single buffer: free once pass 1 has read it
                Fwd<T, TF>::stage(cn, tid, SM_NT);
                p2::cp_async_commit();
            }
        }
        Fwd<T, TF>::store(c, tid, SM_NT);
        c.b = nb;
        c.pt = npt;
    }
}

template <typename T, int TF> __host__ __device__ constexpr int nyq_bytes() { return (int)((TF * sizeof(cx<T>) + 15) & ~(size_t)15); }

// Persistent inverse kernel.  Shared memory: work[LC][TF] | slice rows [TF][ostride] | Nyquist row | plan blob.
template <typename T, int TF>
__global__ void __launch_bounds__(SM_NT, VEILS_SMOOTH_MINB) smooth_inv_kernel(const __grid_constant__ InvArgs<T> args, const int64_t total,
                                                           const int ola_bytes) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x;
    InvCtx<T> c;
    c.a = &args;
    c.work = reinterpret_cast<cx<vec_t<T>> *>(smem);
    unsigned char *q = smem + sizeof(cx<T>) * (size_t)args.s.LC * TF;
    c.ola = reinterpret_cast<T *>(q);
    c.nyq = reinterpret_cast<cx<T> *>(q + ola_bytes);
    c.tab = q + ola_bytes + nyq_bytes<T, TF>();
    stage_tab(q + ola_bytes + nyq_bytes<T, TF>(), args.tab, args.lay.bytes, tid);
    const int own = TF - args.halo;
    c.b = blockIdx.x / args.ntiles;
    int64_t ti = blockIdx.x - c.b * args.ntiles;
    const int64_t db = gridDim.x / args.ntiles, dp = gridDim.x - db * args.ntiles;
    const bool lands = Inv<T, TF>::lands(args);
    if (lands && (int64_t)blockIdx.x < total) {
        c.qa = args.qa0 + ti * own;
        Inv<T, TF>::land(c, tid, SM_NT);
        p2::cp_async_commit();
    }
    __syncthreads();
    for (int64_t w = blockIdx.x; w < total; w += gridDim.x) {
        c.qa = args.qa0 + ti * own;
        if (lands) {
            p2::cp_async_wait_all();
            __syncthreads();               // the S tile has landed
            Inv<T, TF>::merge(c, tid, SM_NT);
        } else {
            Inv<T, TF>::load(c, tid, SM_NT);
        }
        __syncthreads();
        for (int p = 0; p < args.s.np; ++p) {
            Inv<T, TF>::pass(c, p, tid, SM_NT);
            __syncthreads();
        }
        int64_t nb = c.b + db, nti = ti + dp;
        if (nti >= args.ntiles) { nti -= args.ntiles; ++nb; }
        if (lands && w + gridDim.x < total) {      // the work buffer is free: the next S tile lands during the overlap-add
            InvCtx<T> cn = c;
            cn.b = nb;
            cn.qa = args.qa0 + nti * own;
            Inv<T, TF>::land(cn, tid, SM_NT);
            p2::cp_async_commit();
        }
        Inv<T, TF>::gather(c, tid, SM_NT);
        __syncthreads();
        c.b = nb;
        ti = nti;
    }
}

inline int halo_of(int64_t N, int64_t H) { return (int)((N + H - 1) / H) - 1; }
template <typename T> size_t xin_bytes_of(const Shape &s, int tf) {
    const int64_t span = (int64_t)(tf - 1) * s.H + s.N, rows = (span + s.H - 1) / s.H;
    return (size_t)al16((int64_t)sizeof(T) * (rows * (s.H + xpad_of(s.H)) + 4));
}
template <typename T> size_t ola_bytes_of(const Shape &s, int tf) {
    return (size_t)al16((int64_t)sizeof(T) * tf * ostride_of<T>(s.N, s.H));
}
// shared memory without / with the plan blob
template <typename T> void smem_of(const Shape &s, int tf, bool inverse, size_t *core, size_t *with_tab, bool dbuf = true) {
    const size_t work = sizeof(cx<T>) * (size_t)s.LC * tf;
    *core = work + (inverse ? ola_bytes_of<T>(s, tf) + (size_t)tf * sizeof(cx<T>) + 16   // + landed Nyquist row
                            : (dbuf ? 2 : 1) * xin_bytes_of<T>(s, tf));                   // forward: (double-)buffered staging
    *with_tab = *core + (size_t)tab_layout<T>(s, inverse).bytes;
}
// slices per tile: the wide tile when two CTAs of it fit an SM, else the narrow one, else whatever fits (plan blob included).
// Forward: a tile width is tried with two staging buffers first, then with one (*dbuf).
template <typename T> int pick_tf(const Shape &s, bool inverse, bool *dbuf = nullptr) {
    const int cand[3] = {sizeof(T) == 4 ? 32 : 16, sizeof(T) == 4 ? 16 : 8, sizeof(T) == 4 ? 8 : 4};
    const int force = VEILS_ENV_ONCE("VEILS_SMOOTH_TF", 0);
    const size_t budgets[2] = {113 * 1024, SM_SMEM_LIMIT};
    for (int bi = 0; bi < 2; ++bi)
        for (int i = 0; i < 3; ++i) {
            const int tf = cand[i];
            if (force && tf != force) continue;
            if (inverse && halo_of(s.N, s.H) >= tf) continue;
            if (inverse && bi < 1 && 4 * halo_of(s.N, s.H) > tf) continue;     // more than a quarter of the tile recomputed
            for (int d = 1; d >= (inverse ? 1 : 0); --d) {
                size_t core, full;
                smem_of<T>(s, tf, inverse, &core, &full, d != 0);
                if (full <= budgets[bi]) {
                    if (dbuf) *dbuf = d != 0;
                    return tf;
                }
            }
        }
    return 0;
}

template <typename T, int TF>
cudaError_t launch_fwd(FwdArgs<T> &a, int64_t batch, bool dbuf, cudaStream_t st) {
    auto kern = smooth_fwd_kernel<T, TF>;
    static KernelSlot slot;                 // ONE slot per kernel function: the shared-memory attribute is per function
    size_t core, smem;
    smem_of<T>(a.s, TF, false, &core, &smem, dbuf);
    int per_sm = 0, dev = 0;
    cudaError_t e = prepare_kernel(slot, kern, SM_NT, smem, &per_sm, &dev);
    if (e != cudaSuccess) return e;
    a.ptiles = (a.P + TF - 1) / TF;
    const int64_t total = a.ptiles * batch, cap = (int64_t)per_sm * device_sms(dev);
    kern<<<(unsigned)(total < cap ? total : cap), SM_NT, smem, st>>>(a, total, (int)xin_bytes_of<T>(a.s, TF), dbuf ? 1 : 0);
    ++g_launch_count;
    return cudaGetLastError();
}
template <typename T, int TF>
cudaError_t launch_inv(InvArgs<T> &a, int64_t batch, cudaStream_t st) {
    auto kern = smooth_inv_kernel<T, TF>;
    static KernelSlot slot;
    size_t core, smem;
    smem_of<T>(a.s, TF, true, &core, &smem);
    int per_sm = 0, dev = 0;
    cudaError_t e = prepare_kernel(slot, kern, SM_NT, smem, &per_sm, &dev);
    if (e != cudaSuccess) return e;
    const int own = TF - a.halo;
    const int64_t qh0 = p2::fdiv64(a.k0 + a.mid, a.s.H), qh1 = p2::fdiv64(a.k1 - 1 + a.mid, a.s.H);
    a.ntiles = (qh1 - qh0 + 1 + own - 1) / own;
    a.qa0 = qh0 - a.halo;
    const int64_t total = a.ntiles * batch, cap = (int64_t)per_sm * device_sms(dev);
    kern<<<(unsigned)(total < cap ? total : cap), SM_NT, smem, st>>>(a, total, (int)ola_bytes_of<T>(a.s, TF));
    ++g_launch_count;
    return cudaGetLastError();
}
}  // namespace

bool smooth_supported(int64_t M, int64_t N, int64_t H, int precision, bool inverse) {
    sm::Shape s;
    if (N > M || H < 1 || !sm::make_shape(M, N, H, M, 0, &s)) return false;
    if (inverse && H > N) return false;
    return (precision == 0 ? pick_tf<float>(s, inverse) : pick_tf<double>(s, inverse)) > 0;
}
int64_t smooth_table_bytes(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse) {
    sm::Shape s;
    if (!sm::make_shape(M, N, H, M, ps, &s)) return 0;
    return precision == 0 ? sm::tab_layout<float>(s, inverse).bytes : sm::tab_layout<double>(s, inverse).bytes;
}
void smooth_table_fill(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse, const double *vec_n, void *out) {
    sm::Shape s;
    if (!sm::make_shape(M, N, H, M, ps, &s)) return;
    if (precision == 0) sm::tab_fill<float>(s, inverse, vec_n, (unsigned char *)out);
    else sm::tab_fill<double>(s, inverse, vec_n, (unsigned char *)out);
}

template <typename T>
cudaError_t smooth_forward(const T *x, void *out, const FwdParams &p, const void *tab, cudaStream_t st) {
    if (p.P <= 0 || p.batch <= 0) return cudaSuccess;
    sm::FwdArgs<T> a;
    if (!sm::make_shape(p.M, p.N, p.H, p.F, p.ps, &a.s)) return cudaErrorNotSupported;
    a.x = x; a.out = out; a.tab = (const unsigned char *)tab; a.lay = sm::tab_layout<T>(a.s, false);
    a.n = p.n; a.x_pitch = p.x_pitch; a.ldp = p.ldp; a.p0 = p.p0; a.P = p.P; a.k_offset = p.k_offset; a.ptiles = 0;
    a.mid = (int)p.mid; a.mode = p.mode; a.spectrogram = p.spectrogram; a.xpad = sm::xpad_of(p.H);
    a.fac2x = p.mode == 3 ? (T)p.fac2x : (T)1;
    constexpr int TFA = sizeof(T) == 4 ? 32 : 16, TFB = sizeof(T) == 4 ? 16 : 8, TFC = sizeof(T) == 4 ? 8 : 4;
    bool dbuf = true;
    const int tf = pick_tf<T>(a.s, false, &dbuf);
    if (tf == TFA) return launch_fwd<T, TFA>(a, p.batch, dbuf, st);
    if (tf == TFB) return launch_fwd<T, TFB>(a, p.batch, dbuf, st);
    if (tf == TFC) return launch_fwd<T, TFC>(a, p.batch, dbuf, st);
    return cudaErrorNotSupported;
}
template <typename T>
cudaError_t smooth_inverse(const void *S, T *x_out, const InvParams &p, const void *tab, cudaStream_t st) {
    if (p.batch <= 0 || p.k1 <= p.k0) return cudaSuccess;
    sm::InvArgs<T> a;
    if (!sm::make_shape(p.M, p.N, p.H, p.F, p.ps, &a.s)) return cudaErrorNotSupported;
    a.S = S; a.xo = x_out; a.tab = (const unsigned char *)tab; a.lay = sm::tab_layout<T>(a.s, true);
    a.P = p.P; a.ldp = p.ldp; a.out_pitch = p.out_pitch; a.p_min = p.p_min;
    a.k0 = p.k0; a.k1 = p.k1; a.q0 = p.q0; a.q1 = p.q1; a.qa0 = 0; a.ntiles = 0;
    a.mid = (int)p.mid; a.mode = p.mode; a.halo = halo_of(p.N, p.H); a.ostride = sm::ostride_of<T>(p.N, p.H);
    a.rfac2x = p.mode == 3 ? (T)(1.0 / p.fac2x) : (T)1;
    constexpr int TFA = sizeof(T) == 4 ? 32 : 16, TFB = sizeof(T) == 4 ? 16 : 8, TFC = sizeof(T) == 4 ? 8 : 4;
    const int tf = pick_tf<T>(a.s, true);
    if (tf == TFA) return launch_inv<T, TFA>(a, p.batch, st);
    if (tf == TFB) return launch_inv<T, TFB>(a, p.batch, st);
    if (tf == TFC) return launch_inv<T, TFC>(a, p.batch, st);
    return cudaErrorNotSupported;
}
template cudaError_t smooth_forward<float>(const float *, void *, const FwdParams &, const void *, cudaStream_t);
template cudaError_t smooth_forward<double>(const double *, void *, const FwdParams &, const void *, cudaStream_t);
template cudaError_t smooth_inverse<float>(const void *, float *, const InvParams &, const void *, cudaStream_t);
template cudaError_t smooth_inverse<double>(const void *, double *, const InvParams &, const void *, cudaStream_t);

}  // namespace veils
