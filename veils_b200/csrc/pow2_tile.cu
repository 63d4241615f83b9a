// pow2_tile.cu -- CUDA kernels and launchers of the power-of-two "slice-lane" family.
// The per-phase device code lives in pow2_tile_impl.cuh (shared with the host emulation).
#include <cstdlib>
#include <cstring>

#include "pow2_cfg.hpp"
#include "tma.cuh"
#include "launch_cache.cuh"

namespace veils {

namespace {

using namespace p2;

// two CTAs per SM whenever the work buffer (Mc * TF complex) allows it
template <typename T, int LOGMC, int TF>
constexpr int min_ctas() { return sizeof(T) * 2 * (1 << LOGMC) * TF <= 80 * 1024 ? 2 : 1; }

// cooperative copy of a plan table into shared memory
template <typename T>
__device__ __forceinline__ void copy_table(T *dst, const T *src, int n, int tid, int nt) {
    for (int i = tid; i < n; i += nt) dst[i] = __ldg(src + i);
}

// Persistent forward kernel: each CTA walks the (signal, tile) list with stride gridDim.x.
// Shared memory: work[Mc][TF] | tile buffer(s) | (STAB) twiddles, window.
// STAGE (one-sided layouts, S describable by a tensor map): the last pass overwrites the cells it
// has read with the finished bins and every group pair leaves as two dense [RL rows][TF slices]
// boxes through the TMA engine -- the short row runs of the large transforms (TF = 8 / 4 / 2
// slices) never go through the LSU store path.
template <typename T, int LOGMC, int TF, int NSUB, bool STAGE>
__global__ void __launch_bounds__(TF *NSUB, min_ctas<T, LOGMC, TF>())
pow2_fwd_kernel(const __grid_constant__ FwdArgs<T> args, const int64_t total, const int xin_elems,
                const __grid_constant__ CUtensorMap smap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = Fwd<T, LOGMC, TF, NSUB>;
    const int tid = threadIdx.x;
    FwdCtx<T> c;
    c.a = args;
    c.work = reinterpret_cast<cx<T> *>(smem);
    T *xin0 = reinterpret_cast<T *>(smem + sizeof(cx<T>) * (K::MCP + 1) * TF);      // + staged Nyquist row
    T *xin1 = K::DBUF ? xin0 + xin_elems : xin0;
    if (K::STAB) {
        T *stw = xin0 + (K::DBUF ? 2 : 1) * xin_elems;
        T *swin = stw + K::TWN;
        copy_table(stw, args.tw, K::TWN, tid, K::NT);
        copy_table(swin, args.win, args.N, tid, K::NT);
        c.tw = stw;
        c.win = swin;
    } else {
        c.tw = args.tw;
        c.win = args.win;
    }
    int64_t w = blockIdx.x;
    if (w < total) {
        c.b = w / args.ptiles;
        c.pt = w - c.b * args.ptiles;
        c.xin = xin0;
        K::fill(c, tid);
        cp_async_commit();
    }
    for (int it = 0; w < total; w += gridDim.x, ++it) {
        cp_async_wait_all();
        if (STAGE) tma::wait_read_all();   // the previous tile's boxes have left the work buffer
        __syncthreads();                   // tile w staged; previous tile's reads of `work` done
        T *cur = (K::DBUF && (it & 1)) ? xin1 : xin0;
        const int64_t wn = w + gridDim.x;
        if (K::DBUF && wn < total) {       // prefetch the next tile into the other buffer
            c.b = wn / args.ptiles;
            c.pt = wn - c.b * args.ptiles;
            c.xin = (it & 1) ? xin0 : xin1;
            K::fill(c, tid);
            cp_async_commit();
        }
        c.b = w / args.ptiles;
        c.pt = w - c.b * args.ptiles;
        c.xin = cur;
        K::pass1(c, tid);
        __syncthreads();
        if (!K::DBUF && wn < total) {      // single buffer: free once pass 1 is done
            FwdCtx<T> cn = c;
            cn.b = wn / args.ptiles;
            cn.pt = wn - cn.b * args.ptiles;
            K::fill(cn, tid);
            cp_async_commit();
        }
        if (K::NP == 3) {
            K::pass_mid(c, tid);
            __syncthreads();
        }
        if (K::UNPAIRED) {
            // one last-pass group per thread: Z goes back to the work buffer, the mirror values are read from there
            K::last_a(c, tid);
            __syncthreads();
            if (STAGE) {
                const int f = tid % TF;
                const int64_t cb = c.b, cpt = c.pt;
                const unsigned wbase = tma::smem_u32(c.work);
                constexpr int MUL = sizeof(T) == 8 ? 2 : 1;  // complex f64 travels as two 8-byte elements
                auto hook = [&](int g) {
                    tma::fence_async_smem();
                    __syncwarp();
                    if (g >= 0 && f == 0) {
                        const int mul = args.spectrogram ? 1 : MUL;
                        const int p0 = (int)(cpt * TF) * mul;
                        constexpr unsigned RB = TF * (unsigned)sizeof(cx<T>);
                        tma::store_4d(&smap, wbase + K::GEO::goff(g) * RB, p0, 0, g, (int)cb);
                        if (g == 0) tma::store_4d(&smap, wbase + K::MCP * RB, p0, K::RL, 0, (int)cb);
                        tma::commit_group();
                    }
                };
                cx<T> X[K::GPTU][K::RL], xn = mk<T>(0, 0);
                K::last_x(c, tid, X, xn);
                __syncthreads();           // every mirror value has been read: the rows may take the finished bins
                if (args.spectrogram) K::template last_put<true, true, true>(c, tid, X, xn, hook);
                else K::template last_put<true, false, true>(c, tid, X, xn, hook);
            } else {
                K::last_b(c, tid);
            }
        } else if (STAGE) {
            const int f = tid % TF;
            const int64_t cb = c.b, cpt = c.pt;
            const unsigned wbase = tma::smem_u32(c.work);
            constexpr int MUL = sizeof(T) == 8 ? 2 : 1;      // complex f64 travels as two 8-byte elements
            auto hook = [&](int s2) {
                tma::fence_async_smem();
                __syncwarp();
                if (s2 >= 0 && f == 0) {
                    const int g = s2, gp = (s2 == 0) ? K::G / 2 : K::G - s2;
                    const int mul = args.spectrogram ? 1 : MUL;
                    const int p0 = (int)(cpt * TF) * mul;
                    constexpr unsigned RB = TF * (unsigned)sizeof(cx<T>);
                    const bool sp = args.spectrogram != 0;
                    tma::store_4d(&smap, wbase + K::GEO::goff(g) * RB + (sp ? K::sskew(g) : 0), p0, 0, g, (int)cb);
                    tma::store_4d(&smap, wbase + K::GEO::goff(gp) * RB + (sp ? K::sskew(gp) : 0), p0, 0, gp, (int)cb);
                    if (s2 == 0) tma::store_4d(&smap, wbase + K::MCP * RB, p0, K::RL, 0, (int)cb);
                    tma::commit_group();
                }
            };
            if (args.spectrogram) K::template pass_last_t<true, true, true>(c, tid, hook);
            else K::template pass_last_t<true, false, true>(c, tid, hook);
        } else {
            K::pass_last(c, tid);
        }
    }
    if (STAGE) tma::wait_all();
}

// Persistent inverse kernel.  Shared memory: work[Mc][TF] (aliased by ola) | (STAB) twiddles, dual.
template <typename T, int LOGMC, int TF, int NSUB>
__global__ void __launch_bounds__(TF *NSUB, min_ctas<T, LOGMC, TF>())
pow2_inv_kernel(const __grid_constant__ InvArgs<T> args, const int64_t total, const int work_bytes, const int carry_elems) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = Inv<T, LOGMC, TF, NSUB>;
    const int tid = threadIdx.x;
    InvCtx<T> c;
    c.a = args;
    c.work = reinterpret_cast<cx<T> *>(smem);
    c.ola = reinterpret_cast<T *>(smem);
    T *carry0 = reinterpret_cast<T *>(smem + work_bytes);
    T *carry1 = carry0 + carry_elems;
    if (K::STAB) {
        T *stw = carry1 + carry_elems;
        T *sdual = stw + K::TWN;
        copy_table(stw, args.tw, K::TWN, tid, K::NT);
        copy_table(sdual, args.dual, args.N, tid, K::NT);
        c.tw = stw;
        c.dual = sdual;
    } else {
        c.tw = args.tw;
        c.dual = args.dual;
    }
    __syncthreads();
    // streaming overlap-add: a CTA walks the tiles of one chunk of one signal in order; only the first tile of a chunk
    // re-computes the `halo` slices before its first owned one, every other tile owns all TF slices
    const int64_t lc = (int64_t)TF * args.chunk_tiles - args.halo;       // owned slices per chunk
    for (int64_t ch = blockIdx.x; ch < total; ch += gridDim.x) {
        c.b = ch / args.cps;
        const int64_t a0 = args.qh0 + (ch - c.b * args.cps) * lc;         // first owned slice
        c.kown = a0 * (int64_t)args.H - args.mid;
        int flip = 0;
#pragma unroll 1
        for (int t = 0; t < args.chunk_tiles; ++t) {
            c.qa = a0 - args.halo + (int64_t)t * TF;
            if (c.qa > args.qh1) break;
            c.first = t == 0;
            c.last = t + 1 >= args.chunk_tiles || c.qa + TF > args.qh1;
            c.carry_in = flip ? carry1 : carry0;
            c.carry_out = flip ? carry0 : carry1;
            flip ^= 1;
            K::pass1(c, tid);
            __syncthreads();
            if (K::NP == 3) {
                K::pass_mid(c, tid);
                __syncthreads();
            }
            cx<T> r[K::GPT][K::RL];
            K::last_load(c, tid, r);
            __syncthreads();               // all reads of `work` done before it becomes `ola`
            K::last_store(c, tid, r);
            __syncthreads();
            K::gather_s(c, tid);
            __syncthreads();               // `ola` consumed before the next tile overwrites `work`
        }
    }
}

struct FwdLaunch {
    const void *x; void *out; const FwdParams *prm; const void *win; const void *tw; cudaStream_t st;
    template <typename T, int LOGMC, int TF, int NSUB0>
    cudaError_t run() {
        // VEILS_F64_FWD_UNPAIRED: the last pass works one group per thread (Fwd::last_a / last_x), so twice the threads find work
        constexpr int GL = (1 << LOGMC) / Fwd<T, LOGMC, TF, NSUB0>::RL;
        constexpr int NSUB = (VEILS_F64_FWD_UNPAIRED && sizeof(T) == 8 && 2 * TF * NSUB0 <= 512 && 2 * NSUB0 <= GL) ? 2 * NSUB0 : NSUB0;
        using K = Fwd<T, LOGMC, TF, NSUB>;
        const FwdParams &p = *prm;
        FwdArgs<T> a = make_fwd_args<T>((const T *)x, out, p, (const T *)win, (const T *)tw, TF);
        const int xin_elems = (int)fwd_xin_elems(p.N, p.H, TF, a.lay);
        const size_t smem = fwd_smem_bytes_t<T, LOGMC, TF>(p.N, xin_elems);
        CUtensorMap smap;
        memset(&smap, 0, sizeof(smap));
        // measured (profiles/r1c_ablation.txt): +3 % for f64; the f32 tile kernels (mfft >= 2048 here) are
        // 5 % faster with their plain stores, so only f64 takes the staged path
        // (padded narrow tiles: their rows are not 128-byte aligned, which a tensor-map source must be)
        // round 2: with two CTAs per SM (mfft <= 512: rows of 16 / 32 slices = 256 / 512-byte runs) the plain stores win
        // (c2 geometry in f64: 55.0 -> 61.0 %); the narrow tiles of mfft >= 1024 keep the staged path (c1, c5: 4-7 % better)
        bool stage = sizeof(T) == 8 && LOGMC >= 9 && p.mode >= 2 && p.batch < (1LL << 31) && !K::GEO::PADR && !VEILS_ENV_ONCE("VEILS_PK_NO_TMA", 0);
        if (stage) {
            // complex f64 elements are 16 bytes: described as pairs of 8-byte elements
            const int mul = (!p.spectrogram && sizeof(T) == 8) ? 2 : 1;
            const int esz = p.spectrogram ? (int)sizeof(T) : (sizeof(T) == 8 ? 8 : 8);
            const MapKey key{out, {esz, p.P * mul, p.ldp * mul, p.F, p.batch, K::G * 1000 + K::RL, TF * mul, 1}};
            stage = cached_map(&smap, key, [&](CUtensorMap *m) {
                return tma::make_s_map(m, out, esz, p.P * mul, p.ldp * mul, p.F, p.batch, K::G, K::RL, TF * mul, K::RL);
            });
        }
        auto kern = stage ? pow2_fwd_kernel<T, LOGMC, TF, NSUB, true> : pow2_fwd_kernel<T, LOGMC, TF, NSUB, false>;
        static KernelSlot slots[2];
        int per_sm = 0, dev = 0;
        cudaError_t e = prepare_kernel(slots[stage ? 1 : 0], kern, TF * NSUB, smem, &per_sm, &dev);
        if (e != cudaSuccess) return e;
        const int64_t total = a.ptiles * p.batch;
        const int64_t cap = (int64_t)per_sm * device_sms(dev);
        const unsigned grid = (unsigned)(total < cap ? total : cap);
        (void)sizeof(K);
        kern<<<grid, TF * NSUB, smem, st>>>(a, total, xin_elems, smap);
        ++g_launch_count;
        return cudaGetLastError();
    }
};

struct InvLaunch {
    const void *S; void *xo; const InvParams *prm; const void *dual; const void *tw; cudaStream_t st;
    template <typename T, int LOGMC, int TF0, int NSUB00>
    cudaError_t run() {
#ifndef VEILS_F64_L8_INV_TF
#define VEILS_F64_L8_INV_TF 32
#endif
#ifndef VEILS_F64_INV_MAXT
#define VEILS_F64_INV_MAXT 512
#endif
        // the inverse may use a wider tile than the forward (fewer halo slices recomputed per owned slice)
        constexpr bool WIDER = sizeof(T) == 8 && LOGMC == 8 && VEILS_F64_L8_INV_TF != TF0;
        constexpr int TF = WIDER ? VEILS_F64_L8_INV_TF : TF0;
        constexpr int L1H = (1 << LOGMC) / Fac<LOGMC>::R1 / 2;
        constexpr int NSUB0 = WIDER ? (256 / TF < L1H ? 256 / TF : L1H) : NSUB00;
        // f64: pass 1 works one group per thread (Inv::pass1_u), so twice the threads find work
        constexpr int NSUB = (sizeof(T) == 8 && 2 * TF * NSUB0 <= VEILS_F64_INV_MAXT) ? 2 * NSUB0 : NSUB0;
        const InvParams &p = *prm;
        InvArgs<T> a = make_inv_args<T>(S, (T *)xo, p, (const T *)dual, (const T *)tw, TF);
        const int work_bytes = (int)inv_work_bytes_t<T, LOGMC, TF>(p.N);
        const size_t smem = inv_smem_bytes_t<T, LOGMC, TF>(p.N, p.H);
        auto kern = pow2_inv_kernel<T, LOGMC, TF, NSUB>;
        static KernelSlot slot;
        int per_sm = 0, dev = 0;
        cudaError_t e = prepare_kernel(slot, kern, TF * NSUB, smem, &per_sm, &dev);
        if (e != cudaSuccess) return e;
        const int64_t cap = (int64_t)per_sm * device_sms(dev);
        tile_inv_chunks<T>(a, p, TF, p.batch, cap);
        const int64_t total = a.cps * p.batch;
        const unsigned grid = (unsigned)(total < cap ? total : cap);
        kern<<<grid, TF * NSUB, smem, st>>>(a, total, work_bytes, (int)inv_carry_elems(p.N, p.H));
        ++g_launch_count;
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
