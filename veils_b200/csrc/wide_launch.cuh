// wide_launch.cuh -- kernels and launch helpers of the wide family (phase code: wide_impl.cuh).
// Included by wide_fwd.cu and wide_inv.cu (two translation units so the many instantiations
// compile in parallel).
#pragma once
#include <cstdlib>
#include <cstring>

#include "wide_impl.cuh"

namespace veils {
namespace wd {

template <int NT> constexpr int wide_min_ctas() { return NT <= 256 ? 2 : 1; }

inline int env_int(const char *name, int dflt) {
    const char *e = getenv(name);
    return e ? atoi(e) : dflt;
}
// per-device attributes, read once
struct DevInfo { int sms = 0; };
inline const DevInfo &dev_info() {
    static thread_local int cached_dev = -1;
    static thread_local DevInfo info;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return info;
}

// ---- forward: persistent CTAs walk the (signal, tile) list; the next tile's samples are staged
//      (cp.async) while passes 2 and 3 of the current tile run
template <int LOGMC, int TF, int MODE, bool SPEC>
__global__ void __launch_bounds__(WGeo<LOGMC, TF>::NT, wide_min_ctas<WGeo<LOGMC, TF>::NT>())
wide_fwd_kernel(const __grid_constant__ WFwdArgs args, const int64_t total) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = WFwd<LOGMC, TF>;
    const int tid = threadIdx.x;
    WFwdCtx c;
    c.a = args;
    c.work = reinterpret_cast<float *>(smem);
    c.xin = c.work + K::WORK_FLOATS;
    int64_t w = blockIdx.x;
    if (w >= total) return;
    int64_t cb = w / args.ptiles, cpt = w - cb * args.ptiles;
    const int64_t gb = gridDim.x / args.ptiles, gr = gridDim.x - gb * args.ptiles;
    c.b = cb;
    c.pt = cpt;
    K::fill(c, tid);
    cp_async_commit();
    for (; w < total; w += gridDim.x) {
        cp_async_wait_all();
        __syncthreads();                       // the tile is staged; pass 3 of the previous tile has read `work`
        c.b = cb;
        c.pt = cpt;
        K::pass1(c, tid);
        __syncthreads();
        int64_t nb = cb + gb, npt = cpt + gr;
        if (npt >= args.ptiles) { npt -= args.ptiles; ++nb; }
        if (w + gridDim.x < total) {
            WFwdCtx cn = c;
            cn.b = nb;
            cn.pt = npt;
            K::fill(cn, tid);
            cp_async_commit();
        }
        K::pass2(c, tid);
        __syncthreads();
        K::template pass3_t<MODE, SPEC>(c, tid);
        cb = nb;
        cpt = npt;
    }
}

// ---- inverse: a CTA walks the tiles of one chunk of one signal in order (streaming overlap-add:
//      the accumulator rows TF .. TF+2 carry over to the next tile)
template <int LOGMC, int TF, int MODE>
__global__ void __launch_bounds__(WGeo<LOGMC, TF>::NT, wide_min_ctas<WGeo<LOGMC, TF>::NT>())
wide_inv_kernel(const __grid_constant__ WInvArgs args, const int64_t total) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = WInv<LOGMC, TF>;
    const int tid = threadIdx.x;
    WInvCtx c;
    c.a = args;
    c.work = reinterpret_cast<float *>(smem);
    c.acc = c.work + K::WORK_FLOATS;
    const int64_t lc = (int64_t)TF * args.chunk_tiles - args.halo;      // owned slices per chunk
    for (int64_t ch = blockIdx.x; ch < total; ch += gridDim.x) {
        c.b = ch / args.cps;
        const int64_t a0 = args.qh0 + (ch - c.b * args.cps) * lc;      // first owned slice
#pragma unroll 1
        for (int t = 0; t < args.chunk_tiles; ++t) {
            c.qa = a0 - args.halo + (int64_t)t * TF;
            if (c.qa > args.qh1) break;
            c.first = t == 0;
            c.more = (t + 1 < args.chunk_tiles && c.qa + TF <= args.qh1) ? 1 : 0;
            K::template pass1_t<MODE>(c, tid);
            __syncthreads();
            K::pass2(c, tid);
            __syncthreads();
            {
                typename K::Regs r;
                K::pass3_load(c, tid, r);
                if (c.first) K::zero_head(c, tid);
                K::template ola<3>(c, tid, r);
                __syncthreads();
                K::template ola<2>(c, tid, r);
                __syncthreads();
                K::template ola<1>(c, tid, r);
                __syncthreads();
                K::template ola<0>(c, tid, r);
            }
            __syncthreads();
            K::out(c, tid);
            typename K::Tail tail;
            K::take_tail(c, tid, tail);
            __syncthreads();
            K::keep_tail(c, tid, tail);
        }
    }
}

template <typename KernT>
inline cudaError_t wide_prepare(KernT kern, int nt, size_t smem, int *grid_cap) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int cap = env_int("VEILS_WIDE_CTAS_PER_SM", 0);
    if (cap > 0 && cap < per_sm) per_sm = cap;
    *grid_cap = per_sm * dev_info().sms;
    return cudaSuccess;
}

}  // namespace wd
}  // namespace veils
