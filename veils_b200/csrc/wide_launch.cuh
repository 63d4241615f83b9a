// wide_launch.cuh -- kernels and launch helpers of the wide family (phase code: wide_impl.cuh).
// Included by wide_fwd.cu and wide_inv.cu (two translation units so the many instantiations
// compile in parallel).
#pragma once
#include <cstdlib>
#include <cstring>

#include "wide_impl.cuh"
#include "tma.cuh"
#include "launch_cache.cuh"

namespace veils {
namespace wd {

template <int NT> constexpr int wide_min_ctas() { return NT <= 256 ? 2 : 1; }

#define env_int(name, dflt) VEILS_ENV_ONCE(name, dflt)          // environment switches: read once per process

// plan table -> shared memory, once per persistent CTA (offsets: wide_table_offsets)
#define TAB_ENTRIES(M) ((M) / 2 * 3 + (M) / 16 + 2)
#define TAB_XTAB(M) 0
#define TAB_TW1(M) ((M) / 2)
#define TAB_TW2(M) (M)
#define TAB_TWP(M) ((M) + (M) / 16)
__device__ __forceinline__ void copy_table(fl2 *dst, const fl2 *src, int n, int tid, int nt) {
    for (int i = tid; i < n; i += nt) {
        const float2 v = __ldg(reinterpret_cast<const float2 *>(src + i));
        fl2 r; r.x = v.x; r.y = v.y;
        dst[i] = r;
    }
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
    c.xin = c.work + (K::WORK_FLOATS + 64);
    {
        fl2 *tab = reinterpret_cast<fl2 *>(c.xin + K::XIN_FLOATS);
        copy_table(tab, args.tab, TAB_ENTRIES(K::M), tid, K::NT);
        c.wtab = tab + TAB_XTAB(K::M); c.tw1 = tab + TAB_TW1(K::M); c.tw2 = tab + TAB_TW2(K::M); c.twp = tab + TAB_TWP(K::M);
    }
    // staging of a tile: interior tiles whose first sample is 16-byte aligned come in as ROWS bulk copies
    // (one thread each, completion on an mbarrier); border / unaligned tiles through cp.async with zero fill
    const unsigned fill_bar = tma::smem_u32(reinterpret_cast<fl2 *>(c.xin + K::XIN_FLOATS) + TAB_ENTRIES(K::M));
    if (tid == 0) {
        tma::mbar_init(fill_bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    const bool x_al = ((((size_t)args.x) & 15) == 0) && !args.no_bulk;
    auto stage = [&](int64_t b, int64_t pt) -> bool {
        const int64_t s0 = (args.p0 + pt * TF) * (int64_t)K::H + args.k_offset - K::MID;
        const bool bulk = x_al && s0 >= 0 && s0 + (int64_t)K::ROWS * K::H <= args.n && (((b * args.x_pitch + s0) & 3) == 0);
        if (bulk) {
            if (tid < K::ROWS) {
                tma::fence_async_smem();
                if (tid == 0) tma::mbar_expect_tx(fill_bar, 4u * K::ROWS * K::H);
                __syncwarp((1u << K::ROWS) - 1u);
                tma::load_bulk(tma::smem_u32(c.xin + tid * K::HS), args.x + b * args.x_pitch + s0 + (int64_t)tid * K::H,
                               4u * K::H, fill_bar);
            }
        } else {
            WFwdCtx cn = c;
            cn.b = b;
            cn.pt = pt;
            K::fill(cn, tid);
            cp_async_commit();
        }
        return bulk;
    };
    int64_t w = blockIdx.x;
    if (w >= total) return;
    int64_t cb = w / args.ptiles, cpt = w - cb * args.ptiles;
    const int64_t gb = gridDim.x / args.ptiles, gr = gridDim.x - gb * args.ptiles;
    bool bulk_cur = stage(cb, cpt);
    unsigned phase = 0;
    for (; w < total; w += gridDim.x) {
        if (bulk_cur) {
            tma::mbar_wait(fill_bar, phase);
            phase ^= 1;
        } else {
            cp_async_wait_all();
        }
        __syncthreads();                       // the tile is staged; pass 3 of the previous tile has read `work`
        c.b = cb;
        c.pt = cpt;
        K::pass1(c, tid);
        __syncthreads();
        int64_t nb = cb + gb, npt = cpt + gr;
        if (npt >= args.ptiles) { npt -= args.ptiles; ++nb; }
        if (w + gridDim.x < total) bulk_cur = stage(nb, npt);
        K::pass2(c, tid);
        __syncthreads();
        K::template pass3_t<MODE, SPEC>(c, tid);
        cb = nb;
        cpt = npt;
    }
}

// ---- inverse: a CTA walks the tiles of one chunk of one signal in order (streaming overlap-add:
//      the accumulator rows TF .. TF+2 carry over to the next tile).
//      LAND (one-sided layouts, S describable by a tensor map): the S tile of the NEXT work item is
//      requested from the TMA engine as soon as the last pass has read the work buffer and lands over
//      it while the overlap-add and the output of the current tile run; pass 1 then reads shared
//      memory.  Otherwise pass 1 loads S straight from global memory.
struct WTile {
    int64_t ch, b, qa, kend;
    int t, halo;
    bool valid;
};
template <int TF, int H, int MID>
__device__ __forceinline__ WTile wide_tile_of_chunk(const WInvArgs &a, int64_t ch, int64_t total) {
    WTile w;
    w.ch = ch;
    w.t = 0;
    w.valid = ch < total;
    const WChunk c = wide_chunk(a, w.valid ? ch : 0, TF, H, MID);
    w.b = c.b; w.qa = c.qa0; w.kend = c.kend; w.halo = c.halo;
    return w;
}
template <int TF, int H, int MID>
__device__ __forceinline__ WTile wide_next_tile(const WInvArgs &a, const WTile &c, int64_t total) {
    if (c.t + 1 < a.chunk_tiles && c.qa + TF <= a.qh1) {
        WTile w = c;
        w.t = c.t + 1;
        w.qa = c.qa + TF;
        return w;
    }
    return wide_tile_of_chunk<TF, H, MID>(a, c.ch + gridDim.x, total);
}

template <int LOGMC, int TF, int MODE, bool LAND>
__global__ void __launch_bounds__(WGeo<LOGMC, TF>::NT, wide_min_ctas<WGeo<LOGMC, TF>::NT>())
wide_inv_kernel(const __grid_constant__ WInvArgs args, const int64_t total, const __grid_constant__ CUtensorMap smap,
                const __grid_constant__ CUtensorMap smap_nyq) {
    extern __shared__ __align__(1024) unsigned char smem[];
    using K = WInv<LOGMC, TF>;
    const int tid = threadIdx.x;
    WInvCtx c;
    c.a = args;
    c.work = reinterpret_cast<float *>(smem);
    c.acc = c.work + (K::WORK_FLOATS + 64);
    fl2 *tab = reinterpret_cast<fl2 *>(c.acc + K::XIN_FLOATS);
    copy_table(tab, args.tab, TAB_ENTRIES(K::M), tid, K::NT);
    c.dtab = tab + TAB_XTAB(K::M); c.tw1 = tab + TAB_TW1(K::M); c.tw2 = tab + TAB_TW2(K::M); c.twp = tab + TAB_TWP(K::M);
    const unsigned land_bar = tma::smem_u32(tab + TAB_ENTRIES(K::M));
    if (LAND && tid == 0) {
        tma::mbar_init(land_bar, 1);
        tma::fence_mbar_init();
    }
    __syncthreads();
    auto land = [&](const WTile &w) {
        if (tid != 0) return;
        tma::fence_async_smem();               // earlier generic-proxy accesses to the buffer are ordered first
        const unsigned wbase = tma::smem_u32(c.work);
        const int col0 = (int)(w.qa - args.p_min);
        tma::mbar_expect_tx(land_bar, K::LAND_BYTES);
#pragma unroll
        for (int h = 0; h < K::NH; ++h) {
#pragma unroll
            for (int r = 0; r < K::MC; r += K::LAND_ROWS)
                tma::load_3d(&smap, wbase + 4u * (h * K::LAND_BLOCK + 16 * r), land_bar, col0 + 8 * h, r, (int)w.b);
            tma::load_3d(&smap_nyq, wbase + 4u * (K::NH * K::LAND_BLOCK + K::NYQ_STRIDE * h), land_bar, col0 + 8 * h, K::MC, (int)w.b);
        }
    };
    // the tile after the next one is pulled into L2 while the current one is transformed
    auto prefetch = [&](const WTile &w) {
        if (tid != 32) return;
        const int col0 = (int)(w.qa - args.p_min);
#pragma unroll
        for (int h = 0; h < K::NH; ++h) {
#pragma unroll
            for (int r = 0; r < K::MC; r += K::LAND_ROWS) tma::prefetch_l2_3d(&smap, col0 + 8 * h, r, (int)w.b);
            tma::prefetch_l2_3d(&smap_nyq, col0 + 8 * h, K::MC, (int)w.b);
        }
    };
    WTile cur = wide_tile_of_chunk<TF, K::H, K::MID>(args, blockIdx.x, total);
    unsigned phase = 0;
    if (LAND && cur.valid) land(cur);
    while (cur.valid) {
        const WTile nxt = wide_next_tile<TF, K::H, K::MID>(args, cur, total);
        c.b = cur.b;
        c.qa = cur.qa;
        c.kend = cur.kend;
        c.halo = cur.halo;
        c.first = cur.t == 0;
        c.more = (nxt.valid && nxt.ch == cur.ch) ? 1 : 0;
        if (LAND) {
            if (args.prefetch && nxt.valid) prefetch(nxt);
            tma::mbar_wait(land_bar, phase);
            phase ^= 1;
#pragma unroll 1
            for (int it = 0; it < K::P1_ITERS; ++it) {
                typename K::P1Regs r;
                K::template land_load<MODE>(c, tid, it, r);
                __syncthreads();               // every thread holds its bins before anyone overwrites the tile
                K::land_rest(c, tid, it, r);
            }
        } else {
            K::template pass1_t<MODE>(c, tid);
        }
        __syncthreads();
        K::pass2(c, tid);
        __syncthreads();
        {
            typename K::Regs r;
            K::pass3_load(c, tid, r);
            if (LAND) {
                __syncthreads();               // the work buffer is free: the next tile lands during what follows
                if (nxt.valid) land(nxt);
            }
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
        cur = nxt;
    }
}

template <typename KernT>
inline cudaError_t wide_prepare(KernelSlot &slot, KernT kern, int nt, size_t smem, int *grid_cap) {
    int per_sm = 0, dev = 0;
    cudaError_t e = prepare_kernel(slot, kern, nt, smem, &per_sm, &dev);
    if (e != cudaSuccess) return e;
    const int cap = env_int("VEILS_WIDE_CTAS_PER_SM", 0);
    if (cap > 0 && cap < per_sm) per_sm = cap;
    *grid_cap = per_sm * device_sms(dev);
    return cudaSuccess;
}

}  // namespace wd
}  // namespace veils
