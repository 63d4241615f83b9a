// wide_inv.cu -- inverse launcher of the wide family (mfft 1024 / 2048 / 4096, default geometry).
#include "wide_launch.cuh"

namespace veils {

namespace {
using namespace wd;

struct InvLaunch {
    const void *S; float *xo; const InvParams *prm; const float *tab; cudaStream_t st;
    template <int LOGMC, int TF, int MODE>
    cudaError_t go(WInvArgs &a) {
        using K = WInv<LOGMC, TF>;
        // one-sided layouts: the S tiles land through the TMA engine when S can be described by a tensor map
        // (16-byte aligned base and row pitch); P1_ITERS > 1 would overwrite the landed tile between iterations
        CUtensorMap smap, smap_nyq;
        memset(&smap, 0, sizeof(smap));
        memset(&smap_nyq, 0, sizeof(smap_nyq));
        bool land = MODE >= 2 && K::P1_ITERS == 1 && a.vec16 && !env_int("VEILS_WIDE_NO_TMA", 0);
        if (land) {
            const MapKey k1{S, {a.P, a.ldp, a.F, prm->batch, K::LAND_ROWS, 301, 0, 0}};
            land = cached_map(&smap, k1, [&](CUtensorMap *m) { return tma::make_land_map(m, S, a.P, a.ldp, a.F, prm->batch, K::LAND_ROWS); });
        }
        if (land) {
            const MapKey k2{S, {a.P, a.ldp, a.F, prm->batch, 1, 302, 0, 0}};
            land = cached_map(&smap_nyq, k2, [&](CUtensorMap *m) { return tma::make_land_map(m, S, a.P, a.ldp, a.F, prm->batch, 1); });
        }
        using KernT = void (*)(const WInvArgs, const int64_t, const CUtensorMap, const CUtensorMap);
        KernT kern = wide_inv_kernel<LOGMC, TF, MODE, false>;
        if (MODE >= 2 && land) kern = wide_inv_kernel<LOGMC, TF, (MODE >= 2 ? MODE : 2), true>;
        const size_t smem = wide_inv_smem_bytes<LOGMC, TF>();
        int cap = 0;
        static KernelSlot slots[2];
        cudaError_t e = wide_prepare(slots[(MODE >= 2 && land) ? 1 : 0], kern, K::NT, smem, &cap);
        if (e != cudaSuccess) return e;
        // streaming overlap-add: chunks of `ct` tiles; the first tile of a chunk re-computes `halo` slices
        const int64_t Q = a.qh1 - a.qh0 + 1;
#ifndef VEILS_CHUNK_HALO_FACTOR
#define VEILS_CHUNK_HALO_FACTOR 50
#endif
        int ct = (VEILS_CHUNK_HALO_FACTOR * 4 + TF - 1) / TF;      // halo re-computation <= ~2 % of a chunk
        if (ct > 64) ct = 64;
        auto cps = [&](int c) { const int64_t lc = (int64_t)TF * c - 4; return (Q + lc - 1) / lc; };
        while (ct > 2 && prm->batch * cps(ct) < 4 * (int64_t)cap) --ct;   // enough chunks to balance the CTAs
        a.chunk_tiles = ct;
        a.cps = cps(ct);
        const int64_t total = a.cps * prm->batch;
        kern<<<(unsigned)(total < cap ? total : cap), K::NT, smem, st>>>(a, total, smap, smap_nyq);
        ++g_launch_count;
        return cudaGetLastError();
    }
    template <int LOGMC, int TF>
    cudaError_t run() {
        using K = WInv<LOGMC, TF>;
        const InvParams &p = *prm;
        const WTabOff o = wide_table_offsets(p.M);
        const fl2 *t = reinterpret_cast<const fl2 *>(tab);
        WInvArgs a;
        a.S = S; a.xo = xo;
        a.tab = t;
        (void)o;
        a.P = p.P; a.ldp = p.ldp; a.out_pitch = p.out_pitch; a.p_min = p.p_min;
        a.k0 = p.k0; a.k1 = p.k1; a.q0 = p.q0; a.q1 = p.q1; a.F = p.F;
        a.mode = p.mode;
        a.rfac2x = p.mode == 3 ? (float)(1.0 / p.fac2x) : 1.f;
        a.vec16 = (p.ldp % 2 == 0 && ((uintptr_t)S % 16) == 0) ? 1 : 0;
        a.vec_out = ((uintptr_t)xo % 16) == 0 ? 1 : 0;
        a.prefetch = env_int("VEILS_WIDE_PREFETCH", 1);
        a.qh0 = fdiv64(p.k0 + K::MID, K::H);
        a.qh1 = fdiv64(p.k1 - 1 + K::MID, K::H);
        a.chunk_tiles = 0; a.cps = 0;
        switch (p.mode) {
            case 0: return go<LOGMC, TF, 0>(a);
            case 1: return go<LOGMC, TF, 1>(a);
            case 2: return go<LOGMC, TF, 2>(a);
            default: return go<LOGMC, TF, 3>(a);
        }
    }
};
}  // namespace

cudaError_t wide_inverse(const void *S, float *x_out, const InvParams &prm, const float *tab, int tf, cudaStream_t st) {
    if (prm.batch <= 0 || prm.k1 <= prm.k0) return cudaSuccess;
    wd::WCfg c;
    if (!wd::wide_cfg_for(prm.M, tf, &c)) return cudaErrorNotSupported;
    InvLaunch l{S, x_out, &prm, tab, st};
    return wd::wide_dispatch(c, l);
}

}  // namespace veils
