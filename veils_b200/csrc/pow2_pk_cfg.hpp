// pow2_pk_cfg.hpp -- host-side configuration of the packed f32 family (pow2_pk_impl.cuh),
// shared by the CUDA launchers (pow2_pk.cu) and the host emulation (tests/emu/pow2_emu.cu).
#pragma once
#include "pow2_cfg.hpp"
#include "pow2_pk_impl.cuh"

namespace veils {
namespace p2 {

#ifndef VEILS_PK_TF8
#define VEILS_PK_TF8 32                               // slices per tile for mfft = 512 (tuning knob)
#endif
inline bool pk_cfg_for(int64_t M, Cfg *c) {
    const int lm = ilog2_exact(M);
    if (lm < 5 || lm > 13) return false;
    c->logmc = lm - 1;
    static const int tf[13] = {0, 0, 0, 0, 32, 32, 32, 32, VEILS_PK_TF8, 32, 16, 8, 4};
    c->tf = tf[c->logmc];
    c->nsub = 0;                                      // fixed by the dispatch table below
    return true;
}

inline int64_t pk_twiddle_count(int64_t M) {          // floats
    Cfg c;
    if (!pk_cfg_for(M, &c)) return 0;
    int np, r1, r2, r3;
    fac_of(c.logmc, &np, &r1, &r2, &r3);
    const int64_t mc = M / 2, l1 = mc / r1;
    return 4 * (3 * mc + 2 * l1 + 1 + mc / 2);
}
// entries (wr, wi, -wr, -wi), w = product of unit roots exp(-2 pi i m / n):
//   [tw1: MC  at j*R1+c ][tw2: L1 at j2*R2+c2][twp: W_M^k, k <= Mc][tw1x: MC][tw2x: L1]
// forward table:  tw1[j][c] = W_Mc^(j c), times W_L1^(-j) for the reversed groups c > R1/2 of 2-pass
//                 plans (PkGeo::rev_group); tw1x / tw2x unused (zero).
// inverse table:  tw1[j][c] = W_Mc^(j c);  tw1x[j][k] = W_R1^k W_Mc^(j ((R1-k) mod R1)) for the
//                 half-1 groups that run the IDFT on a reversed array (pow2_pk_impl.cuh p1_rest).
inline void pk_twiddle_fill(int64_t M, bool inverse, double *t) {
    Cfg c;
    if (!pk_cfg_for(M, &c)) return;
    int np, r1, r2, r3;
    fac_of(c.logmc, &np, &r1, &r2, &r3);
    const int64_t mc = M / 2, l1 = mc / r1;
    double *p = t;
    auto put2 = [&](int64_t m1, int64_t n1, int64_t m2, int64_t n2) {   // W_n1^m1 * W_n2^m2
        double a, b, cr, ci;
        tw_exact(m1, n1, &a, &b);
        tw_exact(m2, n2, &cr, &ci);
        const double re = a * cr - b * ci, im = a * ci + b * cr;
        p[0] = re; p[1] = im; p[2] = -re; p[3] = -im;
        p += 4;
    };
    for (int64_t j = 0; j < l1; ++j)
        for (int64_t cc = 0; cc < r1; ++cc) {
            const bool rev = !inverse && np == 2 && cc > r1 / 2;
            put2(j * cc, mc, rev ? -j : 0, l1);
        }
    if (np == 3) {
        const int64_t l2 = l1 / r2;
        for (int64_t j2 = 0; j2 < l2; ++j2)
            for (int64_t cc = 0; cc < r2; ++cc) put2(j2 * cc, l1, 0, 1);
    } else {
        for (int64_t i = 0; i < l1; ++i) put2(0, 1, 0, 1);
    }
    for (int64_t k = 0; k <= mc; ++k) put2(k, M, 0, 1);
    for (int64_t j = 0; j < l1; ++j)
        for (int64_t k = 0; k < r1; ++k) {
            if (inverse) put2(k, r1, j * ((r1 - k) % r1), mc);
            else put2(0, 1, 0, 1);
        }
    for (int64_t i = 0; i < l1; ++i) put2(0, 1, 0, 1);
    // tw1p (paired inverse pass 1): (Re W^(j0 c), Re W^(j1 c), Im W^(j0 c), Im W^(j1 c)), W = W_Mc
    for (int64_t w = 0; w < l1 / 2; ++w)
        for (int64_t cc = 0; cc < r1; ++cc) {
            const int64_t j0 = w, j1 = w == 0 ? l1 / 2 : l1 - w;
            double a, b, cr, ci;
            tw_exact(j0 * cc, mc, &a, &b);
            tw_exact(j1 * cc, mc, &cr, &ci);
            p[0] = a; p[1] = cr; p[2] = b; p[3] = ci;
            p += 4;
        }
}

constexpr int PK_NYQ_ROW = 64 + 4; // floats kept behind the work buffer: the staged Nyquist row (TMA stores)
                                   // + 16 bytes for the drain mbarrier
template <int LOGMC, int TF>
inline size_t pk_fwd_smem_bytes(int64_t N, int64_t xin_elems) {
    using G = PkGeo<LOGMC, TF, 2>;
    (void)N;
    size_t b = 4 * ((size_t)G::MC * 2 * TF + PK_NYQ_ROW + (size_t)((xin_elems + 3) & ~3LL) * (G::DBUF ? 2 : 1));
    if (G::STAB) b += 4 * (size_t)G::TWN_FWD + sizeof(PTab) * (size_t)G::MC;
    return b;
}
// per-plan tile table: entry m = sample pair (2m, 2m+1) of the rolled, zero-padded, windowed slice
// (src/lib.rs:336-345, 715-729): tile offset of x[j0], window values * 0.5 (real-FFT split).
// Padded pairs (j0 >= m_num) point at the zero pair stored behind the staged span.
inline void pk_ptab_fill(int64_t M, int64_t N, int64_t H, int64_t ps, const double *win, int tf, PTab *t) {
    const TileLayout lay = tile_layout(H, true);
    const int64_t span = (int64_t)(tf - 1) * H + N;
    for (int64_t m = 0; m < M / 2; ++m) {
        const int64_t j0 = (2 * m + ps) & (M - 1);
        PTab e;
        e.pad = 0;
        if (j0 < N) {
            e.aoff = lay.addr((int)j0);
            e.w0 = (float)(win[j0] * 0.5);
            e.w1 = (float)(win[j0 + 1] * 0.5);
        } else {
            e.aoff = 0;
            e.pad = 1;                                 // zero-padded pair (mfft > m_num): no tile read
            e.w0 = e.w1 = 0.f;
        }
        t[m] = e;
    }
}
// inverse plan table: sample pair m of a slice -> positions in its overlap row and dual / M.
// y index i = 2m, 2m+1 is rolled back to j = (i + ps) mod M (src/lib.rs:499-500); j >= m_num is
// truncated (:503): such samples are written to the dump slot at offset m_num.
inline void pk_itab_fill(int64_t M, int64_t N, int64_t ps, const double *dual_scaled, ITab *t) {
    for (int64_t m = 0; m < M / 2; ++m) {
        const int64_t j0 = (2 * m + ps) & (M - 1), j1 = (j0 + 1) & (M - 1);
        ITab e;
        e.o0 = j0 < N ? (int)j0 : (int)N;
        e.o1 = j1 < N ? (int)j1 : (int)N;
        e.d0 = j0 < N ? (float)dual_scaled[j0] : 0.f;
        e.d1 = j1 < N ? (float)dual_scaled[j1] : 0.f;
        t[m] = e;
    }
}
// overlap-row stride: > m_num (dump slot); pairs: = 2 (mod 4) so the 8-byte stores of the 16 lanes
// of a half-warp hit 32 distinct banks; otherwise odd (4-byte stores).
inline bool pk_pair(int64_t N, int64_t ps) { return ps % 2 == 0 && N % 2 == 0; }
inline int pk_ostride(int64_t N, int64_t ps) {
    if (pk_pair(N, ps)) { int64_t s = N + 2; while (s % 4 != 2) ++s; return (int)s; }
    return (int)((N + 1) | 1);
}
inline int pk_zero_off(int64_t N, int64_t H, int tf) {
    const TileLayout lay = tile_layout(H, true);
    const int64_t span = (int64_t)(tf - 1) * H + N;
    const int a = lay.addr((int)(span - 1)) + 1;
    return a + (a & 1);
}
template <int LOGMC, int TF>
inline size_t pk_inv_work_bytes(int64_t N) {
    using G = PkGeo<LOGMC, TF, 2>;
    size_t work = 4 * (size_t)G::MC * 2 * TF;
    const size_t ola = 4 * (size_t)TF * (size_t)(N + 6);
    if (ola > work) work = ola;
    // TMA path of the one-sided inverse: the S tile lands over this buffer, [Mc + 1 rows][TF + 2 slices]
    const size_t land = (size_t)(G::MC + 1) * (TF + 2) * 8;
    if (LOGMC == 8 && land > work) work = land;        // only mfft = 512 takes the TMA path (pow2_pk.cu); a larger
                                                       // carve-out elsewhere only shrinks L1 (mfft 4096: 2.57 -> 2.88 ms)
    return (work + 127) & ~(size_t)127;
}
// streaming overlap-add: chunk geometry for a launch (owned slices qh0..qh1 per signal)
inline void pk_inv_chunks(InvArgs<float> &a, const InvParams &p, int tf, int64_t batch, int64_t grid_hint) {
    a.qh0 = fdiv64(p.k0 + p.mid, p.H);
    a.qh1 = fdiv64(p.k1 - 1 + p.mid, p.H);
    const int64_t Q = a.qh1 - a.qh0 + 1;
#ifndef VEILS_CHUNK_HALO_FACTOR
#define VEILS_CHUNK_HALO_FACTOR 50
#endif
    int ct = (VEILS_CHUNK_HALO_FACTOR * a.halo + tf - 1) / tf;   // halo re-computation <= ~2 % of a chunk
    if (ct < 2) ct = 2;
    if (ct > 64) ct = 64;
    auto cps = [&](int c) { const int64_t lc = (int64_t)tf * c - a.halo; return (Q + lc - 1) / lc; };
    while (ct > 2 && batch * cps(ct) < 4 * grid_hint) --ct;   // enough chunks to balance the CTAs
    a.chunk_tiles = ct;
    a.cps = cps(ct);
}
template <int LOGMC, int TF>
inline size_t pk_inv_smem_bytes(int64_t N) {
    using G = PkGeo<LOGMC, TF, 2>;
    size_t b = pk_inv_work_bytes<LOGMC, TF>(N) + 2 * 4 * (size_t)((N + 3) & ~3LL);   // + 2 carry buffers
    if (G::STAB) b += 4 * (size_t)G::TWN + sizeof(ITab) * (size_t)G::MC;
    return b;
}
// shared memory of the TMA-landing + accumulator path (pow2_pk.cu): [work | landing][tables][accumulator]
template <int LOGMC, int TF>
inline size_t pk_inv_acc_work_bytes(int64_t N) {
    using G = PkGeo<LOGMC, TF, 2>;
    (void)N;
    size_t work = 4 * (size_t)G::MC * 2 * TF;
    const size_t land = (size_t)(G::MC + 1) * (TF + 2) * 8;
    if (land > work) work = land;
    return (work + 127) & ~(size_t)127;
}
template <int LOGMC, int TF>
inline size_t pk_inv_acc_smem_bytes(int64_t N) {
    using G = PkGeo<LOGMC, TF, 2>;
    size_t b = pk_inv_acc_work_bytes<LOGMC, TF>(N);
    if (G::STAB) b += 4 * (size_t)G::TWN + sizeof(ITab) * (size_t)G::MC;
    b += 4 * (size_t)((TF + 3) * (G::MC / 2 + 2)) + 32;
    return b;
}

// Calls fn.template run<LOGMC, TF, NSUB>() for mfft = 2^(logmc+1)
#define VEILS_PK_CASE(LM, TFV, NS) \
    case LM: return fn.template run<LM, TFV, NS>();
template <typename FN>
inline auto pk_dispatch(int logmc, FN &fn) -> decltype(fn.template run<4, 32, 4>()) {
    switch (logmc) {
        VEILS_PK_CASE(4, 32, 4) VEILS_PK_CASE(5, 32, 8) VEILS_PK_CASE(6, 32, 8) VEILS_PK_CASE(7, 32, 16)
        VEILS_PK_CASE(8, VEILS_PK_TF8, 16) VEILS_PK_CASE(9, 32, 32) VEILS_PK_CASE(10, 16, 64)
        VEILS_PK_CASE(11, 8, 128) VEILS_PK_CASE(12, 4, 256)          // 3-pass plans: 512 threads, 1 CTA/SM
    }
    return decltype(fn.template run<4, 32, 4>())(-1);
}

struct PkSmemQuery {
    int64_t N, H, ps; bool inverse;
    template <int LOGMC, int TF, int NSUB>
    long long run() {
        if (inverse) return (long long)pk_inv_smem_bytes<LOGMC, TF>(N);
        const TileLayout l = tile_layout(H, true);
        return (long long)pk_fwd_smem_bytes<LOGMC, TF>(N, fwd_xin_elems(N, H, TF, l));
    }
};

inline bool pk_supported(int64_t M, int64_t N, int64_t H, int64_t ps, bool inverse) {
    Cfg c;
    if (!pk_cfg_for(M, &c)) return false;
    if (N > M || H < 2 || H >= 65536) return false;
    if (inverse) {
        if (H > N) return false;
        if (inv_halo(N, H) >= c.tf) return false;
    } else if (!fwd_vec(ps, H, N)) {
        return false;                                  // paired tile reads need even ps, hop, m_num
    }
    PkSmemQuery q{N, H, ps, inverse};
    const long long need = pk_dispatch(c.logmc, q);
    return need > 0 && (size_t)need <= SMEM_LIMIT;
}

}  // namespace p2
}  // namespace veils
