// pow2_cfg.hpp -- host-side configuration of the power-of-two family, shared by the CUDA
// launchers (pow2_tile.cu) and the host emulation used by the CPU tests (tests/emu/pow2_emu.cu),
// so both run the very same (TF, NSUB) instantiation, twiddle table and tile layout.
#pragma once
#include <cmath>
#include <cstdint>

#include "pow2_tile_impl.cuh"

namespace veils {
namespace p2 {

// threads per CTA: one CTA per SM for mfft >= 2048 (work buffer 128 KiB) -> 512 threads there
#ifndef VEILS_P2_BIG_THREADS
#define VEILS_P2_BIG_THREADS 512
#endif
#define VEILS_P2_THREADS(logmc) ((logmc) >= 10 ? VEILS_P2_BIG_THREADS : 256)

#ifndef VEILS_F64_L8_TF
#define VEILS_F64_L8_TF 16
#endif
#ifndef VEILS_F64_L10_TF
#define VEILS_F64_L10_TF 8
#endif
struct Cfg {
    int logmc;   // log2(mfft / 2)
    int tf;      // slices per tile
    int nsub;    // butterfly groups worked on concurrently (threads = tf * nsub)
};

inline int ilog2_exact(int64_t v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((int64_t(1) << l) < v) ++l;
    return l;
}

// (TF, NSUB) per precision and size.  The work buffer Mc * TF complex is kept <= 128 KiB.
inline bool cfg_for(int64_t M, int precision, Cfg *c) {
    const int lm = ilog2_exact(M);
    if (lm < 5 || lm > 13) return false;              // mfft in [32, 8192]
    c->logmc = lm - 1;
    static const int tf32[13] = {0, 0, 0, 0, 32, 32, 32, 32, 32, 32, 16, 8, 4};
    static const int tf64[13] = {0, 0, 0, 0, 32, 32, 32, 32, VEILS_F64_L8_TF, 16, VEILS_F64_L10_TF, 4, 2};
    c->tf = precision == 0 ? tf32[c->logmc] : tf64[c->logmc];
    int np, r1, r2, r3;
    fac_of(c->logmc, &np, &r1, &r2, &r3);
    const int l1 = (1 << c->logmc) / r1;
    int nsub = VEILS_P2_THREADS(c->logmc) / c->tf;
    if (nsub > l1 / 2) nsub = l1 / 2;
    if (nsub < 1) nsub = 1;
    c->nsub = nsub;
    return true;
}

inline int64_t twiddle_count(int64_t M) {             // number of reals
    Cfg c;
    if (!cfg_for(M, 0, &c)) return 0;
    int np, r1, r2, r3;
    fac_of(c.logmc, &np, &r1, &r2, &r3);
    const int64_t mc = M / 2, l1 = mc / r1;
    return 2 * (mc + l1 + mc + 1);
}

inline void tw_exact(int64_t m, int64_t M, double *re, double *im) {   // exp(-2 pi i m / M)
    const long double two_pi = 6.283185307179586476925286766559005768L;
    const int64_t mm = ((m % M) + M) % M;
    if ((4 * mm) % M == 0) {
        static const double c4[4] = {1, 0, -1, 0}, s4[4] = {0, 1, 0, -1};
        const int q = (int)(4 * mm / M);
        *re = c4[q];
        *im = -s4[q];
        return;
    }
    const long double a = two_pi * (long double)mm / (long double)M;
    *re = (double)cosl(a);
    *im = (double)(-sinl(a));
}

// layout: [tw1: W_Mc^(j c) at j*R1+c][tw2: W_L1^(j2 c2) at j2*R2+c2][twp: W_M^k, k = 0..Mc]
inline void twiddle_fill(int64_t M, double *t) {
    Cfg c;
    if (!cfg_for(M, 0, &c)) return;
    int np, r1, r2, r3;
    fac_of(c.logmc, &np, &r1, &r2, &r3);
    const int64_t mc = M / 2, l1 = mc / r1;
    double *p = t;
    for (int64_t j = 0; j < l1; ++j)
        for (int64_t cc = 0; cc < r1; ++cc, p += 2) tw_exact(j * cc, mc, p, p + 1);
    if (np == 3) {
        const int64_t l2 = l1 / r2;
        for (int64_t j2 = 0; j2 < l2; ++j2)
            for (int64_t cc = 0; cc < r2; ++cc, p += 2) tw_exact(j2 * cc, l1, p, p + 1);
    } else {
        for (int64_t i = 0; i < l1; ++i, p += 2) { p[0] = 1; p[1] = 0; }
    }
    for (int64_t k = 0; k <= mc; ++k, p += 2) tw_exact(k, M, p, p + 1);
}

// staged-tile layout for hop H: the slice stride H + padw (in elements) must be
//   = 2 (mod 4) for paired (2-element) loads, odd for single-element loads.
inline TileLayout tile_layout(int64_t H, bool vec) {
    TileLayout l;
    l.H = (int)H;
    l.magic = (unsigned)((0x100000000ULL + (unsigned long long)H - 1) / (unsigned long long)H);
    if (vec) l.padw = (int)((2 - (H % 4) + 4) % 4);   // H even -> padw in {2, 0}
    else l.padw = (H % 2 == 0) ? 1 : 0;
    return l;
}
template <typename T> inline TileLayout fwd_layout(int64_t H, bool vec, int tf) {
    TileLayout l = tile_layout(H, vec);
    if (VEILS_P2_PADR && sizeof(T) == 8 && vec && tf < 8) {
        // narrow f64 tiles: a quarter-warp reads 8 / tf consecutive 16-byte pairs of each of its tf slices, so the
        // slice stride must be 128 / tf bytes past a multiple of 128 (ncu, c5: 8 instead of 4 wavefronts per LDS.128)
        const int want = 128 / tf;
        for (int pw = 0; pw < 16; pw += 2)
            if (((H + pw) * 8) % 128 == want) { l.padw = pw; break; }
    }
    return l;
}
inline bool fwd_vec(int64_t ps, int64_t H, int64_t N) { return ps % 2 == 0 && H % 2 == 0 && N % 2 == 0 && H >= 4; }

inline int64_t fwd_xin_elems(int64_t N, int64_t H, int tf, const TileLayout &l) {
    const int64_t span = (int64_t)(tf - 1) * H + N;
    return (span + (span / H + 1) * l.padw + 6 + 1) & ~int64_t(1);   // + zero pair used by padded table entries; even, so
                                                                      // that the tables behind a single buffer stay 16-byte aligned
}
// exact shared-memory footprints of the instantiated kernels (see pow2_tile.cu)
template <typename T, int LOGMC, int TF>
inline size_t fwd_smem_bytes_t(int64_t N, int64_t xin_elems) {
    using G = Geo<T, LOGMC, TF, 1>;
    size_t b = sizeof(cx<T>) * (size_t)(G::MCP + 1) * TF + sizeof(T) * (size_t)xin_elems * (G::DBUF ? 2 : 1);   // + Nyquist row
    if (G::STAB) b += sizeof(T) * ((size_t)G::TWN + (size_t)N);
    return b;
}
template <typename T, int LOGMC, int TF>
inline size_t inv_work_bytes_t(int64_t N) {
    using G = Geo<T, LOGMC, TF, 1>;
    size_t work = sizeof(cx<T>) * (size_t)G::MCP * TF;
    const size_t ola = sizeof(T) * (size_t)TF * (size_t)(N | 1);
    if (ola > work) work = ola;
    return (work + 15) & ~(size_t)15;
}
inline size_t inv_carry_elems(int64_t N, int64_t H) { return (size_t)((N - H + 3) & ~int64_t(3)); }   // one carry buffer (N - H), 16-byte multiple
template <typename T, int LOGMC, int TF>
inline size_t inv_smem_bytes_t(int64_t N, int64_t H) {
    using G = Geo<T, LOGMC, TF, 1>;
    size_t b = inv_work_bytes_t<T, LOGMC, TF>(N) + 2 * sizeof(T) * inv_carry_elems(N, H);   // + the two carry buffers
    if (G::STAB) b += sizeof(T) * ((size_t)G::TWN + (size_t)N);
    return b;
}
struct SmemQuery {
    int64_t N, H, ps; bool inverse;
    template <typename T, int LOGMC, int TF, int NSUB>
    long long run() {
        if (inverse) return (long long)inv_smem_bytes_t<T, LOGMC, TF>(N, H);
        const TileLayout l = fwd_layout<T>(H, fwd_vec(ps, H, N), TF);
        return (long long)fwd_smem_bytes_t<T, LOGMC, TF>(N, fwd_xin_elems(N, H, TF, l));
    }
};
inline int inv_halo(int64_t N, int64_t H) { return (int)((N + H - 1) / H) - 1; }
inline int inv_ostride(int64_t N) { return (int)(N | 1); }   // odd stride: lane writes conflict free

constexpr size_t SMEM_LIMIT = 227 * 1024;

template <typename T>
inline FwdArgs<T> make_fwd_args(const T *x, void *out, const FwdParams &p, const T *win_half, const T *tw, int tf) {
    FwdArgs<T> a;
    a.x = x; a.out = out; a.win = win_half; a.tw = tw;
    a.n = p.n; a.x_pitch = p.x_pitch; a.ldp = p.ldp; a.p0 = p.p0; a.P = p.P; a.k_offset = p.k_offset;
    a.ptiles = (p.P + tf - 1) / tf;
    a.N = (int)p.N; a.H = (int)p.H; a.mid = (int)p.mid; a.ps = (int)p.ps; a.F = (int)p.F;
    a.vec = fwd_vec(p.ps, p.H, p.N) ? 1 : 0;
    a.lay = fwd_layout<T>(p.H, a.vec != 0, tf);
    a.magic_hp = a.vec ? (unsigned)((0x100000000ULL + (unsigned long long)(p.H / 2) - 1) / (unsigned long long)(p.H / 2)) : 0u;
    a.padded = p.N < p.M ? 1 : 0;
    a.mode = p.mode;
    a.shift = p.mode == 1 ? (int)(p.M / 2) : 0;
    a.spectrogram = p.spectrogram;
    a.dbg = 0;
    a.fastgeo = (p.N == p.M && 4 * p.H == p.M && p.ps == p.M / 2) ? 1 : 0;
    a.fac2x = p.mode == 3 ? (T)p.fac2x : (T)1;
    return a;
}
inline int64_t fdiv64(int64_t a, int64_t b) {
    int64_t q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
template <typename T>
inline InvArgs<T> make_inv_args(const void *S, T *xo, const InvParams &p, const T *dual_scaled, const T *tw, int tf) {
    InvArgs<T> a;
    a.S = S; a.xo = xo; a.dual = dual_scaled; a.tw = tw;
    a.P = p.P; a.ldp = p.ldp; a.out_pitch = p.out_pitch; a.p_min = p.p_min;
    a.k0 = p.k0; a.k1 = p.k1; a.q0 = p.q0; a.q1 = p.q1;
    a.N = (int)p.N; a.H = (int)p.H; a.mid = (int)p.mid; a.ps = (int)p.ps; a.F = (int)p.F;
    a.mode = p.mode;
    a.shift = p.mode == 1 ? (int)(p.M / 2) : 0;
    a.rfac2x = p.mode == 3 ? (T)(1.0 / p.fac2x) : (T)1;
    a.magic_h = (unsigned)((0x100000000ULL + (unsigned long long)p.H - 1) / (unsigned long long)p.H);
    a.halo = inv_halo(p.N, p.H);
    a.ostride = inv_ostride(p.N);
    a.pair = 0;
    a.dbg = 0;
    const int own = tf - a.halo;
    const int64_t qh0 = fdiv64(p.k0 + p.mid, p.H), qh1 = fdiv64(p.k1 - 1 + p.mid, p.H);
    a.ntiles = (qh1 - qh0 + 1 + own - 1) / own;
    a.qa0 = qh0 - a.halo;
    return a;
}

// streaming overlap-add: chunk geometry of a launch (owned slices qh0 .. qh1 per signal; a chunk = chunk_tiles
// consecutive tiles walked by one CTA, the first of which re-computes the `halo` slices before its first owned one)
template <typename T>
inline void tile_inv_chunks(InvArgs<T> &a, const InvParams &p, int tf, int64_t batch, int64_t grid_hint) {
    a.qh0 = fdiv64(p.k0 + p.mid, p.H);
    a.qh1 = fdiv64(p.k1 - 1 + p.mid, p.H);
    const int64_t Q = a.qh1 - a.qh0 + 1;
    int ct = (50 * a.halo + tf - 1) / tf;              // halo re-computation <= ~2 % of a chunk
    const int ctmin = a.halo / tf + 1;                 // at least one owned slice per chunk (small problems: one tile
                                                       // per chunk = independent tiles, the shortest critical path)
    if (ct < ctmin) ct = ctmin;
    if (ct > 256) ct = 256;
    auto cps = [&](int c) { const int64_t lc = (int64_t)tf * c - a.halo; return (Q + lc - 1) / lc; };
    while (ct > ctmin && batch * cps(ct) < 4 * grid_hint) --ct;   // enough chunks to balance the CTAs
    a.chunk_tiles = ct;
    a.cps = cps(ct);
}
// Calls fn.template run<T, LOGMC, TF, NSUB>() for the configuration of (M, precision).
#define VEILS_P2_CASE(T, LM, TFV)                                             \
    case LM: {                                                                \
        constexpr int l1_ = (1 << LM) / Fac<LM>::R1;                          \
        constexpr int ns0_ = VEILS_P2_THREADS(LM) / TFV;                      \
        constexpr int ns1_ = ns0_ > l1_ / 2 ? l1_ / 2 : ns0_;                 \
        constexpr int ns_ = ns1_ < 1 ? 1 : ns1_;                              \
        return fn.template run<T, LM, TFV, ns_>();                            \
    }
template <typename FN>
inline auto dispatch(int precision, int logmc, FN &fn) -> decltype(fn.template run<float, 4, 32, 2>()) {
    if (precision == 0) {
        switch (logmc) {
            VEILS_P2_CASE(float, 4, 32) VEILS_P2_CASE(float, 5, 32) VEILS_P2_CASE(float, 6, 32)
            VEILS_P2_CASE(float, 7, 32) VEILS_P2_CASE(float, 8, 32) VEILS_P2_CASE(float, 9, 32)
            VEILS_P2_CASE(float, 10, 16) VEILS_P2_CASE(float, 11, 8) VEILS_P2_CASE(float, 12, 4)
        }
    } else {
        switch (logmc) {
            VEILS_P2_CASE(double, 4, 32) VEILS_P2_CASE(double, 5, 32) VEILS_P2_CASE(double, 6, 32)
            VEILS_P2_CASE(double, 7, 32) VEILS_P2_CASE(double, 8, VEILS_F64_L8_TF) VEILS_P2_CASE(double, 9, 16)
            VEILS_P2_CASE(double, 10, VEILS_F64_L10_TF) VEILS_P2_CASE(double, 11, 4) VEILS_P2_CASE(double, 12, 2)
        }
    }
    return decltype(fn.template run<float, 4, 32, 2>())(-1);
}

inline bool supported(int64_t M, int64_t N, int64_t H, int64_t ps, int precision, bool inverse) {
    Cfg c;
    if (!cfg_for(M, precision, &c)) return false;
    if (N > M || H < 2 || H >= 65536) return false;
    if (inverse) {
        if (H > N) return false;
        // (streaming overlap-add: tiles may be narrower than the overlap)
    }
    SmemQuery q{N, H, ps, inverse};
    const long long need = dispatch(precision, c.logmc, q);
    return need > 0 && (size_t)need <= SMEM_LIMIT;
}

}  // namespace p2
}  // namespace veils
