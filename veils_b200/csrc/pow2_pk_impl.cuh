// pow2_pk_impl.cuh -- f32 "packed slice-lane" kernels: the pow2 tile family with TWO time slices
// per lane, the same component of both kept in one f32x2 register pair, so every butterfly
// add / mul / fma is a single FADD2 / FMUL2 / FFMA2 (see f32x2.cuh).
//
//   * thread t = (sub, l): l = t % LN (LN = TF/2) owns slices l and l + LN of the tile
//     (this pairing keeps the 8-byte tile reads of the 16 lanes of a half-warp on 32 distinct
//     banks; slices 2l, 2l+1 would give 2-way conflicts);
//   * work buffer is planar: work[(e*2 + comp)*TF + 2*l + s]  (comp 0 = re, 1 = im; s = slice
//     l + s*LN), so one 8-byte access moves one component of both slices;
//   * twiddle tables hold (wr, wi, -wr, -wi): no packed negation is ever needed (ptxas does not
//     fold negations into FFMA2 operands);
//   * the mirrored groups (g, G-g) of the real-FFT split live in the two HALVES of a warp row
//     pair (sub, sub^1) and swap one value per bin with __shfl_xor -- each thread keeps ONE
//     group in registers (<= 128 registers, 16 warps per SM) and produces ONE output row per
//     bin pair; both halves run the same instruction stream because the table entry of bin
//     Mc-k is (-wr, wi) of bin k;
//   * per-plan tile table ptab[m] = (tile offset of sample pair m, -, w[2m'], w[2m'+1]) turns
//     frame extraction + roll + zero padding + windowing into one 16-byte table read.
// The AoS <-> packed transposition is free: done by the first scalar op after a load (window
// multiply / Hermitian merge add) and the last scalar op before a store.
// Phase functions compile for the device and for the host emulation (tests/emu).
#pragma once
#include "pow2_tile_impl.cuh"

namespace veils {
namespace p2 {

struct alignas(8) fl2 { float x, y; };
struct alignas(16) fl4 { float x, y, z, w; };
struct alignas(16) PTab { int aoff; int pad; float w0, w1; };   // one sample pair of a slice
struct alignas(16) ITab { int o0, o1; float d0, d1; };           // inverse: ola offsets + dual/M of a pair

template <int LOGMC, int TF_, int NSUB_>
struct PkGeo {
    using F = Fac<LOGMC>;
    static constexpr int MC = 1 << LOGMC, M = 2 * MC;
    static constexpr int TF = TF_, LN = TF_ / 2, NSUB = NSUB_, NT = LN * NSUB_;
    static constexpr int NP = F::NP, R1 = F::R1, R2 = F::R2, R3 = F::R3;
    static constexpr int RL = (NP == 2) ? R2 : R3;
    static constexpr int L1 = MC / R1;
    static constexpr int L2 = (NP == 3) ? L1 / R2 : 1;
    static constexpr int G = MC / RL;
    // twiddle table: entries of 4 floats (wr, wi, -wr, -wi):
    //   [tw1: MC][tw2: L1][twp: MC + 1][tw1x: MC][tw2x: L1]   (pow2_pk_cfg.hpp: separate forward and
    //   inverse tables; the x-tables serve the "reversed" groups of the second warp halves)
    //   [tw1p: MC/2]  inverse, paired pass 1: (Re W^(j0 c), Re W^(j1 c), Im W^(j0 c), Im W^(j1 c)),
    //                 W = W_Mc, for the group pair (j0, j1) of w = 0 .. L1/2-1 at entry w*R1 + c
    static constexpr int TW1 = 0, TW2 = 4 * MC, TWP = 4 * (MC + L1), TW1X = 4 * (2 * MC + L1 + 1),
                         TW2X = 4 * (3 * MC + L1 + 1), TW1P = 4 * (3 * MC + 2 * L1 + 1),
                         TWN = 4 * (3 * MC + 2 * L1 + 1 + MC / 2);
    static constexpr int TWN_FWD = TW1X;              // the forward kernels read tw1, tw2, twp only
    // last-pass groups > G/2 are handled by the second halves (h = 1) of the warp row pairs: they
    // need their DFT outputs in REVERSED order (their partner's u[d] mirrors their u[RL-1-d]).
    // DFT(y)[e] = X[RL-1-e] for y[n] = x[(-n) mod RL] W_RL^n, so the producer pass stores such a
    // group reversed and folds W_RL^(-j) into its twiddle: free at run time.
    static constexpr bool REVG = (NP == 2);           // 2-pass plans: pass 1 reads the tile, no hazard
    VHD static bool rev_group(int g) { return REVG && g > G / 2; }
    VHD static int goff(int g) { return (NP == 2) ? g * L1 : (g % R1) * L1 + (g / R1) * L2; }
    static constexpr bool STAB = LOGMC <= 8;          // plan tables in shared memory
    static constexpr bool DBUF = LOGMC <= 7;          // double-buffered signal tile (measured for mfft = 512:
                                                      // no gain, 0.818 vs 0.803 ms -- the next tile is
                                                      // requested after pass 1 instead)
    VHD static fl4 tw4(const float *q) {
#if defined(__CUDA_ARCH__)
        if (!STAB) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(q));
            fl4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
            return r;
        }
#endif
        return *reinterpret_cast<const fl4 *>(q);
    }
    VHD static fl2 tw2(const float *q) {               // first half (wr, wi) of a table entry
#if defined(__CUDA_ARCH__)
        if (!STAB) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(q));
            fl2 r; r.x = v.x; r.y = v.y;
            return r;
        }
#endif
        return *reinterpret_cast<const fl2 *>(q);
    }
    VHD static PTab pt4(const PTab *q) {
#if defined(__CUDA_ARCH__)
        if (!STAB) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(q));
            PTab r; r.aoff = v.x; r.pad = v.y; r.w0 = __int_as_float(v.z); r.w1 = __int_as_float(v.w);
            return r;
        }
#endif
        return *q;
    }
    VHD static fl2 dw2(const ITab *q) {                // (d0, d1) of an inverse table entry
#if defined(__CUDA_ARCH__)
        if (!STAB) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(&q->d0));
            fl2 r; r.x = v.x; r.y = v.y;
            return r;
        }
#endif
        return *reinterpret_cast<const fl2 *>(&q->d0);
    }
    VHD static ITab it4(const ITab *q) {
#if defined(__CUDA_ARCH__)
        if (!STAB) {
            const int4 v = __ldg(reinterpret_cast<const int4 *>(q));
            ITab r; r.o0 = v.x; r.o1 = v.y; r.d0 = __int_as_float(v.z); r.d1 = __int_as_float(v.w);
            return r;
        }
#endif
        return *q;
    }
};

// Keeps a running pointer a running pointer: without it the compiler materialises all 16
// addresses of an unrolled `p += step` chain up front and spills them.
template <typename P> VHD void keep(P *&p) {
#if defined(__CUDA_ARCH__)
    asm volatile("" : "+l"(p));
#else
    (void)p;
#endif
}
using C2 = cx<f2>;
VHD C2 c2zero() { C2 r; r.x = bc(0.f); r.y = bc(0.f); return r; }
VHD C2 wload(const float *w, int TF) {
    C2 r;
    r.x = *reinterpret_cast<const f2 *>(w);
    r.y = *reinterpret_cast<const f2 *>(w + TF);
    return r;
}
VHD void wstore(float *w, int TF, C2 v) {
    *reinterpret_cast<f2 *>(w) = v.x;
    *reinterpret_cast<f2 *>(w + TF) = v.y;
}

// value of the partner thread (lane ^ LN).  Device: warp shuffle.  Host emulation: the partner's
// copy, handed in by the harness.
// lanes of the (sub, sub^1) pair this thread belongs to
template <int LN>
VHD unsigned pair_mask(int tid) {
    if (LN >= 16) return 0xffffffffu;
    const unsigned m = (1u << (2 * LN)) - 1u;
    return m << (((tid & 31) / (2 * LN)) * (2 * LN));
}
template <int LN>
VHD C2 xchg(C2 v, const C2 *partner_copy, unsigned mask) {
#if defined(__CUDA_ARCH__)
    (void)partner_copy;
    C2 r;
    r.x.v = __shfl_xor_sync(mask, v.x.v, LN);
    r.y.v = __shfl_xor_sync(mask, v.y.v, LN);
    return r;
#else
    (void)v; (void)mask;
    return *partner_copy;
#endif
}
struct Raw {                 // one bin of both slices as loaded: {re0, im0} {re1, im1}
    float r0, i0, r1, i1;
};
template <int LN>
VHD Raw xchg_raw(Raw v, const Raw *partner_copy, unsigned mask) {
#if defined(__CUDA_ARCH__)
    (void)partner_copy;
    Raw r;
    r.r0 = __shfl_xor_sync(mask, v.r0, LN);
    r.i0 = __shfl_xor_sync(mask, v.i0, LN);
    r.r1 = __shfl_xor_sync(mask, v.r1, LN);
    r.i1 = __shfl_xor_sync(mask, v.i1, LN);
    return r;
#else
    (void)v; (void)mask;
    return *partner_copy;
#endif
}

// ================================================================== FORWARD (packed)
struct PkFwdCtx {
    FwdArgs<float> a;
    const float *tw;
    const PTab *ptab;        // [MC] tile offset + window of sample pair m
    float *xin;
    float *work;             // planar [MC][2][TF]
    int64_t b, pt;
};

template <int LOGMC, int TF_, int NSUB_>
struct PkFwd : PkGeo<LOGMC, TF_, NSUB_> {
    using GEO = PkGeo<LOGMC, TF_, NSUB_>;
    using GEO::MC; using GEO::M; using GEO::TF; using GEO::LN; using GEO::NSUB; using GEO::NT;
    using GEO::NP; using GEO::R1; using GEO::R2; using GEO::R3; using GEO::RL; using GEO::L1;
    using GEO::L2; using GEO::G;
    static constexpr int WPT = (G / 2 + NSUB / 2 - 1) / (NSUB / 2);     // last-pass pairs per sub pair

    // staging of the signal tile (paired cp.async, zero fill outside [0, n)); the two floats
    // after the staged span are the "zero pair" that padded table entries point to.
    VHD static void fill(const PkFwdCtx &c, int tid) {
        const FwdArgs<float> &p = c.a;
        const int H = p.H;
        const int span = (TF - 1) * H + p.N;
        const int hs = H + p.lay.padw;
        const int64_t s0 = (p.p0 + c.pt * TF) * (int64_t)H + p.k_offset - p.mid;
        const float *xb = p.x + c.b * p.x_pitch;
        if ((((c.b * p.x_pitch + s0) & 1) == 0) && ((p.n & 1) == 0) && ((((size_t)p.x) & 7) == 0)) {
            const int np = span >> 1, hp = H >> 1;
            if (s0 >= 0 && s0 + span <= p.n && hp <= NT && NT % hp == 0) {
                // interior tile, thread grid (NT/hp rows) x (hp pairs): purely incremental addressing
                const int c2 = tid % hp, r0 = tid / hp, rstep = NT / hp;
                float *dst = c.xin + r0 * hs + 2 * c2;
                const float *src = xb + s0 + (int64_t)r0 * H + 2 * c2;
                int e = tid;
                for (; e < np; e += NT) {
                    cp_async_zfill<8>(dst, src, true);
                    dst += rstep * hs;
                    src += rstep * H;
                }
                return;
            }
            for (int e = tid; e < np; e += NT) {
                const int64_t s = s0 + 2 * (int64_t)e;
                const bool ok = s >= 0 && s < p.n;
                const int r = (int)mulhi_u32((unsigned)e, p.magic_hp);
                cp_async_zfill<8>(c.xin + r * hs + 2 * (e - r * hp), ok ? xb + s : xb, ok);
            }
            return;
        }
        for (int e = tid; e < span; e += NT) {
            const int64_t s = s0 + e;
            const bool ok = s >= 0 && s < p.n;
            const int r = e / H;
            cp_async_zfill<4>(c.xin + r * hs + (e - r * H), ok ? xb + s : xb, ok);
        }
    }

    // ---- pass 1: one group per thread and iteration
    VHD static void pass1(const PkFwdCtx &c, int tid) {
        if (c.a.padded) pass1_t<true>(c, tid);
        else pass1_t<false>(c, tid);
    }
    template <bool PADDED>
    VHD static void pass1_t(const PkFwdCtx &c, int tid) {
        const int sub = tid / LN;
        for (int j = sub; j < L1; j += NSUB) {
            C2 v[R1];
            if (LOGMC >= 8 && c.a.fastgeo) pass1_in_fast(c, tid, j, v);
            else pass1_in<PADDED>(c, tid, j, v);
            pass1_out(c, tid, j, v);
        }
    }
    // pass 1 of group j, first half: tile reads, window, DFT_R1 -- does not touch the work buffer
    VHD static void pass1_in(const PkFwdCtx &c, int tid, int j, C2 (&v)[R1]) {
        if (LOGMC >= 8 && c.a.fastgeo) pass1_in_fast(c, tid, j, v);      // (mfft 256: 3 % slower with it)
        else if (c.a.padded) pass1_in<true>(c, tid, j, v);
        else pass1_in<false>(c, tid, j, v);
    }
    // default geometry (m_num = mfft = 4 hop, roll = m_num / 2): sample pair m = j + L1 a of a slice sits
    // at tile offset 2j + C(a) with a compile-time C(a), its window pair at ptab[m].w0 -- every read is
    // base + immediate, nothing waits for a table entry first
    VHD static void pass1_in_fast(const PkFwdCtx &c, int tid, int j, C2 (&v)[R1]) {
        constexpr int HC = MC / 2;
        constexpr int PADW = (2 - HC % 4 + 4) % 4;
        const int l = tid % LN;
        const int hs = HC + PADW;
        const float *xf0 = c.xin + l * hs + 2 * j, *xf1 = xf0 + LN * hs;
        const float *wj = &c.ptab[j].w0;
#pragma unroll
        for (int a = 0; a < R1; ++a) {
            constexpr int dummy = 0; (void)dummy;
            const int ca = (2 * L1 * a + M / 2) % M;                   // compile-time after unrolling
            const int off = ca + (ca / HC) * PADW;
            const fl2 x0 = *reinterpret_cast<const fl2 *>(xf0 + off);
            const fl2 x1 = *reinterpret_cast<const fl2 *>(xf1 + off);
            const fl2 w = *reinterpret_cast<const fl2 *>(wj + 4 * L1 * a);
            v[a].x = mk2(x0.x * w.x, x1.x * w.x);
            v[a].y = mk2(x0.y * w.y, x1.y * w.y);
        }
        Dft<R1, false, f2>::run(v);
    }
    template <bool PADDED>
    VHD static void pass1_in(const PkFwdCtx &c, int tid, int j, C2 (&v)[R1]) {
        const FwdArgs<float> &p = c.a;
        const int l = tid % LN;
        const int hs = p.H + p.lay.padw;
        const float *xf0 = c.xin + l * hs, *xf1 = xf0 + LN * hs;
        {
            const PTab *pj = c.ptab + j;
#pragma unroll
            for (int a = 0; a < R1; ++a) {
                const PTab e = GEO::pt4(pj + L1 * a);
                if (PADDED && e.pad) {                         // sample pair beyond m_num: exact zeros
                    v[a] = c2zero();
                    continue;
                }
                const fl2 x0 = *reinterpret_cast<const fl2 *>(xf0 + e.aoff);
                const fl2 x1 = *reinterpret_cast<const fl2 *>(xf1 + e.aoff);
                v[a].x = mk2(x0.x * e.w0, x1.x * e.w0);        // scalar multiplies transpose for free
                v[a].y = mk2(x0.y * e.w1, x1.y * e.w1);
            }
            Dft<R1, false, f2>::run(v);
        }
    }
    // second half: twiddle W_Mc^(j c) and the stores into the work buffer
    VHD static void pass1_out(const PkFwdCtx &c, int tid, int j, const C2 (&v)[R1]) {
        const int l = tid % LN;
        {
            float *wj = c.work + (j * 2) * TF + 2 * l;
            // 2-pass plans: column cc IS last-pass group cc; reversed groups are stored at row
            // (L1 - j) mod L1 (the extra W_L1^(-j) is already in the forward tw1 table)
            float *wr = c.work + (((L1 - j) & (L1 - 1)) * 2) * TF + 2 * l;
            const float *twj = c.tw + GEO::TW1 + 4 * j * R1;
            wstore(wj, TF, v[0]);
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) {
                const fl2 t = GEO::tw2(twj + 4 * cc);          // (wr, wi): 8-byte broadcast read
                float *dst = GEO::rev_group(cc) ? wr : wj;
                wstore(dst + (L1 * cc * 2) * TF, TF, cmulw(v[cc], bc(t.x), bc(t.y), bc(-t.y)));
            }
        }
    }

    // middle pass (3-pass plans), in place.  (The reversed-group trick is NOT applied here: a
    // reversed store would land in a row another thread of the block still has to read.)
    VHD static void pass_mid(const PkFwdCtx &c, int tid) {
        if (NP != 3) return;
        const int l = tid % LN, sub = tid / LN;
        for (int u = sub; u < R1 * L2; u += NSUB) {
            const int cb = u / L2, j2 = u % L2;
            float *wb = c.work + ((cb * L1 + j2) * 2) * TF + 2 * l;
            const float *twj = c.tw + GEO::TW2 + 4 * j2 * R2;
            C2 v[R2];
#pragma unroll
            for (int a = 0; a < R2; ++a) v[a] = wload(wb + (L2 * a * 2) * TF, TF);
            Dft<R2, false, f2>::run(v);
            wstore(wb, TF, v[0]);
#pragma unroll
            for (int cc = 1; cc < R2; ++cc) {
                const fl4 t = GEO::tw4(twj + 4 * cc);
                wstore(wb + (L2 * cc * 2) * TF, TF, cmulw(v[cc], bc(t.x), bc(t.y), bc(t.w)));
            }
        }
    }

    struct Out {             // per-thread output cursor
        char *ob;            // byte pointer of (row 0, column of slice l)
        int64_t rb;          // row pitch in bytes
        bool live0, live1;   // slice l / l + LN inside [0, P)
        int dbg;
    };
    // STAGE: the tile is staged in shared memory (in place, over the rows this thread has read) and
    // leaves through the TMA engine (pk_fwd_kernel); q then points into the work buffer.
    template <bool SPEC, bool STAGE>
    VHD static void put_at(const Out &o, char *q, float r0, float i0, float r1, float i1) {
#ifdef VEILS_ABLATE
        if ((o.dbg & 1) && r0 != 12345.678f) return;      // ablation: no global stores
#endif
        if (STAGE) {
            if (SPEC) {
                *reinterpret_cast<float *>(q) = r0 * r0 + i0 * i0;
                *reinterpret_cast<float *>(q + 4 * LN) = r1 * r1 + i1 * i1;
            } else {
                fl2 v; v.x = r0; v.y = i0; *reinterpret_cast<fl2 *>(q) = v;
                fl2 u; u.x = r1; u.y = i1; *reinterpret_cast<fl2 *>(q + 8 * LN) = u;
            }
            return;
        }
        if (SPEC) {
            if (o.live0) *reinterpret_cast<float *>(q) = r0 * r0 + i0 * i0;
            if (o.live1) *reinterpret_cast<float *>(q + 4 * LN) = r1 * r1 + i1 * i1;
        } else {
            if (o.live0) { fl2 v; v.x = r0; v.y = i0; *reinterpret_cast<fl2 *>(q) = v; }
            if (o.live1) { fl2 v; v.x = r1; v.y = i1; *reinterpret_cast<fl2 *>(q + 8 * LN) = v; }
        }
    }
    // mode epilogue of one-sided bin k (src/lib.rs:367-408) and the store(s).
    // MODE: 0 twosided, 1 centered, 2 onesided, 3 onesided2X.  EDGE: k may be 0 or Mc.
    // q: byte pointer of row k (one-sided layouts only).
    // byte offset of one-sided bin k in the staged tile: the RL rows of last-pass group g = k mod G
    // are contiguous at the group's own work rows (row pitch TF * esz); bin Mc has an extra row.
    template <bool SPEC>
    VHD static int srow(int k) {
        if (k >= MC) return MC * TF * 8;
        return GEO::goff(k & (G - 1)) * (TF * 8) + (k / G) * (TF * (SPEC ? 4 : 8));
    }
    template <bool SPEC, bool STAGE>
    VHD static char *rowptr(const Out &o, int k) {
        return STAGE ? o.ob + srow<SPEC>(k) : o.ob + k * o.rb;
    }
    template <int MODE, bool SPEC, bool EDGE, bool STAGE>
    VHD static void store_bin(const FwdArgs<float> &p, const Out &o, char *q, int k, float r0, float i0, float r1, float i1) {
        const bool edge = EDGE && (k == 0 || k == MC);
        if (edge) { i0 = 0.f; i1 = 0.f; }
        if (MODE >= 2) {
            if (MODE == 3 && !edge) { r0 *= p.fac2x; i0 *= p.fac2x; r1 *= p.fac2x; i1 *= p.fac2x; }
            put_at<SPEC, STAGE>(o, q, r0, i0, r1, i1);
            return;
        }
        const int sh = MODE == 1 ? MC : 0;                                    // fftshift for centered
        put_at<SPEC, false>(o, o.ob + ((k + sh) & (M - 1)) * o.rb, r0, i0, r1, i1);
        if (!edge) put_at<SPEC, false>(o, o.ob + ((M - k + sh) & (M - 1)) * o.rb, r0, -i0, r1, -i1);
    }

    // own = Z[k_own]/2 of this thread's group, rcv = Z[Mc - k_own]/2 (partner half or own
    // registers); t = table entry of k_own (= (-wr, wi) of the mirrored bin).  Produces X[k_own].
    template <int MODE, bool SPEC, bool EDGE, bool STAGE>
    VHD static void emit1(const FwdArgs<float> &p, const Out &o, char *q, int k, fl4 t, C2 own, C2 rcv) {
        const f2 Ex = own.x + rcv.x, Oy = own.y + rcv.y, ey = own.y - rcv.y, ox = own.x - rcv.x;
        const f2 ax = vfma(bc(t.y), ox, Ex);                 // + wi * ox
        const f2 ay = vfma(bc(t.y), Oy, ey);                 // + wi * Oy
        // last fma scalar: moves (re,re)(im,im) into {re,im} per slice for free
        const float r0 = t.x * lo(Oy) + lo(ax), r1 = t.x * hi(Oy) + hi(ax);
        const float i0 = t.z * lo(ox) + lo(ay), i1 = t.z * hi(ox) + hi(ay);
        store_bin<MODE, SPEC, EDGE, STAGE>(p, o, q, k, r0, i0, r1, i1);
    }
    // both outputs from values held by one thread (self-mirrored groups 0 and G/2)
    template <int MODE, bool SPEC, bool STAGE>
    VHD static void emit2(const FwdArgs<float> &p, const float *twp, const Out &o, int k, C2 A, C2 Zp, bool both) {
        emit1<MODE, SPEC, true, STAGE>(p, o, rowptr<SPEC, STAGE>(o, k), k, GEO::tw4(twp + 4 * k), A, Zp);
        if (both)
            emit1<MODE, SPEC, true, STAGE>(p, o, rowptr<SPEC, STAGE>(o, MC - k), MC - k, GEO::tw4(twp + 4 * (MC - k)), Zp, A);
    }

    // ---- last pass, part 1: this thread's group -> registers
    //      sub = 2w + h:  w = 0 -> groups (0, G/2);  w >= 1 -> groups (w, G - w)
    VHD static int group_of(int w, int h) { return h == 0 ? w : (w == 0 ? G / 2 : G - w); }
    VHD static void last_dft(const PkFwdCtx &c, int tid, int it, C2 (&u)[RL]) {
        const int l = tid % LN, sub = tid / LN;
        const int w = (sub >> 1) + it * (NSUB / 2), h = sub & 1;
        if (w >= G / 2) return;
        const float *w1 = c.work + (GEO::goff(group_of(w, h)) * 2) * TF + 2 * l;
#pragma unroll
        for (int d = 0; d < RL; ++d) u[d] = wload(w1 + (d * 2) * TF, TF);
        Dft<RL, false, f2>::run(u);
    }
    // ---- last pass, part 2: exchange with the partner half, split, epilogue, store.
    //      pu = the partner thread's u[] (host emulation only)
    template <int MODE, bool SPEC, bool STAGE = false>
    VHD static void last_emit_t(const PkFwdCtx &c, int tid, int it, const C2 (&u)[RL], const C2 *pu) {
        const FwdArgs<float> &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        const int w = (sub >> 1) + it * (NSUB / 2), h = sub & 1;
        if (w >= G / 2) return;
        const int64_t pl = c.pt * TF + l;
        const int64_t esz = SPEC ? 4 : 8;
        Out o;
        if (STAGE) {
            o.live0 = o.live1 = true;                      // the TMA store clips at P
            o.rb = TF * esz;
            o.ob = reinterpret_cast<char *>(c.work) + l * esz;
        } else {
            o.live0 = pl < p.P;
            o.live1 = pl + LN < p.P;
            o.rb = p.ldp * esz;
            o.ob = reinterpret_cast<char *>(p.out) + ((c.b * p.F) * p.ldp + pl) * esz;
        }
        o.dbg = p.dbg;
        const float *twp = c.tw + GEO::TWP;
        const int g = group_of(w, h);
        const unsigned mask = pair_mask<LN>(tid);
        if (w != 0) {
            // bin of u[d]: k = g + G d.  Partner of half 0's u[d] is half 1's u[RL-1-d]: half 1 walks
            // its bins backwards, so row pointer, table pointer and k advance by a per-thread stride.
            int k = g + (h == 0 ? 0 : G * (RL - 1));
            const int kstep = h == 0 ? G : -G;
            char *q = rowptr<SPEC, STAGE>(o, k);
            const int64_t qstep = STAGE ? (h == 0 ? TF * esz : -TF * esz) : kstep * o.rb;
            const float *tq = twp + 4 * k;
            const int tstep = 4 * kstep;
#pragma unroll
            for (int d = 0; d < RL; ++d) {
                // REVG: half 1 already holds its group reversed (u[d] mirrors the partner's u[d])
                const C2 mine = (GEO::REVG || h == 0) ? u[d] : u[RL - 1 - d];
#ifdef VEILS_ABLATE
                const C2 rcv = (p.dbg & 2) ? mine : xchg<LN>(mine, nullptr, mask);   // ablation: no shuffles
#else
                const C2 rcv = xchg<LN>(mine, pu ? pu + ((GEO::REVG || h != 0) ? d : RL - 1 - d) : nullptr, mask);
#endif
                emit1<MODE, SPEC, false, STAGE>(p, o, q, k, GEO::tw4(tq), mine, rcv);
                k += kstep;
                q += qstep;
                tq += tstep;
            }
        } else {
            // self-mirrored groups: half 0 holds group 0 (u[d] <-> u[(RL-d) mod RL]), half 1 group G/2
            // (u[d] <-> u[RL-1-d]).  One instruction stream for both halves (the mirror operand is a
            // register select), two outputs per pair -- same amount of work as the other warps, which
            // would otherwise wait at the next barrier for a warp running two code paths in turn.
#pragma unroll
            for (int t = 0; t < RL / 2; ++t) {
                const C2 a = u[t];
                const C2 bm = h == 0 ? u[(RL - t) & (RL - 1)] : u[RL - 1 - t];
                const int k = h == 0 ? G * t : G / 2 + G * t;
                emit2<MODE, SPEC, STAGE>(p, twp, o, k, a, bm, true);      // t = 0, h = 0: X[0] and X[Mc]
            }
            if (h == 0) emit2<MODE, SPEC, STAGE>(p, twp, o, MC / 2, u[RL / 2], u[RL / 2], false);
        }
    }
    // ================= staged ("paired") last pass: one slice per lane, no partner exchange ===========
    // Thread (w, lam): lam = tid % TF is the slice, w selects the mirrored group pair (w, G - w) -- w = 0:
    // (0, G/2).  The two groups ride in the two halves of one f32x2 register, so ONE packed DFT
    // transforms both and the bins k and Mc - k of the real-FFT split (src/lib.rs:367-375 computes
    // them with a full-length FFT) meet in the same thread: no shuffles, no selects, no redundant
    // adds.  The finished bins overwrite the group's own work rows (TMA staging, pk_fwd_kernel).
    static constexpr int NW2 = NT / TF;                                  // group pairs in flight
    static constexpr int WPT2 = (G / 2 + NW2 - 1) / NW2;
    // lane lam handles slice lam; the pass-1 lanes interleave their two slices (l, l + LN) in the
    // work buffer, so slice lam sits at position wpos(lam).  Reads: 32 lanes on 32 distinct banks;
    // staged stores: 32 consecutive 8-byte slots -- both conflict free.
    VHD static int wpos(int lam) { return lam < LN ? 2 * lam : 2 * (lam - LN) + 1; }
    VHD static void last_dft2(const PkFwdCtx &c, int tid, int it, C2 (&u)[RL]) {
        const int lam = tid % TF, w = tid / TF + it * NW2;
        if (w >= G / 2) return;
        const float *a0 = c.work + (GEO::goff(group_of(w, 0)) * 2) * TF + wpos(lam);
        const float *a1 = c.work + (GEO::goff(group_of(w, 1)) * 2) * TF + wpos(lam);
#pragma unroll
        for (int d = 0; d < RL; ++d) {
            u[d].x = mk2(a0[(d * 2) * TF], a1[(d * 2) * TF]);
            u[d].y = mk2(a0[(d * 2 + 1) * TF], a1[(d * 2 + 1) * TF]);
        }
        Dft<RL, false, f2>::run(u);
    }
    // one slice, one mirrored bin pair: A = Z[k]/2, B = Z[Mc-k]/2, (wr, wi) = W_M^k.
    //   X[k] = (Ex + P, ey + Q),  X[Mc-k] = (Ex - P, Q - ey)   with
    //   Ex = Ar + Br, Oy = Ai + Bi, ey = Ai - Bi, ox = Ar - Br, P = wi ox + wr Oy, Q = wi Oy - wr ox
    template <int MODE, bool SPEC>
    VHD static void put1(const FwdArgs<float> &p, char *q, int k, float re, float im) {
        if (k == 0 || k == MC) im = 0.f;                                  // src/lib.rs:378-384
        else if (MODE == 3) { re *= p.fac2x; im *= p.fac2x; }             // :397-403
        if (SPEC) *reinterpret_cast<float *>(q) = re * re + im * im;
        else { fl2 v; v.x = re; v.y = im; *reinterpret_cast<fl2 *>(q) = v; }
    }
    template <int MODE, bool SPEC>
    VHD static void emit_pair(const FwdArgs<float> &p, const float *twp, char *ob, int k, float ar, float ai,
                              float br, float bi, bool both) {
        const fl2 t = *reinterpret_cast<const fl2 *>(twp + 4 * k);
        const float Ex = ar + br, Oy = ai + bi, ey = ai - bi, ox = ar - br;
        const float P = t.y * ox + t.x * Oy, Q = t.y * Oy - t.x * ox;
        put1<MODE, SPEC>(p, ob + srow<SPEC>(k), k, Ex + P, ey + Q);
        if (both) put1<MODE, SPEC>(p, ob + srow<SPEC>(MC - k), MC - k, Ex - P, Q - ey);
    }
    template <int MODE, bool SPEC>
    VHD static void last_emit2_t(const PkFwdCtx &c, int tid, int it, const C2 (&u)[RL]) {
        const FwdArgs<float> &p = c.a;
        const int lam = tid % TF, w = tid / TF + it * NW2;
        if (w >= G / 2) return;
        char *ob = reinterpret_cast<char *>(c.work) + lam * (SPEC ? 4 : 8);
        const float *twp = c.tw + GEO::TWP;
        if (w != 0) {
            // lo half: group w, bin k = w + G d;  hi half: group G - w, holding Z[Mc - k] at index d
            // (REVG: stored reversed by pass 1) or RL-1-d
#pragma unroll
            for (int d = 0; d < RL; ++d) {
                const int dm = GEO::REVG ? d : RL - 1 - d;
                emit_pair<MODE, SPEC>(p, twp, ob, w + G * d, lo(u[d].x), lo(u[d].y), hi(u[dm].x), hi(u[dm].y), true);
            }
        } else {
            // self-mirrored groups: lo = group 0 (bin G t <-> G (RL - t)), hi = group G/2
            // (bin G/2 + G t <-> G/2 + G (RL-1-t))
#pragma unroll
            for (int t = 0; t < RL / 2; ++t) {
                const int tm = (RL - t) & (RL - 1);
                emit_pair<MODE, SPEC>(p, twp, ob, G * t, lo(u[t].x), lo(u[t].y), lo(u[tm].x), lo(u[tm].y), true);
                emit_pair<MODE, SPEC>(p, twp, ob, G / 2 + G * t, hi(u[t].x), hi(u[t].y), hi(u[RL - 1 - t].x),
                                      hi(u[RL - 1 - t].y), true);
            }
            emit_pair<MODE, SPEC>(p, twp, ob, MC / 2, lo(u[RL / 2].x), lo(u[RL / 2].y), lo(u[RL / 2].x),
                                  lo(u[RL / 2].y), false);
        }
    }
    template <bool SPEC>
    VHD static void last_emit2(const PkFwdCtx &c, int tid, int it, const C2 (&u)[RL]) {
        if (c.a.mode == 3) last_emit2_t<3, SPEC>(c, tid, it, u);
        else last_emit2_t<2, SPEC>(c, tid, it, u);
    }

    // kernel variants: ONES = one-sided layouts (modes 2, 3), SPEC = |X|^2 output.  Keeping the
    // two-sided address arithmetic out of the one-sided kernels matters: the compiler hoists it.
    template <bool ONES, bool SPEC, bool STAGE = false>
    VHD static void last_emit_v(const PkFwdCtx &c, int tid, int it, const C2 (&u)[RL], const C2 *pu) {
        const int mode = c.a.mode;
        if (ONES) {
            if (mode == 3) last_emit_t<3, SPEC, STAGE>(c, tid, it, u, pu);
            else last_emit_t<2, SPEC, STAGE>(c, tid, it, u, pu);
        } else {
            if (mode == 1) last_emit_t<1, SPEC>(c, tid, it, u, pu);
            else last_emit_t<0, SPEC>(c, tid, it, u, pu);
        }
    }
    VHD static void last_emit(const PkFwdCtx &c, int tid, int it, const C2 (&u)[RL], const C2 *pu) {
        const FwdArgs<float> &p = c.a;
        if (p.mode >= 2) {
            if (p.spectrogram) last_emit_v<true, true>(c, tid, it, u, pu);
            else last_emit_v<true, false>(c, tid, it, u, pu);
        } else {
            if (p.spectrogram) last_emit_v<false, true>(c, tid, it, u, pu);
            else last_emit_v<false, false>(c, tid, it, u, pu);
        }
    }
};

// ================================================================== INVERSE (packed)
struct PkInvCtx {
    InvArgs<float> a;
    const float *tw;
    const ITab *itab;        // [MC] ola offsets + dual/M of sample pair m
    float *work;             // planar [MC][2][TF]
    float *ola;              // [TF][ostride] -- ALIASES work
    const float *carry_in;   // [N - H] partial overlap sums of the samples this tile completes
    float *carry_out;        // [N - H] ... and of the samples the next tile completes
    int first;               // first tile of a chunk: no carry-in, its first `halo` slices are a re-computed halo
    int land_shift;          // TMA path: the landed tile starts land_shift (0 / 1) columns before slice 0
    int64_t b, qa;
};

template <int LOGMC, int TF_, int NSUB_>
struct PkInv : PkGeo<LOGMC, TF_, NSUB_> {
    using GEO = PkGeo<LOGMC, TF_, NSUB_>;
    using GEO::MC; using GEO::M; using GEO::TF; using GEO::LN; using GEO::NSUB; using GEO::NT;
    using GEO::NP; using GEO::R1; using GEO::R2; using GEO::R3; using GEO::RL; using GEO::L1;
    using GEO::L2; using GEO::G;
    static constexpr int GPT = (G + NSUB - 1) / NSUB;
    static constexpr int WPT = (L1 / 2 + NSUB / 2 - 1) / (NSUB / 2);    // pass-1 pairs per sub pair

    struct In {
        const char *sb;      // byte pointer of (row 0, column of slice l)
        int64_t rb;
        bool live0, live1;
    };
    VHD static fl2 ldS(const char *q) {              // one complex of S: global memory, read once
#if defined(__CUDA_ARCH__)
        const float2 v = __ldg(reinterpret_cast<const float2 *>(q));
        fl2 r; r.x = v.x; r.y = v.y;
        return r;
#else
        return *reinterpret_cast<const fl2 *>(q);
#endif
    }
    template <bool FULL>
    VHD static Raw get_at(const In &in, const char *q) {
        Raw r;
        if (FULL) {                                  // both slices inside [q0, q1): no predication
            const fl2 a = ldS(q);
            const fl2 b = ldS(q + 8 * LN);
            r.r0 = a.x; r.i0 = a.y; r.r1 = b.x; r.i1 = b.y;
            return r;
        }
        r.r0 = r.i0 = r.r1 = r.i1 = 0.f;
        if (in.live0) { const fl2 v = ldS(q); r.r0 = v.x; r.i0 = v.y; }
        if (in.live1) { const fl2 v = ldS(q + 8 * LN); r.r1 = v.x; r.i1 = v.y; }
        return r;
    }
    // Hermitian one-sided bin k (src/lib.rs:414-487); two-sided layouts: S[k] + conj(S[M-k]).
    // q: byte pointer of row k (one-sided layouts).  Im(DC), Im(Nyquist) are cleared by the caller.
    template <int MODE, bool FULL>
    VHD static Raw load_bin(const InvArgs<float> &p, const In &in, const char *q, int k) {
        Raw v;
        if (MODE >= 2) {
            v = get_at<FULL>(in, q);
            if (MODE == 3 && k >= 1 && k <= MC - 1) {
                v.r0 *= p.rfac2x; v.i0 *= p.rfac2x; v.r1 *= p.rfac2x; v.i1 *= p.rfac2x;
            }
        } else {
            const int sh = MODE == 1 ? MC : 0;                                // ifftshift :516-521
            const Raw a = get_at<FULL>(in, in.sb + ((k + sh) & (M - 1)) * in.rb);
            const Raw bq = get_at<FULL>(in, in.sb + ((M - k + sh) & (M - 1)) * in.rb);
            v.r0 = a.r0 + bq.r0; v.i0 = a.i0 - bq.i0; v.r1 = a.r1 + bq.r1; v.i1 = a.i1 - bq.i1;
        }
        return v;
    }
    // own = X[k_own], rcv = X[Mc - k_own]; t = table entry of k_own.  Returns Z'[k_own].
    // The first adds are scalar and pack the two slices.
    VHD static C2 merge1(fl4 t, Raw own, Raw rcv) {
        const f2 Ex = mk2(own.r0 + rcv.r0, own.r1 + rcv.r1), Dy = mk2(own.i0 + rcv.i0, own.i1 + rcv.i1);
        const f2 ey = mk2(own.i0 - rcv.i0, own.i1 - rcv.i1), dx = mk2(own.r0 - rcv.r0, own.r1 - rcv.r1);
        C2 z;
        z.x = vfma(bc(t.z), Dy, vfma(bc(t.y), dx, Ex));      // Ex - wr Dy + wi dx
        z.y = vfma(bc(t.x), dx, vfma(bc(t.y), Dy, ey));      // ey + wr dx + wi Dy
        return z;
    }

    // ---- pass 1, part 1: load this thread's group of bins (raw)
    //      sub = 2w + h:  w = 0 -> groups (0, L1/2);  w >= 1 -> groups (w, L1 - w)
    VHD static int group_of(int w, int h) { return h == 0 ? w : (w == 0 ? L1 / 2 : L1 - w); }
    VHD static In make_in(const PkInvCtx &c, int l) {
        const InvArgs<float> &p = c.a;
        const int64_t q0 = c.qa + l, q1 = q0 + LN;
        const int64_t col = q0 - p.p_min;
        In in;
        in.live0 = q0 >= p.q0 && q0 < p.q1 && col >= 0 && col < p.P;
        in.live1 = q1 >= p.q0 && q1 < p.q1 && col + LN >= 0 && col + LN < p.P;
        in.rb = p.ldp * 8;
        in.sb = reinterpret_cast<const char *>(p.S) + ((c.b * p.F) * p.ldp + col) * 8;
        return in;
    }
    // Half 1 (h = 1) loads its group BACKWARDS: x[a] = X[jp + L1 (R1-1-a)], so that x[a] of the two
    // halves are each other's mirror bins and no register selects are needed.
    template <int MODE>
    VHD static void p1_load_t(const PkInvCtx &c, int tid, int it, Raw (&x)[R1], Raw &xn) {
        const int l = tid % LN, sub = tid / LN;
        const int w = (sub >> 1) + it * (NSUB / 2);
        if (w >= L1 / 2) return;
        const In in = make_in(c, l);
        if (in.live0 && in.live1) p1_load_f<MODE, true>(c, tid, it, in, x, xn);
        else p1_load_f<MODE, false>(c, tid, it, in, x, xn);
    }
    template <int MODE, bool FULL>
    VHD static void p1_load_f(const PkInvCtx &c, int tid, int it, const In &in, Raw (&x)[R1], Raw &xn) {
        const int sub = tid / LN;
        const int w = (sub >> 1) + it * (NSUB / 2), h = sub & 1;
        const int j = group_of(w, h);
        const bool rev = h != 0 && w != 0;
        int k = rev ? j + L1 * (R1 - 1) : j;
        const int kstep = rev ? -L1 : L1;
        const char *q = in.sb + k * in.rb;
        const int64_t qstep = kstep * in.rb;
#pragma unroll
        for (int a = 0; a < R1; ++a) {
            x[a] = load_bin<MODE, FULL>(c.a, in, q, k);
            q += qstep;
            keep(q);
            k += kstep;
        }
        xn = x[0];
        if (w == 0 && h == 0) {                                           // DC pairs with Nyquist;
            xn = load_bin<MODE, FULL>(c.a, in, in.sb + MC * in.rb, MC);   // irfft ignores their Im
            x[0].i0 = x[0].i1 = 0.f;
            xn.i0 = xn.i1 = 0.f;
        }
    }
    template <bool ONES>
    VHD static void p1_load_v(const PkInvCtx &c, int tid, int it, Raw (&x)[R1], Raw &xn) {
        const int mode = c.a.mode;
        if (ONES) {
            if (mode == 3) p1_load_t<3>(c, tid, it, x, xn);
            else p1_load_t<2>(c, tid, it, x, xn);
        } else {
            if (mode == 1) p1_load_t<1>(c, tid, it, x, xn);
            else p1_load_t<0>(c, tid, it, x, xn);
        }
    }
    VHD static void p1_load(const PkInvCtx &c, int tid, int it, Raw (&x)[R1], Raw &xn) {
        if (c.a.mode >= 2) p1_load_v<true>(c, tid, it, x, xn);
        else p1_load_v<false>(c, tid, it, x, xn);
    }
    // ---- pass 1, part 2: exchange, real-IFFT merge, IDFT_R1, twiddle, store
    //      px = the partner thread's x[] (host emulation only)
    // Half 1 runs the SAME IDFT on its reversed array v~[a] = z[R1-1-a]:
    //   IDFT(v~)[k] = w^(-k) Y[(-k) mod R1],  w = exp(+2 pi i / R1),
    // so its local output k is the true output k' = (R1-k) mod R1 times w^k; the factor is folded
    // into the tw1x table and the store goes to row block k' (per-thread pointer stride).
    VHD static void p1_rest(const PkInvCtx &c, int tid, int it, const Raw (&x)[R1], const Raw &xn, const Raw *px) {
        const int l = tid % LN, sub = tid / LN;
        const int w = (sub >> 1) + it * (NSUB / 2), h = sub & 1;
        if (w >= L1 / 2) return;
        const float *twp = c.tw + GEO::TWP;
        const int j = group_of(w, h);
        const unsigned mask = pair_mask<LN>(tid);
        const bool rev = h != 0 && w != 0;
        C2 v[R1];
        if (w != 0) {
            const float *tq = twp + 4 * (rev ? j + L1 * (R1 - 1) : j);
            const int tstep = rev ? -4 * L1 : 4 * L1;
#pragma unroll
            for (int a = 0; a < R1; ++a) {
#ifdef VEILS_ABLATE
                const Raw rcv = (c.a.dbg & 2) ? x[a] : xchg_raw<LN>(x[a], px ? px + a : nullptr, mask);
#else
                const Raw rcv = xchg_raw<LN>(x[a], px ? px + a : nullptr, mask);
#endif
                v[a] = merge1(GEO::tw4(tq), x[a], rcv);
                tq += tstep;
            }
        } else {
            // self-mirrored groups: half 0 holds group 0 (x[a] <-> x[(R1-a) mod R1], DC <-> Nyquist), half 1
            // group L1/2 (x[a] <-> x[R1-1-a]).  One instruction stream: the mirror operand and the slot of
            // the mirrored result are register selects (cheaper than running the two paths in turn while
            // the seven other warps wait at the barrier).
            C2 m[R1 / 2];
#pragma unroll
            for (int t = 0; t < R1 / 2; ++t) {
                const Raw b = h == 0 ? (t == 0 ? xn : x[R1 - t]) : x[R1 - 1 - t];
                const int k = h == 0 ? L1 * t : j + L1 * t;
                v[t] = merge1(GEO::tw4(twp + 4 * k), x[t], b);
                m[t] = merge1(GEO::tw4(twp + 4 * (MC - k)), b, x[t]);             // Z'[Mc - k]
            }
            if (h == 0) {
                // k = 0: Z'[0] = (X0 + XN) + i (X0 - XN) with real X0, XN (imag parts cleared on load)
                v[0].x = mk2(x[0].r0 + xn.r0, x[0].r1 + xn.r1);
                v[0].y = mk2(x[0].r0 - xn.r0, x[0].r1 - xn.r1);
                v[R1 / 2] = merge1(GEO::tw4(twp + 4 * (MC / 2)), x[R1 / 2], x[R1 / 2]);
            }
#pragma unroll
            for (int t = 0; t < R1 / 2; ++t) {
                // half 0: mirror of bin L1 t is element R1 - t (t >= 1); half 1: element R1 - 1 - t
                if (t >= 1) v[R1 - t] = h == 0 ? m[t] : (t == 1 ? m[0] : m[t - 1]);
            }
            if (h != 0) v[R1 / 2] = m[R1 / 2 - 1];
        }
        Dft<R1, true, f2>::run(v);
        float *wj = c.work + (j * 2) * TF + 2 * l;
        const float *twj = c.tw + (rev ? GEO::TW1X : GEO::TW1) + 4 * j * R1;
        wstore(wj, TF, v[0]);
        float *wq = rev ? wj + (L1 * R1 * 2) * TF : wj;
        const int wstep = rev ? -(L1 * 2) * TF : (L1 * 2) * TF;
#pragma unroll
        for (int cc = 1; cc < R1; ++cc) {
            wq += wstep;
            const fl4 t = GEO::tw4(twj + 4 * cc);
            wstore(wq, TF, cmulwc(v[cc], bc(t.x), bc(t.y), bc(t.w)));
        }
    }

    // ================= paired pass 1 (one-sided layouts): one slice per lane, no partner exchange =====
    // Thread (w, lam): slice lam of the tile, mirrored group pair (j0, j1) = (w, L1 - w), w = 0:
    // (0, L1/2).  It loads the 2 R1 bins of both groups for its slice -- a warp reads one 256-byte
    // row run per instruction -- merges each mirrored bin pair (k, Mc - k) ONCE (src/lib.rs:428-453
    // builds the full Hermitian spectrum instead), and runs one packed IDFT over both groups: group
    // j0 rides in the low halves of the f32x2 registers, group j1 in the high halves.
    static constexpr int NW2 = NT / TF;
    static constexpr int WPT2 = (L1 / 2 + NW2 - 1) / NW2;
    static constexpr int LTF = TF + 2;                 // slices per landed row (TMA path)
    static constexpr size_t LAND_BYTES = (size_t)(MC + 1) * LTF * 8;
    VHD static int wpos(int lam) { return lam < LN ? 2 * lam : 2 * (lam - LN) + 1; }
    struct Cp { float r, i; };
    // SMEM: q points into the tile that the TMA engine has landed in shared memory (pk_inv_kernel)
    template <int MODE, bool LIVE, bool SMEM = false>
    VHD static Cp ldbin(const InvArgs<float> &p, const char *q, int k) {
        Cp v; v.r = 0.f; v.i = 0.f;
        if (LIVE) {
#ifdef VEILS_ABLATE
            if (p.dbg & 1) { v.r = 1.f + k; v.i = 0.5f; return v; }      // ablation: no global loads
#endif
            const fl2 t = SMEM ? *reinterpret_cast<const fl2 *>(q) : ldS(q);
            v.r = t.x; v.i = t.y;
            if (MODE == 3 && k >= 1 && k <= MC - 1) { v.r *= p.rfac2x; v.i *= p.rfac2x; }
        }
        return v;
    }
    // Z'[k] -> (zr, zi), Z'[Mc-k] -> (mr, mi) from A = X[k], B = X[Mc-k], (wr, wi) = W_M^k
    VHD static void merge_pair(fl2 t, Cp A, Cp B, float &zr, float &zi, float &mr, float &mi) {
        const float Ex = A.r + B.r, Dy = A.i + B.i, ey = A.i - B.i, dx = A.r - B.r;
        const float P = t.x * Dy - t.y * dx, Q = t.x * dx + t.y * Dy;
        zr = Ex - P; zi = ey + Q;
        mr = Ex + P; mi = Q - ey;
    }
    template <int MODE, bool LIVE>
    VHD static void p1p_t(const PkInvCtx &c, int tid, int it, const char *sb) {
        const int w = tid / TF + it * NW2;
        if (w >= L1 / 2) return;
        Cp x0[R1], x1[R1], xn;
        p1p_load<MODE, LIVE, false>(c, tid, it, sb, x0, x1, xn);
        p1p_rest(c, tid, it, x0, x1, xn);
    }
    // loads of the 2 R1 (+1) bins of group pair w for this thread's slice.  SMEM = false: from S
    // (sb = byte pointer of (row 0, this slice's column)); SMEM = true: from the tile the TMA engine
    // has landed over the work buffer -- row j R1 + a holds bin j + L1 a, [TF slices] of 8 bytes.
    template <int MODE, bool LIVE, bool SMEM>
    VHD static void p1p_load(const PkInvCtx &c, int tid, int it, const char *sb, Cp (&x0)[R1], Cp (&x1)[R1], Cp &xn) {
        const InvArgs<float> &p = c.a;
        const int lam = tid % TF, w = tid / TF + it * NW2;
        xn.r = xn.i = 0.f;
        if (w >= L1 / 2) return;
        const int j0 = w, j1 = w == 0 ? L1 / 2 : L1 - w;
        if (SMEM) {
            // landed rows are LTF = TF + 2 slices wide (the load starts at an even column)
            const char *base = reinterpret_cast<const char *>(c.work) + (lam + c.land_shift) * 8;
            const char *q0 = base + (size_t)(j0 * R1) * (LTF * 8);
            const char *q1 = base + (size_t)(j1 * R1) * (LTF * 8);
#pragma unroll
            for (int a = 0; a < R1; ++a) {
                x0[a] = ldbin<MODE, LIVE, true>(p, q0 + a * (LTF * 8), j0 + L1 * a);
                x1[a] = ldbin<MODE, LIVE, true>(p, q1 + a * (LTF * 8), j1 + L1 * a);
            }
            if (w == 0) xn = ldbin<MODE, LIVE, true>(p, base + (size_t)MC * (LTF * 8), MC);
            return;
        }
        const int64_t rb = p.ldp * 8;
        const char *q0 = sb + j0 * rb, *q1 = sb + j1 * rb;
        const int64_t st = L1 * rb;
#pragma unroll
        for (int a = 0; a < R1; ++a) {
            x0[a] = ldbin<MODE, LIVE>(p, q0, j0 + L1 * a);
            x1[a] = ldbin<MODE, LIVE>(p, q1, j1 + L1 * a);
            q0 += st;
            keep(q0);
            q1 += st;
            keep(q1);
        }
        if (w == 0) xn = ldbin<MODE, LIVE>(p, sb + MC * rb, MC);
    }
    VHD static void p1p_rest(const PkInvCtx &c, int tid, int it, const Cp (&x0)[R1], const Cp (&x1)[R1], const Cp &xn) {
        const int lam = tid % TF, w = tid / TF + it * NW2;
        if (w >= L1 / 2) return;
        const int j0 = w, j1 = w == 0 ? L1 / 2 : L1 - w;
        const float *twp = c.tw + GEO::TWP;
        float lr[R1], li[R1], hr[R1], hi_[R1];
        if (w != 0) {
            // bin k = j0 + L1 a  <->  Mc - k = j1 + L1 (R1-1-a)
#pragma unroll
            for (int a = 0; a < R1; ++a)
                merge_pair(GEO::tw2(twp + 4 * (j0 + L1 * a)), x0[a], x1[R1 - 1 - a], lr[a], li[a], hr[R1 - 1 - a], hi_[R1 - 1 - a]);
        } else {
            // group 0: bin L1 a <-> L1 (R1 - a), DC <-> Nyquist (their imaginary parts are ignored,
            // irfft semantics); group L1/2: bin L1/2 + L1 a <-> L1/2 + L1 (R1-1-a)
            lr[0] = x0[0].r + xn.r;
            li[0] = x0[0].r - xn.r;
#pragma unroll
            for (int a = 1; a < R1 / 2; ++a)
                merge_pair(GEO::tw2(twp + 4 * (L1 * a)), x0[a], x0[R1 - a], lr[a], li[a], lr[R1 - a], li[R1 - a]);
            {
                float mr, mi;
                merge_pair(GEO::tw2(twp + 4 * (MC / 2)), x0[R1 / 2], x0[R1 / 2], lr[R1 / 2], li[R1 / 2], mr, mi);
            }
#pragma unroll
            for (int a = 0; a < R1 / 2; ++a)
                merge_pair(GEO::tw2(twp + 4 * (j1 + L1 * a)), x1[a], x1[R1 - 1 - a], hr[a], hi_[a], hr[R1 - 1 - a], hi_[R1 - 1 - a]);
        }
        C2 v[R1];
#pragma unroll
        for (int a = 0; a < R1; ++a) {
            v[a].x = mk2(lr[a], hr[a]);
            v[a].y = mk2(li[a], hi_[a]);
        }
        Dft<R1, true, f2>::run(v);
        // twiddle W_Mc^(-j c) (each half its own group) and the stores: group j output c goes to
        // work row j + L1 c, planar [re: TF][im: TF], slice lam at position wpos(lam)
        float *w0 = c.work + (j0 * 2) * TF + wpos(lam), *w1 = c.work + (j1 * 2) * TF + wpos(lam);
        const float *tq = c.tw + GEO::TW1P + 4 * w * R1;
        w0[0] = lo(v[0].x); w0[TF] = lo(v[0].y);
        w1[0] = hi(v[0].x); w1[TF] = hi(v[0].y);
#pragma unroll
        for (int cc = 1; cc < R1; ++cc) {
            const fl4 t = GEO::tw4(tq + 4 * cc);
            const f2 wr = mk2(t.x, t.y), wi = mk2(t.z, t.w);
            const C2 z = cmulwc(v[cc], wr, wi, bc(0.f) - wi);
            w0[(L1 * cc * 2) * TF] = lo(z.x); w0[(L1 * cc * 2 + 1) * TF] = lo(z.y);
            w1[(L1 * cc * 2) * TF] = hi(z.x); w1[(L1 * cc * 2 + 1) * TF] = hi(z.y);
        }
    }
    // slice of this thread inside [q0, q1) and inside S?
    VHD static bool p1p_live(const PkInvCtx &c, int tid) {
        const InvArgs<float> &p = c.a;
        const int64_t q = c.qa + tid % TF, col = q - p.p_min;
        return q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;
    }
    // from the landed tile (TMA path): loads only; the caller synchronises the CTA before p1p_rest
    VHD static void p1p_load_smem(const PkInvCtx &c, int tid, int it, Cp (&x0)[R1], Cp (&x1)[R1], Cp &xn) {
        const bool live = p1p_live(c, tid);
        if (c.a.mode == 3) {
            if (live) p1p_load<3, true, true>(c, tid, it, nullptr, x0, x1, xn);
            else p1p_load<3, false, true>(c, tid, it, nullptr, x0, x1, xn);
        } else {
            if (live) p1p_load<2, true, true>(c, tid, it, nullptr, x0, x1, xn);
            else p1p_load<2, false, true>(c, tid, it, nullptr, x0, x1, xn);
        }
    }
    VHD static void p1p(const PkInvCtx &c, int tid, int it) {
        const InvArgs<float> &p = c.a;
        const int lam = tid % TF;
        const int64_t q = c.qa + lam, col = q - p.p_min;
        const bool live = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;
        const char *sb = reinterpret_cast<const char *>(p.S) + ((c.b * p.F) * p.ldp + col) * 8;
        if (p.mode == 3) {
            if (live) p1p_t<3, true>(c, tid, it, sb);
            else p1p_t<3, false>(c, tid, it, sb);
        } else {
            if (live) p1p_t<2, true>(c, tid, it, sb);
            else p1p_t<2, false>(c, tid, it, sb);
        }
    }

    VHD static void pass_mid(const PkInvCtx &c, int tid) {
        if (NP != 3) return;
        const int l = tid % LN, sub = tid / LN;
        for (int u = sub; u < R1 * L2; u += NSUB) {
            const int cb = u / L2, j2 = u % L2;
            float *wb = c.work + ((cb * L1 + j2) * 2) * TF + 2 * l;
            const float *twj = c.tw + GEO::TW2 + 4 * j2 * R2;
            C2 v[R2];
#pragma unroll
            for (int a = 0; a < R2; ++a) v[a] = wload(wb + (L2 * a * 2) * TF, TF);
            Dft<R2, true, f2>::run(v);
            wstore(wb, TF, v[0]);
#pragma unroll
            for (int cc = 1; cc < R2; ++cc) {
                const fl4 t = GEO::tw4(twj + 4 * cc);
                wstore(wb + (L2 * cc * 2) * TF, TF, cmulwc(v[cc], bc(t.x), bc(t.y), bc(t.w)));
            }
        }
    }

    VHD static void last_load(const PkInvCtx &c, int tid, C2 (&r)[GPT][RL]) {
        const int l = tid % LN, sub = tid / LN;
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = sub + NSUB * i;
            if (g < G) {
                const float *w = c.work + (GEO::goff(g) * 2) * TF + 2 * l;
#pragma unroll
                for (int d = 0; d < RL; ++d) r[i][d] = wload(w + (d * 2) * TF, TF);
                Dft<RL, true, f2>::run(r[i]);
            }
        }
    }
    // roll back (src/lib.rs:499-500), truncate to N, dual window -- all folded into the plan table
    // itab[m] = (ola offset of y[2m], of y[2m+1], dual/M of both); dropped samples go to a dump slot
    VHD static void last_store(const PkInvCtx &c, int tid, const C2 (&r)[GPT][RL]) {
        if (c.a.pair) last_store_t<true>(c, tid, r);
        else last_store_t<false>(c, tid, r);
    }
    template <bool PAIR>
    VHD static void last_store_t(const PkInvCtx &c, int tid, const C2 (&r)[GPT][RL]) {
        const InvArgs<float> &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        float *row0 = c.ola + (size_t)l * p.ostride, *row1 = row0 + (size_t)LN * p.ostride;
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = sub + NSUB * i;
            if (g < G) {
                const ITab *tg = c.itab + g;
#pragma unroll
                for (int d = 0; d < RL; ++d) {
                    const ITab e = GEO::it4(tg + G * d);                     // sample pair m = g + G d
                    if (PAIR) {                                              // y[2m], y[2m+1] adjacent, 8-byte aligned
                        fl2 a, b;
                        a.x = lo(r[i][d].x) * e.d0; a.y = lo(r[i][d].y) * e.d1;
                        b.x = hi(r[i][d].x) * e.d0; b.y = hi(r[i][d].y) * e.d1;
                        *reinterpret_cast<fl2 *>(row0 + e.o0) = a;
                        *reinterpret_cast<fl2 *>(row1 + e.o0) = b;
                    } else {
                        row0[e.o0] = lo(r[i][d].x) * e.d0;
                        row1[e.o0] = hi(r[i][d].x) * e.d0;
                        row0[e.o1] = lo(r[i][d].y) * e.d1;
                        row1[e.o1] = hi(r[i][d].y) * e.d1;
                    }
                }
            }
        }
    }

    // ================= accumulator overlap-add (75 % overlap, m_num = mfft = 4 hop, default roll) ======
    // Instead of TF rows of m_num samples (66 KB) the tile's output lives in ONE padded accumulator
    // acc[u + 2 (u / H)], u in [0, (TF + 3) H): sample u of the tile, slice sl contributes its quarter
    // c = 0..3 to u in [(sl + c) H, (sl + c + 1) H).  Four passes, OLDEST slice first (c = 3, 2, 1, 0),
    // each a conflict-free store / read-modify-write of one quarter of every slice: within a pass every
    // accumulator cell receives exactly one contribution, across passes the slices arrive in increasing
    // q -- the reference's summation order (src/lib.rs:853, 876-910), deterministic, no atomics.  The
    // head [0, 3H) carries the previous tile's tail.  The work buffer is free as soon as the last
    // pass has loaded its inputs, so the next tile's S boxes land during these passes (pk_inv_kernel).
    VHD static float mul_rn(float x, float y) {
#if defined(__CUDA_ARCH__)
        return __fmul_rn(x, y);
#else
        return x * y;
#endif
    }
    static constexpr int HC = MC / 2;                  // hop of this path (m_num = 2 Mc = 4 hop)
    static constexpr int HP = HC + 2;                  // padded quarter: lanes (slices) hit distinct banks
    static constexpr int ACC_FLOATS = (TF + 3) * HP;
    VHD static int apos(int u) { return u + 2 * (u / HC); }
    template <int Q>
    VHD static void ola_acc(const PkInvCtx &c, int tid, const C2 (&r)[GPT][RL], float *acc) {
        const int l = tid % LN, sub = tid / LN;
        float *a0 = acc + l * HP + 2 * Q, *a1 = a0 + LN * HP;
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = sub + NSUB * i;
            if (g >= G) continue;
            const ITab *tg = c.itab + g;
#pragma unroll
            for (int dd = 0; dd < RL / 4; ++dd) {
                constexpr int RQ = RL / 4;
                const int d = (Q * RQ + dd + RL / 2) % RL;       // sample pairs g + G d of quarter Q (roll = m_num / 2)
                // this path's geometry (m_num = mfft, roll = m_num / 2): the pair lands at offset 2g + C(d) of
                // its slice, C compile-time -- only the dual-window pair is read from the table (8 bytes)
                const int o0 = 2 * g + (2 * G * d + M / 2) % M;
                const fl2 dw = GEO::dw2(tg + G * d);
                // products rounded on their own (no contraction into the adds below): the sums are then
                // bit-identical to the row-based overlap-add of the other kernels, whatever path a call takes
                fl2 a, b;
                a.x = mul_rn(lo(r[i][d].x), dw.x); a.y = mul_rn(lo(r[i][d].y), dw.y);
                b.x = mul_rn(hi(r[i][d].x), dw.x); b.y = mul_rn(hi(r[i][d].y), dw.y);
                fl2 *p0 = reinterpret_cast<fl2 *>(a0 + o0), *p1 = reinterpret_cast<fl2 *>(a1 + o0);
                if (Q == 3) {                                     // oldest contribution of every cell it touches
                    *p0 = a;
                    *p1 = b;
                } else {
                    fl2 t0 = *p0, t1 = *p1;
                    t0.x += a.x; t0.y += a.y;
                    t1.x += b.x; t1.y += b.y;
                    *p0 = t0;
                    *p1 = t1;
                }
            }
        }
    }
    VHD static void ola_zero_head(int tid, float *acc) {
        for (int i = tid; i < 3 * HP; i += NT) acc[i] = 0.f;
    }
    // completed samples [0, TF H) -> global (clipped to [k0, k1) and, for the first tile of a chunk, to the
    // samples it owns); returns this thread's pair of the tail (handed to the next tile as its head)
    VHD static fl2 ola_out(const PkInvCtx &c, int tid, const float *acc) {
        const InvArgs<float> &p = c.a;
        constexpr int TFH = TF * HC;
        const int64_t base = c.qa * (int64_t)HC - p.mid;
        float *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        int64_t ka = base + (c.first ? 3 * HC : 0), kb = base + TFH;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const bool inner = ua <= 0 && ub >= TFH;
        for (int u = 2 * tid; u < TFH; u += 2 * NT) {
            const fl2 o = *reinterpret_cast<const fl2 *>(acc + apos(u));
            if (inner || (u >= ua && u + 1 < ub)) *reinterpret_cast<fl2 *>(xo + u) = o;
            else {
                if (u >= ua && u < ub) xo[u] = o.x;
                if (u + 1 >= ua && u + 1 < ub) xo[u + 1] = o.y;
            }
        }
        fl2 tail; tail.x = 0.f; tail.y = 0.f;
        return tail;
    }
    // the tail [TF H, TF H + 3H) becomes the next tile's head: each thread carries TPT pairs in registers
    static constexpr int TPT = (3 * HC / 2 + NT - 1) / NT;
    struct Tail { fl2 v[TPT]; };
    VHD static Tail ola_take_tail(int tid, const float *acc) {
        Tail t;
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            const int u = 2 * (tid + i * NT);
            t.v[i].x = t.v[i].y = 0.f;
            if (u < 3 * HC) t.v[i] = *reinterpret_cast<const fl2 *>(acc + apos(TF * HC + u));
        }
        return t;
    }
    VHD static void ola_keep_tail(int tid, float *acc, const Tail &t) {
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            const int u = 2 * (tid + i * NT);
            if (u < 3 * HC) *reinterpret_cast<fl2 *>(acc + apos(u)) = t.v[i];
        }
    }

    // Streaming overlap-add (src/lib.rs:853-911), increasing q:  local sample u = k + mid - qa*H.
    //   u in [0, TF*H)           : completed by this tile -> global memory.  Contributions of earlier
    //                              slices arrive in carry_in (u < N - H); the first tile of a chunk has
    //                              no carry and only completes u >= halo*H.
    //   u in [TF*H, TF*H + N - H): partial sums over this tile's slices -> carry_out.
    // The summation order (earlier tiles' slices first, then this tile's in increasing q) is the
    // reference's; everything is owned by this CTA: deterministic, no atomics.
    VHD static float ola_sum(const PkInvCtx &c, int u, float acc) {
        const InvArgs<float> &p = c.a;
        const int H = p.H, N = p.N, halo = p.halo;
        const int rq = (int)mulhi_u32((unsigned)u, p.magic_h);          // u / H: latest slice
        int r = rq - halo;
        int jj = u - r * H;
        for (int h = 0; h <= halo; ++h, ++r, jj -= H)
            if (r >= 0 && r < TF && jj < N) acc += c.ola[r * p.ostride + jj];
        return acc;
    }
    VHD static void gather(const PkInvCtx &c, int tid) {
        const InvArgs<float> &p = c.a;
        const int H = p.H, N = p.N, halo = p.halo;
        const int TFH = TF * H, NH = N - H, hH = halo * H;
        const int64_t base = c.qa * (int64_t)H - p.mid;                 // k = base + u
        float *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        // output range of this tile, clipped to [k0, k1)
        int64_t ka = base + (c.first ? hH : 0), kb = base + TFH;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const int u_hot = hH > NH ? hH : NH;                            // from here on: no carry, all slices local
        // (1) the bulk: u in [u_hot, TF*H)
        if (halo == 3 && N == 4 * H && p.pair && (H & 1) == 0 && ((((size_t)xo) & 7) == 0)) {
            const int step = p.ostride - H;
            const bool inner = ua <= u_hot && ub >= TFH;
            const int cnt = TFH - hH;
            const int rs = (int)mulhi_u32((unsigned)(2 * NT), p.magic_h);    // (2 NT) / H
            if (inner && rs * H == 2 * NT) {
                // the thread's column inside a slice never changes and its row advances by rs per
                // iteration: pure pointer increments, no index arithmetic in the loop
                const int e0 = 2 * tid;
                const int r0 = (int)mulhi_u32((unsigned)e0, p.magic_h);
                const float *pa = c.ola + r0 * p.ostride + (e0 - r0 * H) + hH;
                const float *pb = pa + step, *pc = pb + step, *pd = pc + step;
                const int adv = rs * p.ostride;
                float *po = xo + e0 + hH;
                for (int e = e0; e < cnt; e += 2 * NT) {
                    const fl2 a0 = *reinterpret_cast<const fl2 *>(pa);
                    const fl2 a1 = *reinterpret_cast<const fl2 *>(pb);
                    const fl2 a2 = *reinterpret_cast<const fl2 *>(pc);
                    const fl2 a3 = *reinterpret_cast<const fl2 *>(pd);
                    fl2 o;
                    o.x = ((a0.x + a1.x) + a2.x) + a3.x;
                    o.y = ((a0.y + a1.y) + a2.y) + a3.y;
                    *reinterpret_cast<fl2 *>(po) = o;
                    pa += adv; pb += adv; pc += adv; pd += adv;
                    po += 2 * NT;
                }
            } else
            for (int e = 2 * tid; e < cnt; e += 2 * NT) {               // two samples per thread
                const int u = e + hH;
                const int r = (int)mulhi_u32((unsigned)e, p.magic_h);
                const int idx = r * p.ostride + (e - r * H) + hH;
                const fl2 a0 = *reinterpret_cast<const fl2 *>(c.ola + idx);
                const fl2 a1 = *reinterpret_cast<const fl2 *>(c.ola + idx + step);
                const fl2 a2 = *reinterpret_cast<const fl2 *>(c.ola + idx + 2 * step);
                const fl2 a3 = *reinterpret_cast<const fl2 *>(c.ola + idx + 3 * step);
                fl2 o;
                o.x = ((a0.x + a1.x) + a2.x) + a3.x;
                o.y = ((a0.y + a1.y) + a2.y) + a3.y;
                if (inner || (u >= ua && u + 1 < ub)) *reinterpret_cast<fl2 *>(xo + u) = o;
                else {
                    if (u >= ua && u < ub) xo[u] = o.x;
                    if (u + 1 >= ua && u + 1 < ub) xo[u + 1] = o.y;
                }
            }
        } else {
            for (int u = u_hot + tid; u < TFH; u += NT)
                if (u >= ua && u < ub) xo[u] = ola_sum(c, u, 0.f);
        }
        if (halo == 3 && N == 4 * H && p.pair && (H & 1) == 0 && ((((size_t)xo) & 7) == 0)) {
            // (2) head, (3) tail for the 75 % overlap case: block b of H samples sums b+1 (head) or 3-b
            // (tail) local slices; two samples per thread, the same 8-byte accesses as the bulk
            const int step = p.ostride - H, hp = H >> 1;
            for (int e = tid; e < 3 * hp; e += NT) {
                const int b = e >= 2 * hp ? 2 : (e >= hp ? 1 : 0);
                const int c0 = 2 * (e - b * hp);
                if (!c.first) {                                      // head: u = b H + c0
                    const int u = b * H + c0;
                    fl2 o = *reinterpret_cast<const fl2 *>(c.carry_in + u);
                    int idx = u;                                     // row 0, offset u; next row: + step
                    for (int i = 0; i <= b; ++i, idx += step) {
                        const fl2 a = *reinterpret_cast<const fl2 *>(c.ola + idx);
                        o.x += a.x;
                        o.y += a.y;
                    }
                    if (u >= ua && u + 1 < ub) *reinterpret_cast<fl2 *>(xo + u) = o;
                    else {
                        if (u >= ua && u < ub) xo[u] = o.x;
                        if (u + 1 >= ua && u + 1 < ub) xo[u + 1] = o.y;
                    }
                }
                {                                                    // tail: v = b H + c0, rows TF-3+b .. TF-1
                    const int v = b * H + c0;
                    int idx = (TF - 3 + b) * p.ostride + 3 * H + c0;       // earliest slice: its last quarter
                    fl2 o = *reinterpret_cast<const fl2 *>(c.ola + idx);
                    idx += step;
                    for (int i = 1; i < 3 - b; ++i, idx += step) {
                        const fl2 a = *reinterpret_cast<const fl2 *>(c.ola + idx);
                        o.x += a.x;
                        o.y += a.y;
                    }
                    *reinterpret_cast<fl2 *>(c.carry_out + v) = o;
                }
            }
            return;
        }
        // (2) the head completed with the carry of the previous tile: u in [0, u_hot)
        if (!c.first) {
            for (int u = tid; u < u_hot; u += NT)
                if (u >= ua && u < ub) xo[u] = ola_sum(c, u, u < NH ? c.carry_in[u] : 0.f);
        }
        // (3) the tail handed to the next tile
        for (int v = tid; v < NH; v += NT) c.carry_out[v] = ola_sum(c, TFH + v, 0.f);
    }
};

}  // namespace p2
}  // namespace veils
