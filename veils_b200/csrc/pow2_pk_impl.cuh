// pow2_pk_impl.cuh -- f32 "packed slice-lane" kernels: the pow2 tile family with TWO consecutive
// time slices per lane, the same component of both kept in one f32x2 register pair, so every
// butterfly add / mul / fma is a single FADD2 / FMUL2 / FFMA2 (see f32x2.cuh).
//
// Versus the scalar family (pow2_tile_impl.cuh) this halves the issued instructions per slice:
// arithmetic is packed, shared-memory traffic moves 8 bytes (two slices) per lane and access, the
// plan tables are read once per two slices, and S is stored / loaded as 16-byte {re,im,re,im}
// vectors of two adjacent slices (256-byte runs per row).  Layout differences:
//   * thread t = (sub, l): l = t % (TF/2) owns slices 2l, 2l+1 of the tile;
//   * work buffer is planar: work[(e*2 + comp)*TF + slice]  (comp 0 = re, 1 = im);
//   * twiddle tables hold (wr, wi, -wr, -wi) so that no packed negation is ever needed
//     (FFMA2 has no negate modifier that ptxas folds for us).
// The AoS <-> packed transposition is free: it is done by the first scalar op after a load
// (window multiply / Hermitian merge add) and the last scalar op before a store.
// Phase functions compile for the device and for the host emulation (tests/emu).
#pragma once
#include "pow2_tile_impl.cuh"

namespace veils {
namespace p2 {

struct alignas(8) fl2 { float x, y; };
struct alignas(16) fl4 { float x, y, z, w; };

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
    // twiddle table: entries of 4 floats (wr, wi, -wr, -wi): [tw1: MC][tw2: L1][twp: MC + 1]
    static constexpr int TW1 = 0, TW2 = 4 * MC, TWP = 4 * (MC + L1), TWN = 4 * (2 * MC + L1 + 1);
    VHD static int goff(int g) { return (NP == 2) ? g * L1 : (g % R1) * L1 + (g / R1) * L2; }
    static constexpr bool STAB = LOGMC <= 8;
    static constexpr bool DBUF = LOGMC <= 8;
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
    VHD static fl2 t2(const float *q) {
#if defined(__CUDA_ARCH__)
        if (!STAB) {
            const float2 v = __ldg(reinterpret_cast<const float2 *>(q));
            fl2 r; r.x = v.x; r.y = v.y;
            return r;
        }
#endif
        return *reinterpret_cast<const fl2 *>(q);
    }
    VHD static float t1(const float *q) {
#if defined(__CUDA_ARCH__)
        if (!STAB) return __ldg(q);
#endif
        return *q;
    }
};

using C2 = cx<f2>;
VHD C2 c2zero() { C2 r; r.x = bc(0.f); r.y = bc(0.f); return r; }
// planar work-buffer access for the two slices of this lane
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

// ================================================================== FORWARD (packed)
struct PkFwdCtx {
    FwdArgs<float> a;
    const float *tw, *win;
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

    // staging of the signal tile: identical layout to the scalar family (paired cp.async)
    VHD static void fill(const PkFwdCtx &c, int tid) {
        const FwdArgs<float> &p = c.a;
        const int H = p.H;
        const int span = (TF - 1) * H + p.N;
        const int hs = H + p.lay.padw;
        const int64_t s0 = (p.p0 + c.pt * TF) * (int64_t)H + p.k_offset - p.mid;
        const float *xb = p.x + c.b * p.x_pitch;
        if ((((c.b * p.x_pitch + s0) & 1) == 0) && ((p.n & 1) == 0) && ((((size_t)p.x) & 7) == 0)) {
            const int np = span >> 1, hp = H >> 1;
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

    // ---- pass 1
    template <bool PADDED>
    VHD static void pass1_t(const PkFwdCtx &c, int tid) {
        const FwdArgs<float> &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        const int hs = p.H + p.lay.padw;
        const float *xf0 = c.xin + (2 * l) * hs, *xf1 = xf0 + hs;
        for (int j = sub; j < L1; j += NSUB) {
            C2 v[R1];
#pragma unroll
            for (int a = 0; a < R1; ++a) {
                const int j0 = (2 * (j + L1 * a) + p.ps) & (M - 1);
                if (PADDED && j0 >= p.N) {
                    v[a] = c2zero();
                } else {
                    const int ad = p.lay.addr(j0);
                    const fl2 x0 = *reinterpret_cast<const fl2 *>(xf0 + ad);
                    const fl2 x1 = *reinterpret_cast<const fl2 *>(xf1 + ad);
                    const fl2 w = GEO::t2(c.win + j0);
                    v[a].x = mk2(x0.x * w.x, x1.x * w.x);      // scalar multiplies transpose for free
                    v[a].y = mk2(x0.y * w.y, x1.y * w.y);
                }
            }
            Dft<R1, false, f2>::run(v);
            float *wj = c.work + (j * 2) * TF + 2 * l;
            const float *twj = c.tw + GEO::TW1 + 4 * j * R1;
            wstore(wj, TF, v[0]);
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) {
                const fl4 t = GEO::tw4(twj + 4 * cc);
                wstore(wj + (L1 * cc * 2) * TF, TF, cmulw(v[cc], bc(t.x), bc(t.y), bc(t.w)));
            }
        }
    }
    VHD static void pass1(const PkFwdCtx &c, int tid) {
        if (c.a.padded) pass1_t<true>(c, tid);
        else pass1_t<false>(c, tid);
    }

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
        char *ob;            // byte pointer of (row 0, column of slice 2l)
        int64_t rb;          // row pitch in bytes
        bool live0, live1;   // slice 2l / 2l+1 inside [0, P)
    };

    // complex output of one bin for the two slices (r0,i0) (r1,i1); A16: 16-byte vector store
    template <bool A16>
    VHD static void put_c(const Out &o, int row, float r0, float i0, float r1, float i1) {
        char *q = o.ob + row * o.rb;
        if (A16 && o.live1) {
            fl4 v; v.x = r0; v.y = i0; v.z = r1; v.w = i1;
            *reinterpret_cast<fl4 *>(q) = v;
        } else {
            if (o.live0) { fl2 v; v.x = r0; v.y = i0; *reinterpret_cast<fl2 *>(q) = v; }
            if (o.live1) { fl2 v; v.x = r1; v.y = i1; *reinterpret_cast<fl2 *>(q + 8) = v; }
        }
    }
    template <bool A16>
    VHD static void put_r(const Out &o, int row, f2 v) {
        char *q = o.ob + row * o.rb;
        if (A16 && o.live1) {
            *reinterpret_cast<f2 *>(q) = v;
        } else {
            if (o.live0) *reinterpret_cast<float *>(q) = lo(v);
            if (o.live1) *reinterpret_cast<float *>(q + 4) = hi(v);
        }
    }

    // one-sided bin k of both slices, X = E (+/-) D handled by the caller through `sgn`
    //   plus = true :  X[k]     = E + D
    //   plus = false:  X[Mc-k]  = conj(E - D)
    template <bool ONES, bool SPEC, bool A16>
    VHD static void store_bin(const FwdArgs<float> &p, const Out &o, int k, C2 E, C2 D, bool plus) {
        const bool edge = (k == 0 || k == MC);
        const bool scale = ONES && p.mode == 3 && !edge;                 // src/lib.rs:397-403
        if (SPEC) {
            C2 X;
            if (plus) { X.x = E.x + D.x; X.y = E.y + D.y; }
            else { X.x = E.x - D.x; X.y = D.y - E.y; }
            if (edge) X.y = bc(0.f);
            if (scale) { X.x = X.x * bc(p.fac2x); X.y = X.y * bc(p.fac2x); }
            const f2 v = vfma(X.y, X.y, X.x * X.x);
            if (ONES) { put_r<A16>(o, k, v); return; }
            put_r<A16>(o, (k + p.shift) & (M - 1), v);
            if (!edge) put_r<A16>(o, (M - k + p.shift) & (M - 1), v);
            return;
        }
        // the last add is scalar: it moves (re,re)(im,im) into {re,im,re,im} for free
        float r0, i0, r1, i1;
        if (plus) {
            r0 = lo(E.x) + lo(D.x); i0 = lo(E.y) + lo(D.y);
            r1 = hi(E.x) + hi(D.x); i1 = hi(E.y) + hi(D.y);
        } else {
            r0 = lo(E.x) - lo(D.x); i0 = lo(D.y) - lo(E.y);
            r1 = hi(E.x) - hi(D.x); i1 = hi(D.y) - hi(E.y);
        }
        if (edge) { i0 = 0.f; i1 = 0.f; }                                // src/lib.rs:378-384
        if (scale) { r0 *= p.fac2x; i0 *= p.fac2x; r1 *= p.fac2x; i1 *= p.fac2x; }
        if (ONES) { put_c<A16>(o, k, r0, i0, r1, i1); return; }
        put_c<A16>(o, (k + p.shift) & (M - 1), r0, i0, r1, i1);         // fftshift for centered
        if (!edge) put_c<A16>(o, (M - k + p.shift) & (M - 1), r0, -i0, r1, -i1);
    }

    // A = Z[k]/2, Zp = Z[Mc-k]/2  ->  X[k] (and X[Mc-k])
    template <bool ONES, bool SPEC, bool A16>
    VHD static void emit(const FwdArgs<float> &p, const float *twp, const Out &o, int k, C2 A, C2 Zp, bool both) {
        C2 E, Om, D;
        E.x = A.x + Zp.x;  E.y = A.y - Zp.y;                 // A + conj(Zp)
        Om.x = A.x - Zp.x; Om.y = A.y + Zp.y;                // A - conj(Zp)
        const fl4 t = GEO::tw4(twp + 4 * k);                 // (wr, wi, -wr, -wi) of W_M^k
        D.x = vfma(bc(t.x), Om.y, bc(t.y) * Om.x);           // (-i W) Om
        D.y = vfma(bc(t.z), Om.x, bc(t.y) * Om.y);
        store_bin<ONES, SPEC, A16>(p, o, k, E, D, true);
        if (both) store_bin<ONES, SPEC, A16>(p, o, MC - k, E, D, false);
    }

    template <bool ONES, bool SPEC, bool A16>
    VHD static void pass_last_t(const PkFwdCtx &c, int tid) {
        const FwdArgs<float> &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        const int64_t pl = c.pt * TF + 2 * l;
        const int64_t esz = SPEC ? 4 : 8;
        Out o;
        o.live0 = pl < p.P;
        o.live1 = pl + 1 < p.P;
        o.rb = p.ldp * esz;
        o.ob = reinterpret_cast<char *>(p.out) + ((c.b * p.F) * p.ldp + pl) * esz;
        const float *twp = c.tw + GEO::TWP;
        for (int s = sub; s < G / 2; s += NSUB) {
            const int g = s, gp = (s == 0) ? G / 2 : G - s;
            C2 u[RL], up[RL];
            const float *w1 = c.work + (GEO::goff(g) * 2) * TF + 2 * l;
            const float *w2 = c.work + (GEO::goff(gp) * 2) * TF + 2 * l;
#pragma unroll
            for (int d = 0; d < RL; ++d) {
                u[d] = wload(w1 + (d * 2) * TF, TF);
                up[d] = wload(w2 + (d * 2) * TF, TF);
            }
            Dft<RL, false, f2>::run(u);
            Dft<RL, false, f2>::run(up);
            if (!o.live0) continue;
            if (s != 0) {
#pragma unroll
                for (int d = 0; d < RL; ++d)
                    emit<ONES, SPEC, A16>(p, twp, o, g + G * d, u[d], up[RL - 1 - d], true);
            } else {
                emit<ONES, SPEC, A16>(p, twp, o, 0, u[0], u[0], true);
#pragma unroll
                for (int d = 1; d < RL / 2; ++d) emit<ONES, SPEC, A16>(p, twp, o, G * d, u[d], u[RL - d], true);
                emit<ONES, SPEC, A16>(p, twp, o, MC / 2, u[RL / 2], u[RL / 2], false);
#pragma unroll
                for (int d = 0; d < RL / 2; ++d)
                    emit<ONES, SPEC, A16>(p, twp, o, G / 2 + G * d, up[d], up[RL - 1 - d], true);
            }
        }
    }
    VHD static void pass_last(const PkFwdCtx &c, int tid) {
        const FwdArgs<float> &p = c.a;
        const bool ones = p.mode >= 2;
        // vector stores of two slices need an even row pitch and an aligned base
        const int64_t esz = p.spectrogram ? 4 : 8;
        const bool a16 = (p.ldp % 2 == 0) && ((((size_t)p.out) & (2 * esz - 1)) == 0);
        if (p.spectrogram) {
            if (ones) { if (a16) pass_last_t<true, true, true>(c, tid); else pass_last_t<true, true, false>(c, tid); }
            else { if (a16) pass_last_t<false, true, true>(c, tid); else pass_last_t<false, true, false>(c, tid); }
        } else {
            if (ones) { if (a16) pass_last_t<true, false, true>(c, tid); else pass_last_t<true, false, false>(c, tid); }
            else { if (a16) pass_last_t<false, false, true>(c, tid); else pass_last_t<false, false, false>(c, tid); }
        }
    }
};

// ================================================================== INVERSE (packed)
struct PkInvCtx {
    InvArgs<float> a;
    const float *tw, *dual;
    float *work;             // planar [MC][2][TF]
    float *ola;              // [TF][ostride] -- ALIASES work
    int64_t b, qa;
};

struct Raw {                 // one bin of both slices as loaded: {re0, im0, re1, im1}
    float r0, i0, r1, i1;
};

template <int LOGMC, int TF_, int NSUB_>
struct PkInv : PkGeo<LOGMC, TF_, NSUB_> {
    using GEO = PkGeo<LOGMC, TF_, NSUB_>;
    using GEO::MC; using GEO::M; using GEO::TF; using GEO::LN; using GEO::NSUB; using GEO::NT;
    using GEO::NP; using GEO::R1; using GEO::R2; using GEO::R3; using GEO::RL; using GEO::L1;
    using GEO::L2; using GEO::G;
    static constexpr int GPT = (G + NSUB - 1) / NSUB;

    struct In {
        const char *sb;      // byte pointer of (row 0, column of slice 2l)
        int64_t rb;
        bool live0, live1;
    };

    template <bool A16>
    VHD static Raw get(const In &in, int row) {
        const char *q = in.sb + row * in.rb;
        Raw r;
        if (A16 && in.live0 && in.live1) {
            const fl4 v = *reinterpret_cast<const fl4 *>(q);
            r.r0 = v.x; r.i0 = v.y; r.r1 = v.z; r.i1 = v.w;
        } else {
            r.r0 = r.i0 = r.r1 = r.i1 = 0.f;
            if (in.live0) { const fl2 v = *reinterpret_cast<const fl2 *>(q); r.r0 = v.x; r.i0 = v.y; }
            if (in.live1) { const fl2 v = *reinterpret_cast<const fl2 *>(q + 8); r.r1 = v.x; r.i1 = v.y; }
        }
        return r;
    }

    // Hermitian one-sided bin k (src/lib.rs:414-487); two-sided layouts: S[k] + conj(S[M-k])
    template <bool ONES, bool A16>
    VHD static Raw load_bin(const InvArgs<float> &p, const In &in, int k) {
        Raw v;
        if (ONES) {
            v = get<A16>(in, k);
            if (p.mode == 3 && k >= 1 && k <= MC - 1) {
                v.r0 *= p.rfac2x; v.i0 *= p.rfac2x; v.r1 *= p.rfac2x; v.i1 *= p.rfac2x;
            }
        } else {
            const Raw a = get<A16>(in, (k + p.shift) & (M - 1));
            const Raw bq = get<A16>(in, (M - k + p.shift) & (M - 1));
            v.r0 = a.r0 + bq.r0; v.i0 = a.i0 - bq.i0; v.r1 = a.r1 + bq.r1; v.i1 = a.i1 - bq.i1;
        }
        if (k == 0 || k == MC) { v.i0 = 0.f; v.i1 = 0.f; }
        return v;
    }

    // (X[k], X[Mc-k]) -> (Z'[k], Z'[Mc-k]); the first adds are scalar and pack the two slices
    VHD static void merge(const float *twp, int k, Raw A, Raw B, C2 &za, C2 &zb) {
        C2 E, Dm;
        E.x = mk2(A.r0 + B.r0, A.r1 + B.r1);  E.y = mk2(A.i0 - B.i0, A.i1 - B.i1);   // Xk + conj(Xm)
        Dm.x = mk2(A.r0 - B.r0, A.r1 - B.r1); Dm.y = mk2(A.i0 + B.i0, A.i1 + B.i1);  // Xk - conj(Xm)
        const fl4 t = GEO::tw4(twp + 4 * k);
        const C2 O = cmulwc(Dm, bc(t.x), bc(t.y), bc(t.w));                           // conj(W) Dm
        za.x = E.x - O.y; za.y = E.y + O.x;                                           // E + iO
        zb.x = E.x + O.y; zb.y = O.x - E.y;                                           // conj(E - iO)
    }

    template <bool ONES, bool A16>
    VHD static void pass1_t(const PkInvCtx &c, int tid) {
        const InvArgs<float> &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        const int64_t q0 = c.qa + 2 * l;
        const int64_t col = q0 - p.p_min;
        In in;
        in.live0 = q0 >= p.q0 && q0 < p.q1 && col >= 0 && col < p.P;
        in.live1 = q0 + 1 >= p.q0 && q0 + 1 < p.q1 && col + 1 >= 0 && col + 1 < p.P;
        in.rb = p.ldp * 8;
        in.sb = reinterpret_cast<const char *>(p.S) + ((c.b * p.F) * p.ldp + col) * 8;
        const float *twp0 = c.tw + GEO::TWP;
        const bool any = in.live0 || in.live1;
        for (int s = sub; s < L1 / 2; s += NSUB) {
            const int j = s, jp = (s == 0) ? L1 / 2 : L1 - s;
            C2 v[R1], vp[R1];
            if (any) {
                if (s != 0) {
#pragma unroll
                    for (int a = 0; a < R1; ++a) {
                        const Raw xa = load_bin<ONES, A16>(p, in, j + L1 * a);
                        const Raw xb = load_bin<ONES, A16>(p, in, jp + L1 * (R1 - 1 - a));
                        merge(twp0, j + L1 * a, xa, xb, v[a], vp[R1 - 1 - a]);
                    }
                } else {
                    const Raw x0 = load_bin<ONES, A16>(p, in, 0);
                    const Raw xn = load_bin<ONES, A16>(p, in, MC);       // Nyquist pairs with DC
                    v[0].x = mk2(x0.r0 + xn.r0, x0.r1 + xn.r1);
                    v[0].y = mk2(x0.r0 - xn.r0, x0.r1 - xn.r1);
#pragma unroll
                    for (int a = 1; a < R1 / 2; ++a) {
                        const Raw xa = load_bin<ONES, A16>(p, in, L1 * a);
                        const Raw xb = load_bin<ONES, A16>(p, in, L1 * (R1 - a));
                        merge(twp0, L1 * a, xa, xb, v[a], v[R1 - a]);
                    }
                    {
                        const Raw xm = load_bin<ONES, A16>(p, in, MC / 2);
                        C2 dummy;
                        merge(twp0, MC / 2, xm, xm, v[R1 / 2], dummy);
                    }
#pragma unroll
                    for (int a = 0; a < R1 / 2; ++a) {
                        const Raw xa = load_bin<ONES, A16>(p, in, jp + L1 * a);
                        const Raw xb = load_bin<ONES, A16>(p, in, jp + L1 * (R1 - 1 - a));
                        merge(twp0, jp + L1 * a, xa, xb, vp[a], vp[R1 - 1 - a]);
                    }
                }
            } else {
#pragma unroll
                for (int a = 0; a < R1; ++a) v[a] = vp[a] = c2zero();
            }
            Dft<R1, true, f2>::run(v);
            Dft<R1, true, f2>::run(vp);
            float *wj = c.work + (j * 2) * TF + 2 * l, *wp = c.work + (jp * 2) * TF + 2 * l;
            const float *twj = c.tw + GEO::TW1 + 4 * j * R1, *twq = c.tw + GEO::TW1 + 4 * jp * R1;
            wstore(wj, TF, v[0]);
            wstore(wp, TF, vp[0]);
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) {
                const fl4 t = GEO::tw4(twj + 4 * cc), tq = GEO::tw4(twq + 4 * cc);
                wstore(wj + (L1 * cc * 2) * TF, TF, cmulwc(v[cc], bc(t.x), bc(t.y), bc(t.w)));
                wstore(wp + (L1 * cc * 2) * TF, TF, cmulwc(vp[cc], bc(tq.x), bc(tq.y), bc(tq.w)));
            }
        }
    }
    VHD static void pass1(const PkInvCtx &c, int tid) {
        const InvArgs<float> &p = c.a;
        const bool a16 = (p.ldp % 2 == 0) && ((((size_t)p.S) & 15) == 0) && (((c.qa - p.p_min) & 1) == 0);
        if (p.mode >= 2) { if (a16) pass1_t<true, true>(c, tid); else pass1_t<true, false>(c, tid); }
        else { if (a16) pass1_t<false, true>(c, tid); else pass1_t<false, false>(c, tid); }
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
    // roll back (src/lib.rs:499-500), truncate to N, dual window; one ola row per slice
    VHD static void last_store(const PkInvCtx &c, int tid, const C2 (&r)[GPT][RL]) {
        const InvArgs<float> &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        const int ps = p.ps, N = p.N;
        float *row0 = c.ola + (size_t)(2 * l) * p.ostride, *row1 = row0 + p.ostride;
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = sub + NSUB * i;
            if (g < G) {
#pragma unroll
                for (int d = 0; d < RL; ++d) {
                    const int m = g + G * d;
                    const int j0 = (2 * m + ps) & (M - 1);
                    const int j1 = (j0 + 1) & (M - 1);
                    if (j0 < N) {
                        const float dw = GEO::t1(c.dual + j0);
                        row0[j0] = lo(r[i][d].x) * dw;
                        row1[j0] = hi(r[i][d].x) * dw;
                    }
                    if (j1 < N) {
                        const float dw = GEO::t1(c.dual + j1);
                        row0[j1] = lo(r[i][d].y) * dw;
                        row1[j1] = hi(r[i][d].y) * dw;
                    }
                }
            }
        }
    }

    // overlap-add gather over the samples this tile owns, increasing q (same as scalar family)
    VHD static void gather(const PkInvCtx &c, int tid) {
        const InvArgs<float> &p = c.a;
        const int H = p.H, N = p.N, halo = p.halo;
        const int64_t qo = c.qa + halo;
        int64_t ka = qo * H - p.mid, kb = (c.qa + TF) * (int64_t)H - p.mid;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        const int64_t base = c.qa * (int64_t)H - p.mid;
        float *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const int cnt = (TF - halo) * H, hH = halo * H, step = p.ostride - H;
        if (halo == 3 && N == 4 * H) {                       // 75 % overlap: fully unrolled
            for (int e = tid; e < cnt; e += NT) {
                const int u = e + hH;
                if (u < ua || u >= ub) continue;
                const int r = (int)mulhi_u32((unsigned)e, p.magic_h);
                const int idx = r * p.ostride + (e - r * H) + hH;
                float acc = c.ola[idx];
                acc += c.ola[idx + step];
                acc += c.ola[idx + 2 * step];
                acc += c.ola[idx + 3 * step];
                xo[u] = acc;
            }
            return;
        }
        for (int e = tid; e < cnt; e += NT) {
            const int u = e + hH;
            if (u < ua || u >= ub) continue;
            const int r = (int)mulhi_u32((unsigned)e, p.magic_h);
            int jj = e - r * H + hH;
            int idx = r * p.ostride + jj;
            float acc = 0.f;
            for (int h = 0; h <= halo; ++h) {
                if (jj < N) acc += c.ola[idx];
                jj -= H;
                idx += step;
            }
            xo[u] = acc;
        }
    }
};

}  // namespace p2
}  // namespace veils
