// wide_impl.cuh -- the "wide" f32 kernel family for mfft = 1024 / 2048 / 4096 with the default
// geometry (m_num = mfft = 4 hop, phase_shift = 0): the shapes of BASELINE c3 / c4 and the
// 1024/256 and 2048/512 front ends.  Written as phase functions that compile for the device
// (wide.cu) and for the host emulation (tests/emu/wide_emu.cu), which also counts shared-memory
// bank conflicts per warp instruction.
//
// Why a second family: in the older families a tile of mfft >= 2048 holds only 8 / 4 slices, a warp
// touches 8 short rows per shared-memory instruction and the row addresses collide on the banks
// (ncu, profiles/r2a_ncu_full_c3_before.txt: 28 % / 63 % of all shared wavefronts are conflicts,
// LSU pipe 68 %, DRAM 18-26 %).  Here:
//   * thread = (sub, l): lane l owns TWO slices of the tile, packed in f32x2 registers (FADD2 /
//     FFMA2); a warp covers SUBW = 32 / LN consecutive butterfly units (LN = TF / 2);
//   * the work buffer is planar per element, [re: TF floats][im: TF floats], and the position of
//     the two halves is SWIZZLED by the element index so that the rows of any RH = 16 / LN
//     consecutive elements fall on disjoint banks: every LDS.64 / STS.64 wavefront moves 128
//     useful bytes (emulator-checked: 0 conflicts);
//   * three passes R1 x R2 x R3 over Mc = mfft / 2 complex points (real-input FFT through the
//     half-length transform).  The middle pass writes its outputs at j2 ^ (cb & XM) -- a
//     permutation inside groups that one warp owns, so __syncwarp() is all it needs -- which makes
//     the last pass read consecutive elements for consecutive units as well;
//   * forward: pass 1 reads the staged signal (frame extraction, zero padding outside [0, n),
//     roll by m_num / 2 and the window are load-index arithmetic and one table read;
//     src/lib.rs:336-345, 612-652, 715-729); the last pass holds the mirrored groups (g, G - g)
//     in one thread, splits the half-length spectrum into the one-sided bins and stores them
//     straight from registers in the reference's [f][p] order (src/lib.rs:354-408, 735-746);
//   * inverse: pass 1 loads S (Hermitian merge of bins k and Mc - k, src/lib.rs:414-487), the last
//     pass multiplies by the dual window and overlap-adds into a padded accumulator in four
//     quarter passes, oldest slice first = the reference's increasing-q order (src/lib.rs:853-910).
#pragma once
#include "pow2_cfg.hpp"

namespace veils {
namespace wd {

using namespace p2;
using C2 = cx<f2>;
struct alignas(8) fl2 { float x, y; };
struct alignas(16) fl4 { float x, y, z, w; };

// ---- shared-memory accessors (the host emulation can trace them) -----------------------------
#if !defined(__CUDA_ARCH__) && defined(VEILS_WIDE_TRACE)
void wide_trace(int kind, const void *p, int bytes);      // tests/emu/wide_emu.cu
#define WD_TRACE(kind, p, b) wide_trace(kind, p, b)
#else
#define WD_TRACE(kind, p, b)
#endif
VHD f2 lds2(const float *p) {
    WD_TRACE(0, p, 8);
    return *reinterpret_cast<const f2 *>(p);
}
VHD void sts2(float *p, f2 v) {
    WD_TRACE(1, p, 8);
    *reinterpret_cast<f2 *>(p) = v;
}
VHD fl2 ldsx(const float *p) {
    WD_TRACE(0, p, 8);
    return *reinterpret_cast<const fl2 *>(p);
}
VHD void stsx(float *p, fl2 v) {
    WD_TRACE(1, p, 8);
    *reinterpret_cast<fl2 *>(p) = v;
}
// plan-table entry: the kernels copy the tables to shared memory once per (persistent) CTA
VHD fl2 ldg2(const fl2 *p) {
    WD_TRACE(0, p, 8);
    return *p;
}
VHD void syncwarp() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}
VHD float mul_rn(float x, float y) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(x, y);
#else
    return x * y;
#endif
}
VHD float add_rn(float x, float y) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(x, y);
#else
    return x + y;
#endif
}

// ---- radix plans -------------------------------------------------------------------------------
// forward (F1, F2, F3): F3 = 8 because the last pass keeps two groups per thread;
// inverse (I1, I2, I3): I1 = 8 because pass 1 keeps two groups per thread.
template <int LOGMC, int TF> struct WPlan;
template <> struct WPlan<9, 16>  { static constexpr int F1 = 8,  F2 = 8,  F3 = 8, I1 = 8, I2 = 8,  I3 = 8,  NT = 256; };
template <> struct WPlan<10, 16> { static constexpr int F1 = 16, F2 = 8,  F3 = 8, I1 = 8, I2 = 8,  I3 = 16, NT = 512; };
template <> struct WPlan<10, 8>  { static constexpr int F1 = 8,  F2 = 16, F3 = 8, I1 = 8, I2 = 16, I3 = 8,  NT = 256; };
template <> struct WPlan<11, 8>  { static constexpr int F1 = 16, F2 = 16, F3 = 8, I1 = 8, I2 = 16, I3 = 16, NT = 512; };

template <int LOGMC, int TF_>
struct WGeo {
    using PL = WPlan<LOGMC, TF_>;
    static constexpr int MC = 1 << LOGMC, M = 2 * MC, H = M / 4, MID = M / 2;
    static constexpr int TF = TF_, LN = TF_ / 2, NT = PL::NT, NSUB = NT / LN, SUBW = 32 / LN, NWARP = NT / 32;
    static constexpr bool ADJ = TF_ == 8;              // lane l owns slices (2l, 2l+1); else (l, l + LN)
    static constexpr int XM = 16 / LN - 1;             // rows per LDS.64 wavefront - 1: swizzle mask
    static constexpr int HS = H + 4;                   // row pitch of the staged signal / the accumulator:
                                                       // lanes (slices) land on disjoint banks
    static constexpr int ROWS = TF + 3;                // hop rows under one tile (75 % overlap)
    static constexpr int WORK_FLOATS = MC * 2 * TF;
    static constexpr int XIN_FLOATS = ROWS * HS;
    VHD static int slice_of(int l, int h) { return ADJ ? 2 * l + h : l + h * LN; }
    // float offset of the re / im row of element e; im = re ^ TF
    VHD static int swz(int e) { return TF == 8 ? ((e >> 1) & 1) : (e & 1); }
    VHD static int off_re(int e) { return 2 * TF * e + TF * swz(e); }
    VHD static int off_im(int e) { return off_re(e) ^ TF; }

    // ---- pass 2: unit (cb, j2) transforms elements cb L1 + j2 + L2 a in place, output c2 goes to
    //      cb L1 + c2 L2 + (j2 ^ (cb & XM)); the XOR stays inside a group of XM + 1 consecutive j2,
    //      i.e. inside one warp: __syncwarp() orders its reads before its writes.
    template <bool INV, int RA, int RB, int LB>
    VHD static void pass2_load(const float *work, int u, int l, C2 (&v)[RB]) {
        constexpr int LA = RB * LB;                    // elements per block
        const int cb = u / LB, j2 = u - cb * LB;
        const float *wre = work + off_re(cb * LA + j2) + 2 * l;
        const float *wim = work + off_im(cb * LA + j2) + 2 * l;
#pragma unroll
        for (int a = 0; a < RB; ++a) {
            v[a].x = lds2(wre + a * (2 * TF * LB));
            v[a].y = lds2(wim + a * (2 * TF * LB));
        }
        Dft<RB, INV, f2>::run(v);
    }
    template <bool INV, int RA, int RB, int LB>
    VHD static void pass2_store(float *work, const fl2 *tw2, int u, int l, const C2 (&v)[RB]) {
        constexpr int LA = RB * LB;
        const int cb = u / LB, j2 = u - cb * LB;
        const int jx = j2 ^ (cb & XM);
        float *wre = work + off_re(cb * LA + jx) + 2 * l;
        float *wim = work + off_im(cb * LA + jx) + 2 * l;
        const fl2 *tw = tw2 + j2;
        sts2(wre, v[0].x);
        sts2(wim, v[0].y);
#pragma unroll
        for (int cc = 1; cc < RB; ++cc) {
            const fl2 t = ldg2(tw + cc * LB);
            const C2 z = INV ? cmulwc(v[cc], bc(t.x), bc(t.y), bc(-t.y)) : cmulw(v[cc], bc(t.x), bc(t.y), bc(-t.y));
            sts2(wre + cc * (2 * TF * LB), z.x);
            sts2(wim + cc * (2 * TF * LB), z.y);
        }
    }
    template <int RA, int LB> static constexpr int pass2_iters() { return (RA * LB + NSUB - 1) / NSUB; }
    // device flavour: the whole pass (units u = sub + it NSUB); the host emulation runs the two halves of
    // every iteration for all threads in turn
    template <bool INV, int RA, int RB, int LB>
    VHD static void pass2_t(float *work, const fl2 *tw2, int tid) {
        const int l = tid % LN, sub = tid / LN;
#pragma unroll 1
        for (int u = sub; u < RA * LB; u += NSUB) {
            C2 v[RB];
            pass2_load<INV, RA, RB, LB>(work, u, l, v);
            syncwarp();
            pass2_store<INV, RA, RB, LB>(work, tw2, u, l, v);
        }
    }

    // the RL elements of last-pass group (cb, c2): element d sits at cb L1 + c2 L2 + (d ^ (cb & XM))
    template <int RR, int LA, int LB>
    VHD static void ld_grp(const float *work, int cb, int c2, int l, C2 (&u)[RR]) {
        const int x = cb & XM;
        const float *base = work + 2 * TF * (cb * LA + c2 * LB) + 2 * l;
        const float *pre[XM + 1], *pim[XM + 1];
#pragma unroll
        for (int dl = 0; dl <= XM; ++dl) {
            pre[dl] = base + off_re(dl ^ x);
            pim[dl] = base + off_im(dl ^ x);
        }
#pragma unroll
        for (int d = 0; d < RR; ++d) {
            u[d].x = lds2(pre[d & XM] + (d & ~XM) * (2 * TF));
            u[d].y = lds2(pim[d & XM] + (d & ~XM) * (2 * TF));
        }
    }
};

// ================================================================== FORWARD
struct WFwdArgs {
    const float *x;
    void *out;
    const fl2 *tab;          // plan table in global memory (layout: wide_table_fill)
    int64_t n, x_pitch, ldp, p0, P, k_offset, ptiles, F;
    int mode;                // 0 twosided, 1 centered, 2 onesided, 3 onesided2X
    int spectrogram;
    int vec16;               // S base and row pitch allow 16-byte stores
    int no_bulk;             // debug: stage every tile through cp.async
    float fac2x;
};
struct WFwdCtx {
    WFwdArgs a;
    const fl2 *wtab;         // [Mc]  0.5 * (w[j0], w[j0 + 1]), j0 = (2m + M/2) mod M  (roll folded in)
    const fl2 *tw1;          // [F1][L1]   W_Mc^(j c)
    const fl2 *tw2;          // [F2][L2]   W_L1^(j2 c2)
    const fl2 *twp;          // [Mc + 1]   W_M^k          (shared-memory copies on the device)
    float *xin;              // staged signal [ROWS][HS]
    float *work;             // [Mc][2][TF], swizzled
    int64_t b, pt;
};

template <int LOGMC, int TF_>
struct WFwd : WGeo<LOGMC, TF_> {
    using GEO = WGeo<LOGMC, TF_>;
    using GEO::MC; using GEO::M; using GEO::H; using GEO::MID; using GEO::TF; using GEO::LN; using GEO::NT;
    using GEO::NSUB; using GEO::SUBW; using GEO::NWARP; using GEO::ADJ; using GEO::XM; using GEO::HS; using GEO::ROWS;
    static constexpr int R1 = GEO::PL::F1, R2 = GEO::PL::F2, RL = GEO::PL::F3;
    static constexpr int L1 = MC / R1, L2 = L1 / R2, G = MC / RL;
    static_assert(L2 == RL, "three passes");
    static_assert(R1 % SUBW == 0 && L2 % SUBW == 0 && L1 % SUBW == 0, "a warp owns whole unit groups");

    // ---- staging of the tile's samples, zero outside [0, n) (src/lib.rs:631-645); cp.async
    VHD static void fill(const WFwdCtx &c, int tid) {
        const WFwdArgs &p = c.a;
        const int64_t s0 = (p.p0 + c.pt * TF) * (int64_t)H + p.k_offset - MID;
        const float *xb = p.x + c.b * p.x_pitch;
        const bool al = (((c.b * p.x_pitch + s0) & 3) == 0) && ((((size_t)p.x) & 15) == 0);
        constexpr int CPR = H / 4;                     // 16-byte chunks per row
        for (int e = tid; e < ROWS * CPR; e += NT) {
            const int r = e / CPR, cq = e - r * CPR;
            const int64_t s = s0 + (int64_t)r * H + 4 * cq;
            float *dst = c.xin + r * HS + 4 * cq;
            if (al && s >= 0 && s + 4 <= p.n) {
                cp_async_zfill<16>(dst, xb + s, true);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool ok = s + i >= 0 && s + i < p.n;
                    cp_async_zfill<4>(dst + i, ok ? xb + s + i : xb, ok);
                }
            }
        }
    }

    // ---- pass 1: group jg reads sample pairs m = jg + L1 a of its two slices, DFT_R1, twiddle
    VHD static void pass1(const WFwdCtx &c, int tid) {
        const int l = tid % LN, sub = tid / LN;
#pragma unroll 1
        for (int jg = sub; jg < L1; jg += NSUB) {
            C2 v[R1];
            const float *xa = c.xin + GEO::slice_of(l, 0) * HS + 2 * jg;
            const float *xb = c.xin + GEO::slice_of(l, 1) * HS + 2 * jg;
            const fl2 *wt = c.wtab + jg;
#pragma unroll
            for (int a = 0; a < R1; ++a) {
                // FFT slot 2m, m = jg + L1 a, holds sample j0 = (2m + M/2) mod M of the slice (the roll
                // of src/lib.rs:343-345): hop row j0 / H, column j0 % H -- compile-time but for 2 jg
                const int c0 = (2 * L1 * a + M / 2) % M;
                const int off = (c0 / H) * HS + (c0 % H);
                const fl2 x0 = ldsx(xa + off), x1 = ldsx(xb + off);
                const fl2 w = ldg2(wt + L1 * a);
                v[a].x = mk2(x0.x * w.x, x1.x * w.x);
                v[a].y = mk2(x0.y * w.y, x1.y * w.y);
            }
            Dft<R1, false, f2>::run(v);
            float *wre = c.work + GEO::off_re(jg) + 2 * l;
            float *wim = c.work + GEO::off_im(jg) + 2 * l;
            const fl2 *tw = c.tw1 + jg;
            sts2(wre, v[0].x);
            sts2(wim, v[0].y);
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) {
                const fl2 t = ldg2(tw + cc * L1);
                const C2 z = cmulw(v[cc], bc(t.x), bc(t.y), bc(-t.y));
                sts2(wre + cc * (2 * TF * L1), z.x);
                sts2(wim + cc * (2 * TF * L1), z.y);
            }
        }
    }

    VHD static void pass2(const WFwdCtx &c, int tid) { GEO::template pass2_t<false, R1, R2, L2>(c.work, c.tw2, tid); }

    // ---- output of one bin for the thread's two slices
    struct Out {
        char *base;          // S + ((b F) ldp + column of slice h = 0) * esz
        int64_t rb;          // row pitch in bytes
        bool live0, live1, vec;
    };
    // FULL: both slices of the lane are inside [0, P) (and, adjacent pairing, the pair is one aligned
    // 16-byte store): no predication.  q = pointer of (row, column of slice h = 0).
    template <bool SPEC, bool FULL>
    VHD static void put(const Out &o, char *q, float r0, float i0, float r1, float i1) {
        if (SPEC) {
            const float v0 = r0 * r0 + i0 * i0, v1 = r1 * r1 + i1 * i1;   // scipy _short_time_fft.py:1398-1400
            if (ADJ && FULL) {
                fl2 t; t.x = v0; t.y = v1;
                *reinterpret_cast<fl2 *>(q) = t;
            } else {
                if (FULL || o.live0) *reinterpret_cast<float *>(q) = v0;
                if (FULL || o.live1) *reinterpret_cast<float *>(q + (ADJ ? 4 : 4 * LN)) = v1;
            }
        } else {
            if (ADJ && FULL) {
                fl4 t; t.x = r0; t.y = i0; t.z = r1; t.w = i1;
                *reinterpret_cast<fl4 *>(q) = t;
            } else {
                if (FULL || o.live0) { fl2 t; t.x = r0; t.y = i0; *reinterpret_cast<fl2 *>(q) = t; }
                if (FULL || o.live1) { fl2 t; t.x = r1; t.y = i1; *reinterpret_cast<fl2 *>(q + (ADJ ? 8 : 8 * LN)) = t; }
            }
        }
    }
    // mode epilogue of one-sided bin k (src/lib.rs:354-408); q = pointer of row k (one-sided layouts)
    template <int MODE, bool SPEC, bool FULL>
    VHD static void emit(const WFwdArgs &p, const Out &o, int k, char *q, float r0, float i0, float r1, float i1) {
        const bool edge = k == 0 || k == MC;
        if (edge) { i0 = 0.f; i1 = 0.f; }                                  // :378-384
        if (MODE >= 2) {
            if (MODE == 3 && !edge) { r0 *= p.fac2x; i0 *= p.fac2x; r1 *= p.fac2x; i1 *= p.fac2x; }   // :397-403
            put<SPEC, FULL>(o, q, r0, i0, r1, i1);
            return;
        }
        const int sh = MODE == 1 ? MC : 0;                                  // fftshift :360-366
        put<SPEC, FULL>(o, o.base + (int64_t)((k + sh) & (M - 1)) * o.rb, r0, i0, r1, i1);
        if (!edge) put<SPEC, FULL>(o, o.base + (int64_t)((M - k + sh) & (M - 1)) * o.rb, r0, -i0, r1, -i1);
    }
    // A = Z[k] / 2, B = Z[Mc - k] / 2 of both slices, t = W_M^k:  X[k] -> row pointer qk and (both)
    // X[Mc - k] -> qm
    template <int MODE, bool SPEC, bool FULL>
    VHD static void split(const WFwdArgs &p, const Out &o, int k, char *qk, char *qm, fl2 t, C2 A, C2 B, bool both) {
        const f2 Ex = A.x + B.x, Oy = A.y + B.y, ey = A.y - B.y, ox = A.x - B.x;
        const f2 P = vfma(bc(t.y), ox, bc(t.x) * Oy);                       // wi ox + wr Oy
        const f2 Q = vfma(bc(t.y), Oy, bc(-t.x) * ox);                      // wi Oy - wr ox
        // the last adds are scalar: they place (re, im) of each slice in store order for free
        emit<MODE, SPEC, FULL>(p, o, k, qk, lo(Ex) + lo(P), lo(ey) + lo(Q), hi(Ex) + hi(P), hi(ey) + hi(Q));
        if (both) emit<MODE, SPEC, FULL>(p, o, MC - k, qm, lo(Ex) - lo(P), lo(Q) - lo(ey), hi(Ex) - hi(P), hi(Q) - hi(ey));
    }

    // ---- pass 3: warp unit = (c2 < R2/2, run of SUBW consecutive cb): the thread holds last-pass group
    //      gA = cb + R1 c2 and its mirror gB = G - gA, whose bins are each other's mirror images:
    //      Mc - (gA + G d) = gB + G (RL - 1 - d)
    static constexpr int CBR = R1 / SUBW, NU3 = (R2 / 2) * CBR;
    template <int MODE, bool SPEC>
    VHD static void pass3_t(const WFwdCtx &c, int tid) {
        const WFwdArgs &p = c.a;
        const int l = tid % LN;
        Out o;
        {
            const int64_t col0 = c.pt * TF + GEO::slice_of(l, 0), col1 = c.pt * TF + GEO::slice_of(l, 1);
            const int64_t esz = SPEC ? 4 : 8;
            o.live0 = col0 < p.P;
            o.live1 = col1 < p.P;
            o.vec = p.vec16 != 0;
            o.rb = p.ldp * esz;
            o.base = reinterpret_cast<char *>(p.out) + ((c.b * p.F) * p.ldp + col0) * esz;
        }
        if (o.live0 && o.live1 && (!ADJ || o.vec)) pass3_f<MODE, SPEC, true>(c, tid, o);
        else pass3_f<MODE, SPEC, false>(c, tid, o);
    }
    template <int MODE, bool SPEC, bool FULL>
    VHD static void pass3_f(const WFwdCtx &c, int tid, const Out &o) {
        const WFwdArgs &p = c.a;
        const int l = tid % LN, sub = tid / LN, warp = sub / SUBW, sw = sub % SUBW;
#pragma unroll 1
        for (int unit = warp; unit < NU3; unit += NWARP) {
            const int c2 = unit / CBR, cb = (unit - c2 * CBR) * SUBW + sw;
            const int gA = cb + R1 * c2;
            C2 uA[RL], uB[RL];
            if (gA != 0) {
                const int gB = G - gA;
                GEO::template ld_grp<RL, L1, L2>(c.work, cb, c2, l, uA);
                GEO::template ld_grp<RL, L1, L2>(c.work, gB % R1, gB / R1, l, uB);
                Dft<RL, false, f2>::run(uA);
                Dft<RL, false, f2>::run(uB);
                const fl2 *tq = c.twp + gA;
                char *qk = o.base + gA * o.rb, *qm = o.base + (MC - gA) * o.rb;     // rows k and Mc - k
                const int64_t st = G * o.rb;
#pragma unroll
                for (int d = 0; d < RL; ++d) {
                    split<MODE, SPEC, FULL>(p, o, gA + G * d, qk, qm, ldg2(tq + G * d), uA[d], uB[RL - 1 - d], true);
                    qk += st;
                    qm -= st;
                }
            } else {
                // self-mirrored groups 0 (bins G d <-> G (RL - d), DC <-> Nyquist) and G/2
                GEO::template ld_grp<RL, L1, L2>(c.work, 0, 0, l, uA);
                GEO::template ld_grp<RL, L1, L2>(c.work, 0, R2 / 2, l, uB);
                Dft<RL, false, f2>::run(uA);
                Dft<RL, false, f2>::run(uB);
                auto row = [&](int k) { return o.base + (int64_t)k * o.rb; };
                split<MODE, SPEC, FULL>(p, o, 0, row(0), row(MC), ldg2(c.twp), uA[0], uA[0], true);      // X[0], X[Mc]
#pragma unroll
                for (int d = 1; d < RL / 2; ++d)
                    split<MODE, SPEC, FULL>(p, o, G * d, row(G * d), row(MC - G * d), ldg2(c.twp + G * d), uA[d], uA[RL - d], true);
                split<MODE, SPEC, FULL>(p, o, MC / 2, row(MC / 2), row(MC / 2), ldg2(c.twp + MC / 2), uA[RL / 2], uA[RL / 2], false);
#pragma unroll
                for (int d = 0; d < RL / 2; ++d)
                    split<MODE, SPEC, FULL>(p, o, G / 2 + G * d, row(G / 2 + G * d), row(MC - G / 2 - G * d),
                                            ldg2(c.twp + G / 2 + G * d), uB[d], uB[RL - 1 - d], true);
            }
        }
    }
    VHD static void pass3(const WFwdCtx &c, int tid) {
        const int m = c.a.mode;
        if (c.a.spectrogram) {
            if (m == 3) pass3_t<3, true>(c, tid);
            else if (m == 2) pass3_t<2, true>(c, tid);
            else if (m == 1) pass3_t<1, true>(c, tid);
            else pass3_t<0, true>(c, tid);
        } else {
            if (m == 3) pass3_t<3, false>(c, tid);
            else if (m == 2) pass3_t<2, false>(c, tid);
            else if (m == 1) pass3_t<1, false>(c, tid);
            else pass3_t<0, false>(c, tid);
        }
    }
};

// ================================================================== INVERSE
struct WInvArgs {
    const void *S;
    float *xo;
    const fl2 *tab;          // plan table in global memory (layout: wide_table_fill)
    int64_t P, ldp, out_pitch, p_min, k0, k1, q0, q1, qh0, qh1, cps, F;
    int chunk_tiles;         // streaming overlap-add: tiles per chunk (a chunk owns TF * chunk_tiles - 4 slices)
    int mode, vec16, vec_out, prefetch;
    float rfac2x;
};
struct WInvCtx {
    WInvArgs a;
    const fl2 *dtab;         // [Mc]  scale * (dual[j0], dual[j0 + 1]), j0 = (2m + M/2) mod M
    const fl2 *tw1;          // [I1][L1]   W_Mc^(j c)
    const fl2 *tw2;          // [I2][L2]   W_L1^(j2 c2)
    const fl2 *twp;          // [Mc + 1]   W_M^k          (shared-memory copies on the device)
    float *work;             // [Mc][2][TF], swizzled
    float *acc;              // [TF + 3][HS] overlap-add accumulator: tile sample u at (u / H) HS + u % H
    int first;               // first tile of a chunk: no carry in the head, owns samples from halo * H on
    int halo;                // slices the first tile of this chunk re-computes (3 or 4: even first column)
    int more;                // another tile of this chunk follows (L2 prefetch of its S columns)
    int64_t b, qa;
    int64_t kend;            // first sample the NEXT chunk owns
};
// Chunk ch of the launch: signal b, first owned slice a0, first computed slice qa0 = a0 - halo.  The halo is 3
// slices (75 % overlap) or 4, whichever puts the tile on an even column of S: the slice pairs of a lane
// are then one aligned 16-byte access and the tile can be described to the TMA engine.
struct WChunk { int64_t b, qa0, kend; int halo; };
VHD WChunk wide_chunk(const WInvArgs &a, int64_t ch, int tf, int H, int mid) {
    const int64_t lc = (int64_t)tf * a.chunk_tiles - 4;
    WChunk c;
    c.b = ch / a.cps;
    const int64_t a0 = a.qh0 + (ch - c.b * a.cps) * lc;
    c.halo = 3 + (int)((a0 - 3 - a.p_min) & 1);
    c.qa0 = a0 - c.halo;
    c.kend = (a0 + lc) * (int64_t)H - mid;
    return c;
}

template <int LOGMC, int TF_>
struct WInv : WGeo<LOGMC, TF_> {
    using GEO = WGeo<LOGMC, TF_>;
    using GEO::MC; using GEO::M; using GEO::H; using GEO::MID; using GEO::TF; using GEO::LN; using GEO::NT;
    using GEO::NSUB; using GEO::SUBW; using GEO::NWARP; using GEO::ADJ; using GEO::XM; using GEO::HS; using GEO::ROWS;
    static constexpr int R1 = GEO::PL::I1, R2 = GEO::PL::I2, RL = GEO::PL::I3;
    static constexpr int L1 = MC / R1, L2 = L1 / R2, G = MC / RL;
    static_assert(L2 == RL, "three passes");
    static_assert(R1 % SUBW == 0 && L2 % SUBW == 0 && (L1 / 2) % SUBW == 0, "a warp owns whole unit groups");

    struct Raw { float r0, i0, r1, i1; };              // one bin of the thread's two slices
    struct In {
        const char *base;    // S + ((b F) ldp + column of slice h = 0) * 8
        int64_t rb;
        bool live0, live1, vec;
    };
    VHD static In make_in(const WInvCtx &c, int l) {
        const WInvArgs &p = c.a;
        const int64_t qa0 = c.qa + GEO::slice_of(l, 0), qa1 = c.qa + GEO::slice_of(l, 1);
        const int64_t col0 = qa0 - p.p_min, col1 = qa1 - p.p_min;
        In in;
        in.live0 = qa0 >= p.q0 && qa0 < p.q1 && col0 >= 0 && col0 < p.P;   // src/lib.rs:855-857
        in.live1 = qa1 >= p.q0 && qa1 < p.q1 && col1 >= 0 && col1 < p.P;
        in.rb = p.ldp * 8;
        in.base = reinterpret_cast<const char *>(p.S) + ((c.b * p.F) * p.ldp + col0) * 8;
        // one 16-byte load for the slice pair: even column, inside the row pitch
        in.vec = ADJ && p.vec16 && (col0 & 1) == 0 && col0 >= 0 && col0 + 1 < p.ldp;
        return in;
    }
    VHD static fl2 ldS(const char *q) {
#if defined(__CUDA_ARCH__)
        float2 v;
        asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(q));
        fl2 r; r.x = v.x; r.y = v.y;
        return r;
#else
        return *reinterpret_cast<const fl2 *>(q);
#endif
    }
    VHD static fl4 ldS4(const char *q) {
#if defined(__CUDA_ARCH__)
        float4 v;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(q));
        fl4 r; r.x = v.x; r.y = v.y; r.z = v.z; r.w = v.w;
        return r;
#else
        return *reinterpret_cast<const fl4 *>(q);
#endif
    }
    VHD static void prefetch_l2(const char *q) {
#if defined(__CUDA_ARCH__)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
#else
        (void)q;
#endif
    }
    VHD static Raw row_at(const In &in, int64_t row, bool pf) {
        const char *q = in.base + row * in.rb;
        Raw v;
        v.r0 = v.i0 = v.r1 = v.i1 = 0.f;
        if (in.vec) {
            if (in.live0 || in.live1) {
                const fl4 t = ldS4(q);
                if (in.live0) { v.r0 = t.x; v.i0 = t.y; }
                if (in.live1) { v.r1 = t.z; v.i1 = t.w; }
            }
        } else {
            if (in.live0) { const fl2 t = ldS(q); v.r0 = t.x; v.i0 = t.y; }
            if (in.live1) { const fl2 t = ldS(q + (ADJ ? 8 : 8 * LN)); v.r1 = t.x; v.i1 = t.y; }
        }
        if (pf) {                                      // the same row of the next tile: TF columns on
            prefetch_l2(q + TF * 8);
            if (!ADJ) prefetch_l2(q + TF * 8 + 8 * LN);
        }
        return v;
    }
    // Hermitian one-sided bin k of the stored layout (src/lib.rs:414-487)
    template <int MODE>
    VHD static Raw load_bin(const WInvArgs &p, const In &in, int k, bool pf) {
        Raw v;
        if (MODE >= 2) {
            v = row_at(in, k, pf);
            if (MODE == 3 && k >= 1 && k <= MC - 1) {                      // :454-487
                v.r0 *= p.rfac2x; v.i0 *= p.rfac2x; v.r1 *= p.rfac2x; v.i1 *= p.rfac2x;
            }
        } else {
            const int sh = MODE == 1 ? MC : 0;                             // ifftshift :421-427, 516-521
            const Raw a = row_at(in, (k + sh) & (M - 1), pf);
            const Raw bq = row_at(in, (M - k + sh) & (M - 1), pf);
            v.r0 = a.r0 + bq.r0; v.i0 = a.i0 - bq.i0; v.r1 = a.r1 + bq.r1; v.i1 = a.i1 - bq.i1;
        }
        return v;
    }
    // A = X[k], B = X[Mc - k], t = W_M^k:  z = Z'[k], zm = Z'[Mc - k]  (Z' = twice the half-length
    // spectrum of z[m] = x[2m] + i x[2m+1]); the first adds are scalar and pack the two slices
    VHD static void merge(fl2 t, const Raw &A, const Raw &B, C2 &z, C2 &zm) {
        const f2 Ex = mk2(A.r0 + B.r0, A.r1 + B.r1), Dy = mk2(A.i0 + B.i0, A.i1 + B.i1);
        const f2 ey = mk2(A.i0 - B.i0, A.i1 - B.i1), dx = mk2(A.r0 - B.r0, A.r1 - B.r1);
        const f2 P = vfma(bc(-t.y), dx, bc(t.x) * Dy);                     // wr Dy - wi dx
        const f2 Q = vfma(bc(t.y), Dy, bc(t.x) * dx);                      // wr dx + wi Dy
        z.x = Ex - P; z.y = ey + Q;
        zm.x = Ex + P; zm.y = Q - ey;
    }

    // ---- pass 1: unit w = pass-1 group pair (jA, jB) = (w, L1 - w), w = 0: (0, L1/2).  The mirror of bin
    //      jA + L1 a is jB + L1 (R1 - 1 - a): both groups of a pair are merged from the same loads.
    template <int MODE>
    VHD static void pass1_t(const WInvCtx &c, int tid) {
        const WInvArgs &p = c.a;
        const int l = tid % LN, sub = tid / LN;
        const In in = make_in(c, l);
        const bool pf = p.prefetch && c.more;
#pragma unroll 1
        for (int w = sub; w < L1 / 2; w += NSUB) {
            C2 vA[R1], vB[R1];
            int jA = w, jB = L1 - w;
            if (w != 0) {
                const fl2 *tq = c.twp + jA;
#pragma unroll
                for (int a = 0; a < R1; ++a) {
                    const Raw XA = load_bin<MODE>(p, in, jA + L1 * a, pf);
                    const Raw XB = load_bin<MODE>(p, in, jB + L1 * (R1 - 1 - a), pf);
                    merge(ldg2(tq + L1 * a), XA, XB, vA[a], vB[R1 - 1 - a]);
                }
            } else {
                jB = L1 / 2;
                // group 0: bin L1 a <-> L1 (R1 - a), DC <-> Nyquist (imaginary parts ignored, irfft semantics)
                {
                    const Raw X0 = load_bin<MODE>(p, in, 0, pf), XN = load_bin<MODE>(p, in, MC, pf);
                    vA[0].x = mk2(X0.r0 + XN.r0, X0.r1 + XN.r1);
                    vA[0].y = mk2(X0.r0 - XN.r0, X0.r1 - XN.r1);
                }
#pragma unroll
                for (int a = 1; a < R1 / 2; ++a) {
                    const Raw XA = load_bin<MODE>(p, in, L1 * a, pf), XB = load_bin<MODE>(p, in, L1 * (R1 - a), pf);
                    merge(ldg2(c.twp + L1 * a), XA, XB, vA[a], vA[R1 - a]);
                }
                {
                    const Raw Xh = load_bin<MODE>(p, in, MC / 2, pf);
                    C2 dump;
                    merge(ldg2(c.twp + MC / 2), Xh, Xh, vA[R1 / 2], dump);
                }
                // group L1/2: bin L1/2 + L1 a <-> L1/2 + L1 (R1 - 1 - a)
#pragma unroll
                for (int a = 0; a < R1 / 2; ++a) {
                    const Raw XA = load_bin<MODE>(p, in, L1 / 2 + L1 * a, pf);
                    const Raw XB = load_bin<MODE>(p, in, L1 / 2 + L1 * (R1 - 1 - a), pf);
                    merge(ldg2(c.twp + L1 / 2 + L1 * a), XA, XB, vB[a], vB[R1 - 1 - a]);
                }
            }
            Dft<R1, true, f2>::run(vA);
            Dft<R1, true, f2>::run(vB);
            store_group(c, l, jA, vA);
            store_group(c, l, jB, vB);
        }
    }
    // ================= TMA path (one-sided layouts): the S tile has LANDED over the work buffer ==========
    // NH = TF / 8 column blocks; block h holds columns [8h, 8h + 8) of the tile as [Mc rows][8 slices][re, im]
    // (row k at k * 64 bytes -- for consecutive units a wavefront reads 128 contiguous bytes), the NH
    // Nyquist rows sit behind the blocks.  Columns outside S arrive as zeros (tensor-map bounds).
    static constexpr int NH = TF / 8;
    static constexpr int LAND_BLOCK = MC * 16;                           // floats per column block
    static constexpr int NYQ_STRIDE = 32;                                // Nyquist rows: 128 bytes apart (TMA alignment)
    static constexpr int LAND_ROWS = MC < 256 ? MC : 256;                // rows per tensor box
    static constexpr unsigned LAND_BYTES = 4u * NH * (MC * 16 + 16);     // bytes one tile brings in
    struct Live { bool l0, l1; };
    VHD static Live make_live(const WInvCtx &c, int l) {
        const WInvArgs &p = c.a;
        const int64_t qa0 = c.qa + GEO::slice_of(l, 0), qa1 = c.qa + GEO::slice_of(l, 1);
        const int64_t col0 = qa0 - p.p_min, col1 = qa1 - p.p_min;
        Live v;
        v.l0 = qa0 >= p.q0 && qa0 < p.q1 && col0 >= 0 && col0 < p.P;
        v.l1 = qa1 >= p.q0 && qa1 < p.q1 && col1 >= 0 && col1 < p.P;
        return v;
    }
    template <int MODE>
    VHD static Raw land_bin(const WInvCtx &c, const Live &lv, int l, int k) {
        Raw v;
        const float *rowp = k < MC ? c.work + 16 * k : c.work + NH * LAND_BLOCK;      // Nyquist rows: [NH][32]
        const int bs = k < MC ? LAND_BLOCK : NYQ_STRIDE;                                // block stride
        if (ADJ) {
            const float *q = rowp + 4 * l;
            WD_TRACE(0, q, 16);
            const fl4 t = *reinterpret_cast<const fl4 *>(q);
            v.r0 = t.x; v.i0 = t.y; v.r1 = t.z; v.i1 = t.w;
        } else {
            const fl2 a = ldsx(rowp + 2 * l), b = ldsx(rowp + bs + 2 * l);
            v.r0 = a.x; v.i0 = a.y; v.r1 = b.x; v.i1 = b.y;
        }
        if (!lv.l0) { v.r0 = 0.f; v.i0 = 0.f; }
        if (!lv.l1) { v.r1 = 0.f; v.i1 = 0.f; }
        if (MODE == 3 && k >= 1 && k <= MC - 1) {
            v.r0 *= c.a.rfac2x; v.i0 *= c.a.rfac2x; v.r1 *= c.a.rfac2x; v.i1 *= c.a.rfac2x;
        }
        return v;
    }
    // loads of unit w (both groups of the pair, + Nyquist for w = 0) from the landed tile; the caller
    // synchronises the CTA before land_rest overwrites the buffer
    struct P1Regs { Raw XA[R1], XB[R1], XN; };
    static constexpr int P1_ITERS = (L1 / 2 + NSUB - 1) / NSUB;
    template <int MODE>
    VHD static void land_load(const WInvCtx &c, int tid, int it, P1Regs &r) {
        const int l = tid % LN, w = tid / LN + it * NSUB;
        if (w >= L1 / 2) return;
        const Live lv = make_live(c, l);
        const int jA = w, jB = w == 0 ? L1 / 2 : L1 - w;
#pragma unroll
        for (int a = 0; a < R1; ++a) {
            r.XA[a] = land_bin<MODE>(c, lv, l, jA + L1 * a);
            r.XB[a] = land_bin<MODE>(c, lv, l, jB + L1 * a);
        }
        r.XN = r.XA[0];
        if (w == 0) r.XN = land_bin<MODE>(c, lv, l, MC);
    }
    VHD static void land_rest(const WInvCtx &c, int tid, int it, const P1Regs &r) {
        const int l = tid % LN, w = tid / LN + it * NSUB;
        if (w >= L1 / 2) return;
        C2 vA[R1], vB[R1];
        int jA = w, jB = L1 - w;
        if (w != 0) {
            const fl2 *tq = c.twp + jA;
#pragma unroll
            for (int a = 0; a < R1; ++a) merge(ldg2(tq + L1 * a), r.XA[a], r.XB[R1 - 1 - a], vA[a], vB[R1 - 1 - a]);
        } else {
            jB = L1 / 2;
            vA[0].x = mk2(r.XA[0].r0 + r.XN.r0, r.XA[0].r1 + r.XN.r1);       // DC, Nyquist: imaginary parts ignored
            vA[0].y = mk2(r.XA[0].r0 - r.XN.r0, r.XA[0].r1 - r.XN.r1);
#pragma unroll
            for (int a = 1; a < R1 / 2; ++a) merge(ldg2(c.twp + L1 * a), r.XA[a], r.XA[R1 - a], vA[a], vA[R1 - a]);
            {
                C2 dump;
                merge(ldg2(c.twp + MC / 2), r.XA[R1 / 2], r.XA[R1 / 2], vA[R1 / 2], dump);
            }
#pragma unroll
            for (int a = 0; a < R1 / 2; ++a)
                merge(ldg2(c.twp + L1 / 2 + L1 * a), r.XB[a], r.XB[R1 - 1 - a], vB[a], vB[R1 - 1 - a]);
        }
        Dft<R1, true, f2>::run(vA);
        Dft<R1, true, f2>::run(vB);
        store_group(c, l, jA, vA);
        store_group(c, l, jB, vB);
    }

    VHD static void store_group(const WInvCtx &c, int l, int j, const C2 (&v)[R1]) {
        float *wre = c.work + GEO::off_re(j) + 2 * l;
        float *wim = c.work + GEO::off_im(j) + 2 * l;
        const fl2 *tw = c.tw1 + j;
        sts2(wre, v[0].x);
        sts2(wim, v[0].y);
#pragma unroll
        for (int cc = 1; cc < R1; ++cc) {
            const fl2 t = ldg2(tw + cc * L1);
            const C2 z = cmulwc(v[cc], bc(t.x), bc(t.y), bc(-t.y));
            sts2(wre + cc * (2 * TF * L1), z.x);
            sts2(wim + cc * (2 * TF * L1), z.y);
        }
    }
    VHD static void pass1(const WInvCtx &c, int tid) {
        const int m = c.a.mode;
        if (m == 3) pass1_t<3>(c, tid);
        else if (m == 2) pass1_t<2>(c, tid);
        else if (m == 1) pass1_t<1>(c, tid);
        else pass1_t<0>(c, tid);
    }
    VHD static void pass2(const WInvCtx &c, int tid) { GEO::template pass2_t<true, R1, R2, L2>(c.work, c.tw2, tid); }

    // ---- pass 3: warp unit = (c2, run of SUBW consecutive cb): the thread's group g = cb + R1 c2 yields
    //      z'[m], m = g + G d, i.e. the time samples 2m, 2m+1 of its two slices.  NIT units per thread.
    static constexpr int CBR = R1 / SUBW, NU3 = R2 * CBR, NIT = (NU3 + NWARP - 1) / NWARP;
    struct Regs { C2 y[NIT][RL]; };
    VHD static void pass3_load(const WInvCtx &c, int tid, Regs &r) {
        const int l = tid % LN, sub = tid / LN, warp = sub / SUBW, sw = sub % SUBW;
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int unit = warp + it * NWARP;
            if (unit < NU3) {
                const int c2 = unit / CBR, cb = (unit - c2 * CBR) * SUBW + sw;
                GEO::template ld_grp<RL, L1, L2>(c.work, cb, c2, l, r.y[it]);
                Dft<RL, true, f2>::run(r.y[it]);
            }
        }
    }
    // ---- overlap-add, quarter Q of every slice (src/lib.rs:499-503 roll back + truncate, :869-910 dual
    //      window and +=).  Sample pair m of slice s lands at tile sample u = s H + j0, j0 = (2m + M/2) mod M;
    //      quarter Q = j0 / H goes to accumulator row s + Q.  Within one quarter pass every cell receives
    //      exactly one contribution; the passes run Q = 3, 2, 1, 0, i.e. oldest slice first: the
    //      reference's increasing-q order, deterministic, no atomics.  The products are rounded on their
    //      own so that the sums do not depend on the tiling.
    template <int Q>
    VHD static void ola(const WInvCtx &c, int tid, const Regs &r) {
        const int l = tid % LN, sub = tid / LN, warp = sub / SUBW, sw = sub % SUBW;
        constexpr int DQ = RL / 4, STEP = M / RL;      // d's per quarter, samples between consecutive d
#pragma unroll
        for (int it = 0; it < NIT; ++it) {
            const int unit = warp + it * NWARP;
            if (unit >= NU3) continue;
            const int c2 = unit / CBR, cb = (unit - c2 * CBR) * SUBW + sw;
            const int g = cb + R1 * c2;
            float *a0 = c.acc + (GEO::slice_of(l, 0) + Q) * HS + 2 * g;
            float *a1 = c.acc + (GEO::slice_of(l, 1) + Q) * HS + 2 * g;
            const fl2 *dq = c.dtab + g;
#pragma unroll
            for (int dd = 0; dd < DQ; ++dd) {
                const int d = (((Q + 2) & 3) * DQ + dd);                     // j0 / H = (d / DQ + 2) mod 4 = Q
                const fl2 dw = ldg2(dq + G * d);
                const C2 y = r.y[it][d];
                fl2 va, vb;
                va.x = mul_rn(lo(y.x), dw.x); va.y = mul_rn(lo(y.y), dw.y);
                vb.x = mul_rn(hi(y.x), dw.x); vb.y = mul_rn(hi(y.y), dw.y);
                float *p0 = a0 + dd * STEP, *p1 = a1 + dd * STEP;
                if (Q == 3) {                                                // oldest contribution of its cell
                    stsx(p0, va);
                    stsx(p1, vb);
                } else {
                    fl2 t0 = ldsx(p0), t1 = ldsx(p1);
                    t0.x = add_rn(t0.x, va.x); t0.y = add_rn(t0.y, va.y);
                    t1.x = add_rn(t1.x, vb.x); t1.y = add_rn(t1.y, vb.y);
                    stsx(p0, t0);
                    stsx(p1, t1);
                }
            }
        }
    }
    VHD static void zero_head(const WInvCtx &c, int tid) {
        for (int i = tid; i < 3 * HS; i += NT) c.acc[i] = 0.f;
    }
    // completed samples u in [0, TF H) -> global, clipped to [k0, k1) and to what this tile owns
    VHD static void out(const WInvCtx &c, int tid) {
        const WInvArgs &p = c.a;
        constexpr int TFH = TF * H;
        const int64_t base = c.qa * (int64_t)H - MID;                       // k = base + u
        float *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        int64_t ka = base + (c.first ? (int64_t)c.halo * H : 0), kb = base + TFH;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        if (kb > c.kend) kb = c.kend;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        if (ub <= ua) return;
        const bool al = p.vec_out && ((((c.b * p.out_pitch - p.k0 + base)) & 3) == 0);
        for (int u = 4 * tid; u < TFH; u += 4 * NT) {
            const int r = u / H;
            const float *q = c.acc + r * HS + (u - r * H);
            WD_TRACE(0, q, 16);
            const fl4 v = *reinterpret_cast<const fl4 *>(q);
            if (al && u >= ua && u + 4 <= ub) {
                *reinterpret_cast<fl4 *>(xo + u) = v;
            } else {
                if (u >= ua && u < ub) xo[u] = v.x;
                if (u + 1 >= ua && u + 1 < ub) xo[u + 1] = v.y;
                if (u + 2 >= ua && u + 2 < ub) xo[u + 2] = v.z;
                if (u + 3 >= ua && u + 3 < ub) xo[u + 3] = v.w;
            }
        }
    }
    // rows TF .. TF+2 (partial sums of the samples the next tile completes) become rows 0 .. 2
    static constexpr int TPT = (3 * H / 4 + NT - 1) / NT;
    struct Tail { fl4 v[TPT]; };
    VHD static void take_tail(const WInvCtx &c, int tid, Tail &t) {
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            const int u = 4 * (tid + i * NT);
            if (u < 3 * H) {
                const int r = u / H;
                t.v[i] = *reinterpret_cast<const fl4 *>(c.acc + (TF + r) * HS + (u - r * H));
            }
        }
    }
    VHD static void keep_tail(const WInvCtx &c, int tid, const Tail &t) {
#pragma unroll
        for (int i = 0; i < TPT; ++i) {
            const int u = 4 * (tid + i * NT);
            if (u < 3 * H) {
                const int r = u / H;
                *reinterpret_cast<fl4 *>(c.acc + r * HS + (u - r * H)) = t.v[i];
            }
        }
    }
};

// ================================================================== host side: tables, geometry
struct WCfg { int logmc, tf; };
inline bool wide_cfg_for(int64_t M, int tf_pref, WCfg *c) {
    if (M == 1024) { c->logmc = 9; c->tf = 16; return true; }
    if (M == 2048) { c->logmc = 10; c->tf = tf_pref == 16 ? 16 : 8; return true; }   // 2 CTAs / SM: 52.7 vs 49.3 %
    if (M == 4096) { c->logmc = 11; c->tf = 8; return true; }
    return false;
}
// default geometry only: m_num = mfft = 4 hop, roll = m_num / 2 (phase_shift = 0)
inline bool wide_geometry_ok(int64_t M, int64_t N, int64_t H, int64_t ps) {
    return N == M && 4 * H == M && ps == M / 2 && (M == 1024 || M == 2048 || M == 4096);
}
template <typename FN>
inline auto wide_dispatch(const WCfg &c, FN &fn) -> decltype(fn.template run<9, 16>()) {
    if (c.logmc == 9 && c.tf == 16) return fn.template run<9, 16>();
    if (c.logmc == 10 && c.tf == 16) return fn.template run<10, 16>();
    if (c.logmc == 10 && c.tf == 8) return fn.template run<10, 8>();
    if (c.logmc == 11 && c.tf == 8) return fn.template run<11, 8>();
    return decltype(fn.template run<9, 16>())(-1);
}
// table layout (fl2 = 2 floats per entry), one set per direction:
//   [xtab: Mc][tw1: R1 L1 = Mc][tw2: R2 L2 = L1][twp: Mc + 1]
// xtab = window * 0.5 (forward, the 1/2 of the real-FFT split) or dual * scale (inverse), permuted so
// that entry m holds the samples j0, j0 + 1 with j0 = (2m + M/2) mod M of FFT slots 2m, 2m + 1.
// (tw2 gets Mc / 8 >= L1 entries: R1 >= 8 in every plan)
inline int64_t wide_table_entries(int64_t M) { const int64_t mc = M / 2; return mc + mc + mc / 8 + (mc + 1) + 1; }
struct WTabOff { int64_t xtab, tw1, tw2, twp; };
inline WTabOff wide_table_offsets(int64_t M) {
    const int64_t mc = M / 2;
    return WTabOff{0, mc, 2 * mc, 2 * mc + mc / 8};
}
inline void wide_table_fill(int64_t M, int r1, int r2, const double *vec_n, double scale, double *out2) {
    const int64_t mc = M / 2, l1 = mc / r1, l2 = l1 / r2;
    const WTabOff o = wide_table_offsets(M);
    for (int64_t i = 0; i < 2 * wide_table_entries(M); ++i) out2[i] = 0.0;
    for (int64_t m = 0; m < mc; ++m) {
        const int64_t j0 = (2 * m + M / 2) % M;
        out2[2 * (o.xtab + m)] = vec_n[j0] * scale;
        out2[2 * (o.xtab + m) + 1] = vec_n[j0 + 1] * scale;
    }
    for (int64_t cc = 0; cc < r1; ++cc)
        for (int64_t j = 0; j < l1; ++j) tw_exact(j * cc, mc, &out2[2 * (o.tw1 + cc * l1 + j)], &out2[2 * (o.tw1 + cc * l1 + j) + 1]);
    for (int64_t cc = 0; cc < r2; ++cc)
        for (int64_t j = 0; j < l2; ++j) tw_exact(j * cc, l1, &out2[2 * (o.tw2 + cc * l2 + j)], &out2[2 * (o.tw2 + cc * l2 + j) + 1]);
    for (int64_t k = 0; k <= mc; ++k) tw_exact(k, M, &out2[2 * (o.twp + k)], &out2[2 * (o.twp + k) + 1]);
}
// shared memory: [work (+ the landed Nyquist rows)][staged signal | accumulator][plan table][mbarrier]
template <int LOGMC, int TF>
inline size_t wide_work_floats() {
    using G = WGeo<LOGMC, TF>;
    return (size_t)G::WORK_FLOATS + 64;               // + NH Nyquist rows of the landed tile (inverse, TMA path)
}
template <int LOGMC, int TF>
inline size_t wide_smem_bytes() {
    using G = WGeo<LOGMC, TF>;
    return 4 * (wide_work_floats<LOGMC, TF>() + (size_t)G::XIN_FLOATS) + 8 * (size_t)wide_table_entries(G::M) + 16;
}
template <int LOGMC, int TF> inline size_t wide_fwd_smem_bytes() { return wide_smem_bytes<LOGMC, TF>(); }
template <int LOGMC, int TF> inline size_t wide_inv_smem_bytes() { return wide_smem_bytes<LOGMC, TF>(); }

}  // namespace wd
}  // namespace veils
