// pow2_tile_impl.cuh -- the power-of-two "slice-lane" kernel family, written as phase functions
// that compile both for the device (pow2_tile.cu) and for a host emulation used by the CPU tests
// (tests/emu/pow2_emu.cu), so the index algebra can be checked without a GPU.
//
// Design:
//   * One CTA owns a tile of TF consecutive time slices of one signal.  Thread t = (sub, f):
//     f = t % TF is the slice ("lane <-> slice"), sub = t / TF selects the butterfly group.
//     Because all threads of a row share (group, element) and differ only in the slice,
//       - every shared-memory access of a row is to TF consecutive words  -> conflict free,
//       - twiddles and window samples are uniform across a row            -> broadcast loads,
//       - the final store of bin k for TF consecutive slices is contiguous in the reference's
//         [f_pts][P] layout (src/lib.rs:735-746)                          -> coalesced, no
//         shared-memory transpose of the output.
//   * The real input (window * signal, src/lib.rs:715-729) is transformed with a half-length
//     complex FFT (Mc = mfft/2): z[m] = x'[2m] + i x'[2m+1]; the split into the mfft/2+1 one-sided
//     bins happens in registers in the last pass because a thread holds the groups (g, G-g)
//     whose bins are each other's mirror images.
//   * In-place decimation: pass 1 reads straight from the staged signal tile (frame extraction,
//     zero padding, windowing and the phase-shift roll of src/lib.rs:336-345, 612-652 are pure
//     load-index arithmetic), middle pass in place, last pass from shared memory to global
//     memory (mode epilogue src/lib.rs:354-408).
//   * The inverse mirrors this: pass 1 loads S columns (coalesced over slices) and undoes the
//     mode (src/lib.rs:414-487), the last pass multiplies by the dual window (:869-873) and the
//     overlap-add (:876-910) is a deterministic gather in increasing q owned by this CTA.
#pragma once
#include "f32x2.cuh"
#include "kernels.cuh"

namespace veils {
namespace p2 {

// ------------------------------------------------------------------ complex helpers
template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};
template <typename T> VHD cx<T> mk(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename T> VHD cx<T> operator+(cx<T> a, cx<T> b) { return mk<T>(a.x + b.x, a.y + b.y); }
template <typename T> VHD cx<T> operator-(cx<T> a, cx<T> b) { return mk<T>(a.x - b.x, a.y - b.y); }
template <typename T> VHD cx<T> cmul(cx<T> a, cx<T> b) {
    return mk<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <typename T> VHD cx<T> cmulconj(cx<T> a, cx<T> w) {      // a * conj(w)
    return mk<T>(a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y);
}
template <typename T> VHD cx<T> conj(cx<T> a) { return mk<T>(a.x, -a.y); }
// multiply by -i (forward) / +i (inverse)
template <bool INV, typename T> VHD cx<T> rot(cx<T> a) {
    return INV ? mk<T>(-a.y, a.x) : mk<T>(a.y, -a.x);
}
// multiply by the constant (c, -s) forward, (c, +s) inverse; T may be float, double or f2
template <bool INV, typename T> VHD cx<T> cmulc(cx<T> a, double c, double s) {
    const T cc = vconst<T>(c), ss = vconst<T>(INV ? -s : s), ns = vconst<T>(INV ? s : -s);
    return mk<T>(vfma(a.y, ss, a.x * cc), vfma(a.x, ns, a.y * cc));   // w = cc - i ss
}
// multiply by a table twiddle given as (wr, wi, -wi):  a * w   and   a * conj(w)
template <typename T> VHD cx<T> cmulw(cx<T> a, T wr, T wi, T nwi) {
    return mk<T>(vfma(a.y, nwi, a.x * wr), vfma(a.y, wr, a.x * wi));
}
template <typename T> VHD cx<T> cmulwc(cx<T> a, T wr, T wi, T nwi) {
    return mk<T>(vfma(a.y, wi, a.x * wr), vfma(a.x, nwi, a.y * wr));
}

template <typename T> VHD T ldg(const T *p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
template <typename T> VHD cx<T> ldgc(const T *p) {   // complex at p[0], p[1] (aligned)
#if defined(__CUDA_ARCH__)
    if (sizeof(T) == 4) {
        const float2 v = __ldg((const float2 *)p);
        return mk<T>((T)v.x, (T)v.y);
    } else {
        const double2 v = __ldg((const double2 *)p);
        return mk<T>((T)v.x, (T)v.y);
    }
#else
    return mk<T>(p[0], p[1]);
#endif
}
VHD unsigned mulhi_u32(unsigned a, unsigned b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (unsigned)(((unsigned long long)a * b) >> 32);
#endif
}

// Asynchronous global -> shared copy of BYTES (4, 8 or 16) with zero fill when !valid.
// Device: cp.async (LDGSTS), completed by cp_async_wait_all(); host emulation: synchronous.
template <int BYTES> VHD void cp_async_zfill(void *sdst, const void *gsrc, bool valid) {
#if defined(__CUDA_ARCH__)
    const unsigned d = (unsigned)__cvta_generic_to_shared(sdst);
    const int sz = valid ? BYTES : 0;
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
#else
    for (int i = 0; i < BYTES; ++i) ((char *)sdst)[i] = valid ? ((const char *)gsrc)[i] : 0;
#endif
}
VHD void cp_async_commit() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
VHD void cp_async_wait_all() {
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// ------------------------------------------------------------------ in-register DFTs
// Natural-order in, natural-order out.  INV selects exp(+i...) (unnormalised).
template <bool INV, typename T> VHD void dft2(cx<T> &a, cx<T> &b) {
    const cx<T> t = a - b;
    a = a + b;
    b = t;
}
template <bool INV, typename T> VHD void dft4(cx<T> &a0, cx<T> &a1, cx<T> &a2, cx<T> &a3) {
    const cx<T> t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3;
    // t3 = (-/+ i)(a1 - a3) without a negation
    const cx<T> t3 = INV ? mk<T>(a3.y - a1.y, a1.x - a3.x) : mk<T>(a1.y - a3.y, a3.x - a1.x);
    a0 = t0 + t2;
    a2 = t0 - t2;
    a1 = t1 + t3;
    a3 = t1 - t3;
}
template <int R, bool INV, typename T> struct Dft;
template <bool INV, typename T> struct Dft<2, INV, T> {
    VHD static void run(cx<T> (&v)[2]) { dft2<INV>(v[0], v[1]); }
};
template <bool INV, typename T> struct Dft<4, INV, T> {
    VHD static void run(cx<T> (&v)[4]) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV, typename T> struct Dft<8, INV, T> {
    // n = 2 n1 + n2, k = k1 + 4 k2
    VHD static void run(cx<T> (&v)[8]) {
        const double h = 0.70710678118654752440;
        dft4<INV>(v[0], v[2], v[4], v[6]);                // n2 = 0 -> t0[k1]
        dft4<INV>(v[1], v[3], v[5], v[7]);                // n2 = 1 -> t1[k1]
        v[3] = cmulc<INV>(v[3], h, h);                    // W8^1
        v[5] = rot<INV>(v[5]);                            // W8^2
        v[7] = cmulc<INV>(v[7], -h, h);                   // W8^3
        cx<T> o[8];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            o[k1] = v[2 * k1] + v[2 * k1 + 1];
            o[k1 + 4] = v[2 * k1] - v[2 * k1 + 1];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = o[i];
    }
};
template <bool INV, typename T> struct Dft<16, INV, T> {
    // n = 4 n1 + n2, k = k1 + 4 k2
    VHD static void run(cx<T> (&v)[16]) {
        const double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
        const double h = 0.70710678118654752440;
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) dft4<INV>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
        // now v[4 k1 + n2] = t[n2][k1]; twiddle W16^(n2 k1)
        v[4 + 1] = cmulc<INV>(v[4 + 1], c1, s1);          // W^1
        v[4 + 2] = cmulc<INV>(v[4 + 2], h, h);            // W^2
        v[4 + 3] = cmulc<INV>(v[4 + 3], s1, c1);          // W^3
        v[8 + 1] = cmulc<INV>(v[8 + 1], h, h);            // W^2
        v[8 + 2] = rot<INV>(v[8 + 2]);                    // W^4
        v[8 + 3] = cmulc<INV>(v[8 + 3], -h, h);           // W^6
        v[12 + 1] = cmulc<INV>(v[12 + 1], s1, c1);        // W^3
        v[12 + 2] = cmulc<INV>(v[12 + 2], -h, h);         // W^6
        v[12 + 3] = cmulc<INV>(v[12 + 3], -c1, -s1);      // W^9
        cx<T> o[16];
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            cx<T> a0 = v[4 * k1], a1 = v[4 * k1 + 1], a2 = v[4 * k1 + 2], a3 = v[4 * k1 + 3];
            dft4<INV>(a0, a1, a2, a3);                    // over n2 -> k2
            o[k1] = a0;
            o[k1 + 4] = a1;
            o[k1 + 8] = a2;
            o[k1 + 12] = a3;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = o[i];
    }
};

// ------------------------------------------------------------------ factorisation of Mc = mfft/2
template <int LOGMC> struct Fac;
template <> struct Fac<4>  { static constexpr int NP = 2, R1 = 4,  R2 = 4,  R3 = 1; };
template <> struct Fac<5>  { static constexpr int NP = 2, R1 = 8,  R2 = 4,  R3 = 1; };
template <> struct Fac<6>  { static constexpr int NP = 2, R1 = 8,  R2 = 8,  R3 = 1; };
template <> struct Fac<7>  { static constexpr int NP = 2, R1 = 16, R2 = 8,  R3 = 1; };
template <> struct Fac<8>  { static constexpr int NP = 2, R1 = 16, R2 = 16, R3 = 1; };
template <> struct Fac<9>  { static constexpr int NP = 3, R1 = 8,  R2 = 8,  R3 = 8; };
template <> struct Fac<10> { static constexpr int NP = 3, R1 = 16, R2 = 8,  R3 = 8; };
template <> struct Fac<11> { static constexpr int NP = 3, R1 = 16, R2 = 16, R3 = 8; };
template <> struct Fac<12> { static constexpr int NP = 3, R1 = 16, R2 = 16, R3 = 16; };

// radices / table offsets for a runtime log2(Mc) (host side: twiddle table layout)
inline void fac_of(int logmc, int *np, int *r1, int *r2, int *r3) {
    static const int tab[13][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0},
                                   {2, 4, 4, 1}, {2, 8, 4, 1}, {2, 8, 8, 1}, {2, 16, 8, 1},
                                   {2, 16, 16, 1}, {3, 8, 8, 8}, {3, 16, 8, 8}, {3, 16, 16, 8},
                                   {3, 16, 16, 16}};
    *np = tab[logmc][0]; *r1 = tab[logmc][1]; *r2 = tab[logmc][2]; *r3 = tab[logmc][3];
}

template <typename T, int LOGMC, int TF_, int NSUB_>
struct Geo {
    using F = Fac<LOGMC>;
    static constexpr int MC = 1 << LOGMC, M = 2 * MC;
    static constexpr int TF = TF_, NSUB = NSUB_, NT = TF_ * NSUB_;
    static constexpr int NP = F::NP, R1 = F::R1, R2 = F::R2, R3 = F::R3;
    static constexpr int RL = (NP == 2) ? R2 : R3;        // radix of the last pass
    static constexpr int L1 = MC / R1;                    // groups of pass 1
    static constexpr int L2 = (NP == 3) ? L1 / R2 : 1;    // groups per block of the middle pass
    static constexpr int G = MC / RL;                     // groups of the last pass
    // twiddle table (complex T, interleaved): [tw1: MC][tw2: L1][twp: MC + 1]
    static constexpr int TW1 = 0, TW2 = 2 * MC, TWP = 2 * (MC + L1);
    // Narrow tiles (a row of TF slices is shorter than 128 bytes: the 8192 / 4096-point f64 and f32 plans): the
    // warp rows of the last pass are whole pass-1 blocks apart, a power-of-two stride that lands every row on
    // the same banks (ncu, c5: 16 instead of 4 wavefronts per LDS.128).  One extra row per pass-1 block skews
    // them over the banks.
    // Measured on c5 (profiles/r2o_ncu_full_c5_padded.txt): conflicts 57 % -> 6 % of the shared wavefronts, but the
    // kernel time does not move (2.13 -> 2.18 ms: padded rows are no tensor-map source, so the stores leave through
    // the LSU, and the tile of two slices is barrier- / latency-bound with 8 warps) -- kept behind a switch.
#ifndef VEILS_P2_PADR
#define VEILS_P2_PADR 0
#endif
    static constexpr int PADR = (VEILS_P2_PADR && NP == 3 && TF * (int)sizeof(cx<T>) < 128) ? 1 : 0;
    static constexpr int L1P = L1 + PADR;                 // pitch (rows) of a pass-1 output block in the work buffer
    static constexpr int MCP = MC + R1 * PADR;            // rows of the work buffer
    // work-memory offset (in complex elements) of last-pass group g: bins k = g + G d
    VHD static int goff(int g) { return (NP == 2) ? g * L1 : (g % R1) * L1P + (g / R1) * L2; }
    // small transforms: persistent CTAs keep the plan tables in shared memory and double-buffer
    // the staged signal tile (cp.async); large ones read the tables through the read-only path.
    static constexpr bool STAB = LOGMC <= 8;
    // f64 mfft 512: one staged-signal buffer keeps the CTA under half an SM's shared memory -- two resident CTAs
    // are worth more than the prefetch (measured: forward 39.4 -> 55.0 % of the HBM roofline)
#ifndef VEILS_F64_L8_DBUF
#define VEILS_F64_L8_DBUF 0
#endif
    static constexpr bool DBUF = LOGMC <= 8 && !(sizeof(T) == 8 && LOGMC == 8 && !VEILS_F64_L8_DBUF);
    static constexpr int TWN = 2 * (MC + L1 + MC + 1);    // reals in the twiddle table
    VHD static cx<T> tlc(const T *q) {                     // complex table entry
        if (STAB) return *reinterpret_cast<const cx<T> *>(q);
        return ldgc(q);
    }
    VHD static T tl(const T *q) {
        if (STAB) return *q;
        return ldg(q);
    }
};

// Shared-memory layout of the staged signal: sample i = r * H + c of the tile lives at
// r * (H + padw) + c, so that the slice stride H + padw keeps lane accesses conflict free.
struct TileLayout {
    int H, padw;
    unsigned magic;           // ceil(2^32 / H): j / H == umulhi(j, magic) for j, H < 2^16
    VHD int row(int j) const { return (int)mulhi_u32((unsigned)j, magic); }
    VHD int addr(int j) const { return j + row(j) * padw; }      // j < 2^16
};

// ================================================================== FORWARD
template <typename T>
struct FwdArgs {
    const T *x;
    void *out;
    const T *win;            // [N] window * 0.5 (the 1/2 of the real-FFT split, exact)
    const T *tw;
    int64_t n, x_pitch, ldp, p0, P, k_offset, ptiles;
    int N, H, mid, ps, F;
    TileLayout lay;
    unsigned magic_hp;       // ceil(2^32 / (H/2)) (paired staging)
    int vec;                 // ps, H, N even -> paired shared-memory loads
    int padded;              // N < M
    int mode;                // Mode
    int shift;               // centered: Mc (fftshift), else 0
    int spectrogram;         // 1: |X|^2 real output
    int dbg;                 // ablation mask (VEILS_ABLATE builds only)
    int fastgeo;             // m_num = mfft = 4 hop and roll = m_num / 2 (the default geometry)
    T fac2x;                 // onesided2X factor, 1 otherwise
};

template <typename T>
struct FwdCtx {
    FwdArgs<T> a;
    const T *tw, *win;       // plan tables: shared-memory copies (STAB) or the global tables
    T *xin;                  // staged signal tile (shared)
    cx<T> *work;             // [MC][TF] (shared)
    int64_t b, pt;           // signal, tile index
};

template <typename T, int LOGMC, int TF_, int NSUB_>
struct Fwd : Geo<T, LOGMC, TF_, NSUB_> {
    using GEO = Geo<T, LOGMC, TF_, NSUB_>;
    using C = cx<T>;
    using GEO::MC; using GEO::M; using GEO::TF; using GEO::NSUB; using GEO::NT; using GEO::NP;
    using GEO::R1; using GEO::R2; using GEO::R3; using GEO::RL; using GEO::L1; using GEO::L2;
    using GEO::G;

    // ---- phase 0: stage the tile's samples, zero outside [0, n)  (src/lib.rs:631-645)
    // Asynchronous (cp.async with zero fill); the caller commits and waits.  Global index of tile
    // sample i is s0 + i; its shared address is (i / H) * (H + padw) + i % H.
    VHD static void fill(const FwdCtx<T> &c, int tid) {
        const FwdArgs<T> &p = c.a;
        const int H = p.H;
        const int span = (TF - 1) * H + p.N;
        const int hs = H + p.lay.padw;
        const int64_t s0 = (p.p0 + c.pt * TF) * (int64_t)H + p.k_offset - p.mid;
        const T *xb = p.x + c.b * p.x_pitch;
        if (p.vec && (((c.b * p.x_pitch + s0) & 1) == 0) && ((p.n & 1) == 0) &&
            ((((size_t)p.x) & (2 * sizeof(T) - 1)) == 0)) {
            // pairs: never straddle the signal borders (s0, n even) nor a tile row (H even)
            const int np = span >> 1, hp = H >> 1;
            for (int e = tid; e < np; e += NT) {
                const int64_t s = s0 + 2 * (int64_t)e;
                const bool ok = s >= 0 && s < p.n;
                const int r = (int)mulhi_u32((unsigned)e, p.magic_hp);         // e / (H/2)
                cp_async_zfill<2 * (int)sizeof(T)>(c.xin + r * hs + 2 * (e - r * hp), ok ? xb + s : xb, ok);
            }
            return;
        }
        for (int e = tid; e < span; e += NT) {
            const int64_t s = s0 + e;
            const bool ok = s >= 0 && s < p.n;
            const int r = e / H;
            cp_async_zfill<(int)sizeof(T)>(c.xin + r * hs + (e - r * H), ok ? xb + s : xb, ok);
        }
    }

    // windowed, rolled, zero-padded sample pair x'[2m], x'[2m+1] of the slice whose tile base is xf
    template <bool VEC, bool PADDED>
    VHD static C load_pair(const FwdArgs<T> &p, const T *win, const T *xf, int m) {
        const int j0 = (2 * m + p.ps) & (M - 1);
        if (VEC) {
            if (PADDED && j0 >= p.N) return mk<T>(0, 0);
            const C xv = *reinterpret_cast<const C *>(xf + p.lay.addr(j0));
            const C w = GEO::tlc(win + j0);
            return mk<T>(xv.x * w.x, xv.y * w.y);
        }
        const int j1 = (j0 + 1) & (M - 1);
        T v0 = 0, v1 = 0;
        if (j0 < p.N) v0 = xf[p.lay.addr(j0)] * GEO::tl(win + j0);
        if (j1 < p.N) v1 = xf[p.lay.addr(j1)] * GEO::tl(win + j1);
        return mk<T>(v0, v1);
    }

    // ---- pass 1: groups j in [0, L1): DFT_R1 over z[j + L1 a], twiddle W_Mc^(j c)
    template <bool VEC, bool PADDED>
    VHD static void pass1_t(const FwdCtx<T> &c, int tid) {
        const FwdArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        const T *xf = c.xin + f * (p.H + p.lay.padw);
        for (int j = sub; j < L1; j += NSUB) {
            C v[R1];
#pragma unroll
            for (int a = 0; a < R1; ++a) v[a] = load_pair<VEC, PADDED>(p, c.win, xf, j + L1 * a);
            Dft<R1, false, T>::run(v);
            C *wj = c.work + j * TF + f;
            const T *twj = c.tw + GEO::TW1 + 2 * j * R1;
            wj[0] = v[0];
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) wj[GEO::L1P * cc * TF] = cmul(v[cc], GEO::tlc(twj + 2 * cc));
        }
    }
    VHD static void pass1(const FwdCtx<T> &c, int tid) {
        if (c.a.vec) {
            if (c.a.padded) pass1_t<true, true>(c, tid);
            else pass1_t<true, false>(c, tid);
        } else {
            pass1_t<false, true>(c, tid);
        }
    }

    // ---- middle pass (3-pass plans): block cb, groups j2 in [0, L2): DFT_R2, twiddle W_L1^(j2 c2)
    VHD static void pass_mid(const FwdCtx<T> &c, int tid) {
        if (NP != 3) return;
        const FwdArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        for (int u = sub; u < R1 * L2; u += NSUB) {
            const int cb = u / L2, j2 = u % L2;
            C *wb = c.work + (cb * GEO::L1P + j2) * TF + f;
            const T *twj = c.tw + GEO::TW2 + 2 * j2 * R2;
            C v[R2];
#pragma unroll
            for (int a = 0; a < R2; ++a) v[a] = wb[L2 * a * TF];
            Dft<R2, false, T>::run(v);
            wb[0] = v[0];
#pragma unroll
            for (int cc = 1; cc < R2; ++cc) wb[L2 * cc * TF] = cmul(v[cc], GEO::tlc(twj + 2 * cc));
        }
    }

    // store one-sided bin k (0 <= k <= Mc) of this thread's slice column.
    // ONES: one-sided layouts (modes 2, 3); SPEC: |X|^2 real output.  ob: byte pointer of
    // (row 0, this column); rb: row pitch in bytes.
    // byte offset of one-sided bin k in the STAGED tile (one-sided layouts, TMA store path): the RL
    // rows of last-pass group g = k mod G sit on the group's own work rows; bin Mc has an extra row
    // narrow tiles, |X|^2 staging: the real cells fill only half of the group's rows, so every other run of
    // groups starts one cell row later -- together with the padded block pitch the half-warp's rows fall on
    // distinct banks (ncu, c5: 16 wavefronts per STS.64 before)
    VHD static int sskew(int g) {
        constexpr int CW = TF * (int)sizeof(T);
        return GEO::PADR ? ((g / (64 / CW)) & 1) * CW : 0;
    }
    template <bool SPEC>
    VHD static int64_t srow(int k) {
        if (k >= MC) return (int64_t)GEO::MCP * TF * (int64_t)sizeof(C);
        return (int64_t)GEO::goff(k & (G - 1)) * TF * (int64_t)sizeof(C) + (SPEC ? sskew(k & (G - 1)) : 0) +
               (int64_t)(k / G) * TF * (int64_t)(SPEC ? sizeof(T) : sizeof(C));
    }
    template <bool ONES, bool SPEC, bool STAGE = false>
    VHD static void store_bin(const FwdArgs<T> &p, char *ob, int64_t rb, int k, C X) {
        if (k == 0 || k == MC) X.y = (T)0;                  // src/lib.rs:378-384
        if (ONES) {
            if (k >= 1 && k <= MC - 1) {                    // src/lib.rs:397-403 (fac2x == 1 for mode 2)
                X.x *= p.fac2x;
                X.y *= p.fac2x;
            }
            char *q = STAGE ? ob + srow<SPEC>(k) : ob + k * rb;
            if (SPEC) *reinterpret_cast<T *>(q) = X.x * X.x + X.y * X.y;
            else *reinterpret_cast<C *>(q) = X;
            return;
        }
        // two-sided layouts: bin k and its mirror M - k (x is real); fftshift for centered
        const int kk = (k + p.shift) & (M - 1);
        const int km = (M - k + p.shift) & (M - 1);
        const bool mir = !(k == 0 || k == MC);
        if (SPEC) {
            const T v = X.x * X.x + X.y * X.y;
            *reinterpret_cast<T *>(ob + kk * rb) = v;
            if (mir) *reinterpret_cast<T *>(ob + km * rb) = v;
        } else {
            *reinterpret_cast<C *>(ob + kk * rb) = X;
            if (mir) *reinterpret_cast<C *>(ob + km * rb) = conj(X);
        }
    }

    // split of the half-length spectrum: A = Z[k]/2, Zp = Z[Mc-k]/2  ->  X[k], X[Mc-k]
    template <bool ONES, bool SPEC, bool STAGE = false>
    VHD static void emit(const FwdArgs<T> &p, const T *twp, char *ob, int64_t rb, int k, C A, C Zp, bool both) {
        const C B = conj(Zp);
        const C E = A + B, Om = A - B;
        const C w = GEO::tlc(twp + 2 * k);                   // W_M^k
        const C D = cmul(mk<T>(w.y, -w.x), Om);             // (-i W) (A - B)
        store_bin<ONES, SPEC, STAGE>(p, ob, rb, k, E + D);
        if (both) store_bin<ONES, SPEC, STAGE>(p, ob, rb, MC - k, conj(E - D));
    }

    // ---- last pass: groups (g, G-g): DFT_RL, real-FFT split, mode epilogue, global store
    // STAGE: the finished bins overwrite the cells (row, slice) this thread has read (one-sided
    // layouts); pair(s) is called after every pair of groups so that the caller can hand the two
    // dense [RL rows][TF slices] boxes to the TMA engine (pow2_tile.cu)
    struct NoHook { VHD void operator()(int) const {} };
    template <bool ONES, bool SPEC, bool STAGE = false, typename HOOK = NoHook>
    VHD static void pass_last_t(const FwdCtx<T> &c, int tid, HOOK hook = HOOK()) {
        const FwdArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        const int64_t pl = c.pt * TF + f;
        const bool live = STAGE || pl < p.P;
        const int64_t esz = SPEC ? sizeof(T) : sizeof(C);
        const int64_t rb = p.ldp * esz;
        char *ob = STAGE ? reinterpret_cast<char *>(c.work) + f * esz
                         : reinterpret_cast<char *>(p.out) + ((c.b * p.F) * p.ldp + pl) * esz;
        const T *twp = c.tw + GEO::TWP;
        for (int s = sub; s < (G / 2 + NSUB - 1) / NSUB * NSUB; s += NSUB) {
            if (s >= G / 2) {                  // STAGE: every thread of the warp reaches the hook
                if (STAGE) hook(-1);
                continue;
            }
            const int g = s, gp = (s == 0) ? G / 2 : G - s;
            C u[RL], up[RL];
            const C *w1 = c.work + GEO::goff(g) * TF + f, *w2 = c.work + GEO::goff(gp) * TF + f;
#pragma unroll
            for (int d = 0; d < RL; ++d) {
                u[d] = w1[d * TF];
                up[d] = w2[d * TF];
            }
            Dft<RL, false, T>::run(u);
            Dft<RL, false, T>::run(up);
            if (!live) continue;
#if defined(__CUDA_ARCH__)
            if (STAGE && SPEC) __syncwarp();   // |X|^2 cells are half as wide: lanes overwrite each other's inputs
#endif
            if (s != 0) {
#pragma unroll
                for (int d = 0; d < RL; ++d)
                    emit<ONES, SPEC, STAGE>(p, twp, ob, rb, g + G * d, u[d], up[RL - 1 - d], true);
            } else {
                emit<ONES, SPEC, STAGE>(p, twp, ob, rb, 0, u[0], u[0], true);                  // X[0], X[Mc]
#pragma unroll
                for (int d = 1; d < RL / 2; ++d) emit<ONES, SPEC, STAGE>(p, twp, ob, rb, G * d, u[d], u[RL - d], true);
                emit<ONES, SPEC, STAGE>(p, twp, ob, rb, MC / 2, u[RL / 2], u[RL / 2], false);  // self-mirrored
#pragma unroll
                for (int d = 0; d < RL / 2; ++d)
                    emit<ONES, SPEC, STAGE>(p, twp, ob, rb, G / 2 + G * d, up[d], up[RL - 1 - d], true);
            }
            if (STAGE) hook(s);
        }
    }
    // ---- last pass, one group per thread (f64): keeping both mirrored groups in one thread costs 2 x RL complex
    // doubles (212 registers, 8 warps per SM).  Here every thread transforms ONE group and writes its Z back over
    // the rows it read (last_a); after a CTA barrier it reads the RL mirror values Z[Mc - k] from the partner
    // group's rows and splits only its own bins (last_x) -- half the registers, twice the threads; the price is one
    // shared-memory round trip of Z and 5 more flops per bin.  last_put stores the bins (STAGE: over the thread's
    // own rows, behind a second barrier, for the TMA boxes).
    // Measured on B200: NOT a win -- c2 geometry in f64 54.4 % (paired: 55.1 %), c1 forward 74 us (60), c5 3.2 ms
    // (2.1): the second pass over shared memory, two more barriers and the spills at 128 registers cost more than
    // the doubled warps bring.  Kept behind a switch (the host emulation follows the same switch).
#ifndef VEILS_F64_FWD_UNPAIRED
#define VEILS_F64_FWD_UNPAIRED 0
#endif
    static constexpr bool UNPAIRED = VEILS_F64_FWD_UNPAIRED && sizeof(T) == 8;
    static constexpr int GPTU = (G + NSUB - 1) / NSUB;       // groups per thread
    VHD static void last_a(const FwdCtx<T> &c, int tid) {
        const int f = tid % TF, sub = tid / TF;
#pragma unroll 1
        for (int g = sub; g < G; g += NSUB) {
            C u[RL];
            C *w = c.work + GEO::goff(g) * TF + f;
#pragma unroll
            for (int d = 0; d < RL; ++d) u[d] = w[d * TF];
            Dft<RL, false, T>::run(u);
#pragma unroll
            for (int d = 0; d < RL; ++d) w[d * TF] = u[d];
        }
    }
    // X[k], k = g + G d, of this thread's groups; xn = X[Mc] (held by the thread of group 0)
    VHD static void last_x(const FwdCtx<T> &c, int tid, C (&X)[GPTU][RL], C &xn) {
        const int f = tid % TF, sub = tid / TF;
        const T *twp = c.tw + GEO::TWP;
#pragma unroll
        for (int i = 0; i < GPTU; ++i) {
            const int g = sub + NSUB * i;
            if (g >= G) continue;
            const int gp = g == 0 ? 0 : G - g;
            const C *w = c.work + GEO::goff(g) * TF + f, *wp = c.work + GEO::goff(gp) * TF + f;
#pragma unroll
            for (int d = 0; d < RL; ++d) {
                const int dp = g == 0 ? (RL - d) % RL : RL - 1 - d;      // Mc - (g + G d) = gp + G dp
                const C A = w[d * TF], Zp = wp[dp * TF];
                const C B = conj(Zp);
                const C E = A + B, Om = A - B;
                const C tw = GEO::tlc(twp + 2 * (g + G * d));             // W_M^k
                const C D = cmul(mk<T>(tw.y, -tw.x), Om);               // (-i W) (A - B)
                X[i][d] = E + D;
                if (g == 0 && d == 0) xn = conj(E - D);
            }
        }
    }
    template <bool ONES, bool SPEC, bool STAGE, typename HOOK>
    VHD static void last_put(const FwdCtx<T> &c, int tid, const C (&X)[GPTU][RL], C xn, HOOK hook) {
        const FwdArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        const int64_t pl = c.pt * TF + f;
        const bool live = STAGE || pl < p.P;
        const int64_t esz = SPEC ? sizeof(T) : sizeof(C);
        const int64_t rb = p.ldp * esz;
        char *ob = STAGE ? reinterpret_cast<char *>(c.work) + f * esz
                         : reinterpret_cast<char *>(p.out) + ((c.b * p.F) * p.ldp + pl) * esz;
#pragma unroll
        for (int i = 0; i < GPTU; ++i) {
            const int g = sub + NSUB * i;
            if (g >= G) {
                if (STAGE) hook(-1);
                continue;
            }
            if (live) {
#pragma unroll
                for (int d = 0; d < RL; ++d) store_bin<ONES, SPEC, STAGE>(p, ob, rb, g + G * d, X[i][d]);
                if (g == 0) store_bin<ONES, SPEC, STAGE>(p, ob, rb, MC, xn);
            }
            if (STAGE) hook(g);
        }
    }
    template <bool ONES, bool SPEC>
    VHD static void last_b_t(const FwdCtx<T> &c, int tid) {
        C X[GPTU][RL], xn = mk<T>(0, 0);
        last_x(c, tid, X, xn);
        last_put<ONES, SPEC, false>(c, tid, X, xn, NoHook());
    }
    // host emulation / plain-store device path: last_a for every thread, barrier, last_b for every thread
    VHD static void last_b(const FwdCtx<T> &c, int tid) {
        const int spectrogram = c.a.spectrogram;
        if (c.a.mode >= 2) {
            if (spectrogram) last_b_t<true, true>(c, tid);
            else last_b_t<true, false>(c, tid);
        } else {
            if (spectrogram) last_b_t<false, true>(c, tid);
            else last_b_t<false, false>(c, tid);
        }
    }
    VHD static void pass_last(const FwdCtx<T> &c, int tid) {
        const int spectrogram = c.a.spectrogram;
        if (c.a.mode >= 2) {
            if (spectrogram) pass_last_t<true, true>(c, tid);
            else pass_last_t<true, false>(c, tid);
        } else {
            if (spectrogram) pass_last_t<false, true>(c, tid);
            else pass_last_t<false, false>(c, tid);
        }
    }
};

// ================================================================== INVERSE
template <typename T>
struct InvArgs {
    const void *S;
    T *xo;
    const T *dual;           // [N] dual window / M (two-sided modes: / 2M), exact scaling
    const T *tw;
    int64_t P, ldp, out_pitch, p_min, k0, k1, q0, q1, qa0, ntiles;
    int N, H, mid, ps, F;
    int mode, shift;         // shift: centered Mc (ifftshift), else 0
    T rfac2x;                // 1 / fac2x for onesided2X, 1 otherwise
    unsigned magic_h;        // ceil(2^32 / H)
    int halo;                // ceil(N/H) - 1 leading slices recomputed for the overlap
    int ostride;             // row stride of the overlap buffer
    int pair;                // packed family: rolled-back sample pairs are adjacent (ps, N even)
    // packed family, streaming overlap-add: a signal's owned slices [qh0, qh1] are cut into chunks of
    // chunk_tiles consecutive tiles that one CTA walks in order, carrying the overlap tail
    int64_t qh0, qh1, cps;   // first / last owned (latest contributing) slice, chunks per signal
    int chunk_tiles;
    int dbg;                 // ablation mask (VEILS_ABLATE builds only)
};

template <typename T>
struct InvCtx {
    InvArgs<T> a;
    const T *tw, *dual;      // plan tables: shared-memory copies (STAB) or the global tables
    cx<T> *work;             // [MC][TF]  (shared)
    T *ola;                  // [TF][ostride] -- ALIASES work
    int64_t b, qa;           // signal, first slice computed by this tile
    // streaming overlap-add (tile family, round 2): a CTA walks the tiles of one chunk of one signal in order
    const T *carry_in;       // [N - H] partial sums (earlier tiles' slices) of the samples that follow the previous tile
    T *carry_out;            // [N - H] ... of the samples that follow this tile
    int first, last;         // first tile of its chunk: no carry in; last tile: no carry out
    int64_t kown;            // first sample the chunk owns (its leading `halo` slices are re-computed for it)
};

template <typename T, int LOGMC, int TF_, int NSUB_>
struct Inv : Geo<T, LOGMC, TF_, NSUB_> {
    using GEO = Geo<T, LOGMC, TF_, NSUB_>;
    using C = cx<T>;
    using GEO::MC; using GEO::M; using GEO::TF; using GEO::NSUB; using GEO::NT; using GEO::NP;
    using GEO::R1; using GEO::R2; using GEO::R3; using GEO::RL; using GEO::L1; using GEO::L2;
    using GEO::G;
    static constexpr int GPT = (G + NSUB - 1) / NSUB;       // last-pass groups per thread

    // Hermitian one-sided bin k (0 <= k <= Mc) of this thread's column, mode undone
    // (src/lib.rs:414-487); scaled by 2 for the two-sided layouts (folded into the dual window).
    // sb: byte pointer of (row 0, this column); rb: row pitch in bytes.
    template <bool ONES>
    VHD static C load_bin(const InvArgs<T> &p, const char *sb, int64_t rb, int k) {
        C v;
        if (ONES) {
            v = *reinterpret_cast<const C *>(sb + k * rb);
            if (k >= 1 && k <= MC - 1) {                    // src/lib.rs:454-470 (rfac2x == 1 for mode 2)
                v.x *= p.rfac2x;
                v.y *= p.rfac2x;
            }
        } else {
            const int k1 = (k + p.shift) & (M - 1);         // X[k] + conj(X[M-k]); ifftshift :516-521
            const int k2 = (M - k + p.shift) & (M - 1);
            const C a = *reinterpret_cast<const C *>(sb + k1 * rb);
            const C bq = *reinterpret_cast<const C *>(sb + k2 * rb);
            v = a + conj(bq);
        }
        if (k == 0 || k == MC) v.y = (T)0;                  // irfft ignores Im(DC), Im(Nyquist)
        return v;
    }

    // Z'[k] = 2 Z[k], Z'[Mc-k] from X[k], X[Mc-k]
    VHD static void merge(const T *twp, int k, C &a, C &bm) {
        const C Xk = a, Xm = conj(bm);
        const C E = Xk + Xm, Dm = Xk - Xm;
        const C O = cmulconj(Dm, GEO::tlc(twp + 2 * k));     // conj(W_M^k) (Xk - conj Xm)
        const C iO = mk<T>(-O.y, O.x);
        a = E + iO;
        bm = conj(E - iO);
    }

    // ---- pass 1: load S columns (coalesced over slices), undo mode, real-IFFT merge, IDFT_R1
    template <bool ONES>
    VHD static void pass1_t(const InvCtx<T> &c, int tid) {
        const InvArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        const int64_t q = c.qa + f;
        const int64_t col = q - p.p_min;
        const bool live = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;
        const int64_t rb = p.ldp * (int64_t)sizeof(C);
        const char *sb = reinterpret_cast<const char *>(p.S) + ((c.b * p.F) * p.ldp + col) * (int64_t)sizeof(C);
        const T *twp0 = c.tw + GEO::TWP;
        for (int s = sub; s < L1 / 2; s += NSUB) {
            const int j = s, jp = (s == 0) ? L1 / 2 : L1 - s;
            C v[R1], vp[R1];
            if (live) {
#pragma unroll
                for (int a = 0; a < R1; ++a) {
                    v[a] = load_bin<ONES>(p, sb, rb, j + L1 * a);
                    vp[a] = load_bin<ONES>(p, sb, rb, jp + L1 * a);
                }
                if (s != 0) {
#pragma unroll
                    for (int a = 0; a < R1; ++a) merge(twp0, j + L1 * a, v[a], vp[R1 - 1 - a]);
                } else {
                    const C xn = load_bin<ONES>(p, sb, rb, MC);   // Nyquist pairs with DC
                    v[0] = mk<T>(v[0].x + xn.x, v[0].x - xn.x);
#pragma unroll
                    for (int a = 1; a < R1 / 2; ++a) merge(twp0, L1 * a, v[a], v[R1 - a]);
                    {
                        C t = v[R1 / 2];
                        merge(twp0, MC / 2, v[R1 / 2], t);
                    }
#pragma unroll
                    for (int a = 0; a < R1 / 2; ++a) merge(twp0, jp + L1 * a, vp[a], vp[R1 - 1 - a]);
                }
            } else {
#pragma unroll
                for (int a = 0; a < R1; ++a) v[a] = vp[a] = mk<T>(0, 0);
            }
            Dft<R1, true, T>::run(v);
            Dft<R1, true, T>::run(vp);
            C *wj = c.work + j * TF + f, *wp = c.work + jp * TF + f;
            const T *twj = c.tw + GEO::TW1 + 2 * j * R1, *twq = c.tw + GEO::TW1 + 2 * jp * R1;
            wj[0] = v[0];
            wp[0] = vp[0];
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) {
                wj[GEO::L1P * cc * TF] = cmulconj(v[cc], GEO::tlc(twj + 2 * cc));
                wp[GEO::L1P * cc * TF] = cmulconj(vp[cc], GEO::tlc(twq + 2 * cc));
            }
        }
    }
    // One group per thread (f64): Z'[k] = E + i O needs X[k] and its mirror X[Mc - k]; instead of keeping the
    // mirrored group in the same thread (2 x R1 complex doubles: 255 registers, 8 warps per SM) every thread
    // loads the R1 mirror bins as well (they are the other threads' own bins: L1 / L2 hits) and merges only its
    // own group -- half the registers, twice the resident warps, one code path for every group (k = 0 pairs with
    // the Nyquist bin and k = Mc / 2 with itself through the same formula).
#ifndef VEILS_P1U_BATCH
#define VEILS_P1U_BATCH 8
#endif
    VHD static C merge_one(const T *twp, int k, C Xk, C Xp) {
        const C Xm = conj(Xp);
        const C E = Xk + Xm, Dm = Xk - Xm;
        const C O = cmulconj(Dm, GEO::tlc(twp + 2 * k));     // conj(W_M^k) (Xk - conj X[Mc-k])
        return mk<T>(E.x - O.y, E.y + O.x);                  // E + i O
    }
    template <bool ONES>
    VHD static void pass1_u(const InvCtx<T> &c, int tid) {
        const InvArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        const int64_t q = c.qa + f;
        const int64_t col = q - p.p_min;
        const bool live = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;
        const int64_t rb = p.ldp * (int64_t)sizeof(C);
        const char *sb = reinterpret_cast<const char *>(p.S) + ((c.b * p.F) * p.ldp + (live ? col : 0)) * (int64_t)sizeof(C);
        const T *twp0 = c.tw + GEO::TWP;
        for (int j = sub; j < L1; j += NSUB) {
            C v[R1];
            if (live) {
#if VEILS_P1U_BATCH
                // batches of VEILS_P1U_BATCH own + mirror bins in flight, merged before the next loads are issued
                constexpr int HB = R1 >= VEILS_P1U_BATCH ? VEILS_P1U_BATCH : R1;
#pragma unroll
                for (int a0 = 0; a0 < R1; a0 += HB) {
                    C w[HB];
#pragma unroll
                    for (int a = 0; a < HB; ++a) {
                        v[a0 + a] = load_bin<ONES>(p, sb, rb, j + L1 * (a0 + a));
                        w[a] = load_bin<ONES>(p, sb, rb, MC - (j + L1 * (a0 + a)));
                    }
#pragma unroll
                    for (int a = 0; a < HB; ++a) v[a0 + a] = merge_one(twp0, j + L1 * (a0 + a), v[a0 + a], w[a]);
                }
#else
                C w[R1];
#pragma unroll
                for (int a = 0; a < R1; ++a) {
                    v[a] = load_bin<ONES>(p, sb, rb, j + L1 * a);
                    w[a] = load_bin<ONES>(p, sb, rb, MC - (j + L1 * a));
                }
#pragma unroll
                for (int a = 0; a < R1; ++a) v[a] = merge_one(twp0, j + L1 * a, v[a], w[a]);
#endif
            } else {
#pragma unroll
                for (int a = 0; a < R1; ++a) v[a] = mk<T>(0, 0);
            }
            Dft<R1, true, T>::run(v);
            C *wj = c.work + j * TF + f;
            const T *twj = c.tw + GEO::TW1 + 2 * j * R1;
            wj[0] = v[0];
#pragma unroll
            for (int cc = 1; cc < R1; ++cc) wj[GEO::L1P * cc * TF] = cmulconj(v[cc], GEO::tlc(twj + 2 * cc));
        }
    }
    static constexpr bool UNPAIRED = sizeof(T) == 8;
    VHD static void pass1(const InvCtx<T> &c, int tid) {
        if (UNPAIRED) {
            if (c.a.mode >= 2) pass1_u<true>(c, tid);
            else pass1_u<false>(c, tid);
            return;
        }
        if (c.a.mode >= 2) pass1_t<true>(c, tid);
        else pass1_t<false>(c, tid);
    }

    VHD static void pass_mid(const InvCtx<T> &c, int tid) {
        if (NP != 3) return;
        const InvArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        for (int u = sub; u < R1 * L2; u += NSUB) {
            const int cb = u / L2, j2 = u % L2;
            C *wb = c.work + (cb * GEO::L1P + j2) * TF + f;
            const T *twj = c.tw + GEO::TW2 + 2 * j2 * R2;
            C v[R2];
#pragma unroll
            for (int a = 0; a < R2; ++a) v[a] = wb[L2 * a * TF];
            Dft<R2, true, T>::run(v);
            wb[0] = v[0];
#pragma unroll
            for (int cc = 1; cc < R2; ++cc) wb[L2 * cc * TF] = cmulconj(v[cc], GEO::tlc(twj + 2 * cc));
        }
    }

    // ---- last pass, part A: IDFT_RL into registers (work is about to be reused as `ola`)
    VHD static void last_load(const InvCtx<T> &c, int tid, C (&r)[GPT][RL]) {
        const int f = tid % TF, sub = tid / TF;
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = sub + NSUB * i;
            if (g < G) {
                const C *w = c.work + GEO::goff(g) * TF + f;
#pragma unroll
                for (int d = 0; d < RL; ++d) r[i][d] = w[d * TF];
                Dft<RL, true, T>::run(r[i]);
            }
        }
    }
    // ---- last pass, part B: roll back (src/lib.rs:499-500), truncate to N, dual window
    VHD static void last_store(const InvCtx<T> &c, int tid, const C (&r)[GPT][RL]) {
        const InvArgs<T> &p = c.a;
        const int f = tid % TF, sub = tid / TF;
        const int ps = p.ps, N = p.N;
        T *row = c.ola + (size_t)f * p.ostride;
        if (N == M && (ps & 1) == 0) {
            // no truncation, even roll: the pair (y[2m], y[2m+1]) lands on adjacent row samples j0, j0 + 1 (j0 even)
            // and its dual-window factors are one aligned pair of the table
#pragma unroll
            for (int i = 0; i < GPT; ++i) {
                const int g = sub + NSUB * i;
                if (g < G) {
                    int j0 = (2 * g + ps) & (M - 1);
#pragma unroll
                    for (int d = 0; d < RL; ++d) {
                        const C w = GEO::tlc(c.dual + j0);
                        row[j0] = r[i][d].x * w.x;
                        row[j0 + 1] = r[i][d].y * w.y;
                        j0 = (j0 + 2 * G) & (M - 1);
                    }
                }
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < GPT; ++i) {
            const int g = sub + NSUB * i;
            if (g < G) {
#pragma unroll
                for (int d = 0; d < RL; ++d) {
                    const int m = g + G * d;                // z[m] = (y[2m], y[2m+1])
                    const int j0 = (2 * m + ps) & (M - 1);
                    const int j1 = (j0 + 1) & (M - 1);
                    if (j0 < N) row[j0] = r[i][d].x * GEO::tl(c.dual + j0);
                    if (j1 < N) row[j1] = r[i][d].y * GEO::tl(c.dual + j1);
                }
            }
        }
    }

    // ---- overlap-add as a gather over the samples this tile owns, increasing q
    template <int HALO>
    VHD static void gather_t(const InvCtx<T> &c, int tid) {
        const InvArgs<T> &p = c.a;
        const int H = p.H, N = p.N, halo = HALO >= 0 ? HALO : p.halo;
        const int64_t qo = c.qa + halo;                      // first owned "latest slice"
        int64_t ka = qo * H - p.mid, kb = (c.qa + TF) * (int64_t)H - p.mid;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        // local sample index u = k + mid - qa * H = (r + halo) * H + c0, r = first contributing slice
        const int64_t base = c.qa * (int64_t)H - p.mid;      // k = base + u
        T *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const int cnt = (TF - halo) * H, hH = halo * H, step = p.ostride - H;
        for (int e = tid; e < cnt; e += NT) {
            const int u = e + hH;
            if (u < ua || u >= ub) continue;
            const int r = (int)mulhi_u32((unsigned)e, p.magic_h);   // e / H
            const int jj = e - r * H + hH;                   // offset inside slice r (the oldest contributor)
            const T *q = c.ola + r * p.ostride + jj;
            // increasing q (src/lib.rs:853); only the oldest slice can fall beyond the window length (halo H < N)
            T acc = jj < N ? q[0] : (T)0;
            if (HALO >= 0) {
#pragma unroll
                for (int h = 1; h <= (HALO >= 0 ? HALO : 0); ++h) acc += q[h * step];
            } else {
                for (int h = 1; h <= halo; ++h) acc += q[h * step];
            }
            xo[u] = acc;
        }
    }
    VHD static void gather(const InvCtx<T> &c, int tid) {
        switch (c.a.halo) {
            case 1: gather_t<1>(c, tid); break;
            case 3: gather_t<3>(c, tid); break;
            default: gather_t<-1>(c, tid); break;
        }
    }

    // ---- STREAMING overlap-add (src/lib.rs:853-911), increasing q: a CTA walks the tiles of one chunk of one signal
    // in order; local sample u = k + mid - qa H.
    //   u in [0, TF H)             : completed by this tile -> global memory; what earlier tiles' slices contributed
    //                                arrives in carry_in (u < N - H).  Only the first tile of a chunk re-computes the
    //                                `halo` slices before its first owned one (kown clips its output).
    //   u in [TF H, TF H + N - H)  : partial sums over the slices so far -> carry_out (a carry that reaches beyond
    //                                the next tile -- tiles narrower than the overlap -- is forwarded).
    // The sums run over earlier tiles' slices first, then this tile's in increasing q: the reference's order whatever
    // the tiling; everything is owned by this CTA: deterministic, no atomics.  Tiles need not be wider than the
    // overlap, which is what lets the two- and four-slice f64 tiles of mfft 8192 / 4096 invert 75 % overlap at all.
    // (a plain loop over the overlap, not unrolled, and nothing else: unrolled / templated forms and a second loop form
    // for long overlaps each measured 8 % slower on the c2 geometry in f64 -- the kernel sits at its 128-register cap)
    VHD static T ola_sum(const InvCtx<T> &c, int u, T acc) {
        const InvArgs<T> &p = c.a;
        const int H = p.H, N = p.N, halo = p.halo;
        const int rq = (int)mulhi_u32((unsigned)u, p.magic_h);           // u / H: latest contributing slice (tile index)
        int r = rq - halo;
        int jj = u - r * H;
        for (int h = 0; h <= halo; ++h, ++r, jj -= H)
            if (r >= 0 && r < TF && jj < N) acc += c.ola[r * p.ostride + jj];
        return acc;
    }
    template <int HALO>
    VHD static void gather_s_t(const InvCtx<T> &c, int tid) {
        const InvArgs<T> &p = c.a;
        const int H = p.H, N = p.N, halo = HALO >= 0 ? HALO : p.halo;
        const int TFH = TF * H, NH = N - H, hH = halo * H, step = p.ostride - H;
        const int64_t base = c.qa * (int64_t)H - p.mid;                  // k = base + u
        T *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        int64_t ka = base, kb = base + TFH;
        if (ka < c.kown) ka = c.kown;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        if (ka > kb) ka = kb;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        // (1) the bulk u in [halo H, TF H): every contributing slice is local
        for (int e = tid; e < TFH - hH; e += NT) {
            const int u = e + hH;
            if (u < ua || u >= ub) continue;
            const int r = (int)mulhi_u32((unsigned)e, p.magic_h);         // e / H
            const int jj = e - r * H + hH;                               // offset inside slice r (the oldest contributor)
            const T *q = c.ola + r * p.ostride + jj;
            T acc = jj < N ? q[0] : (T)0;                                // only the oldest can fall beyond the window length
            if (HALO >= 0) {
#pragma unroll
                for (int h = 1; h <= (HALO >= 0 ? HALO : 0); ++h) acc += q[h * step];
            } else {
                for (int h = 1; h <= halo; ++h) acc += q[h * step];
            }
            xo[u] = acc;
        }
        // (2) the head u in [0, halo H): completed with the previous tiles' carry
        const int hend = hH < TFH ? hH : TFH;
        for (int u = tid; u < hend; u += NT)
            if (u >= ua && u < ub) xo[u] = ola_sum(c, u, (!c.first && u < NH) ? c.carry_in[u] : (T)0);
        // (3) the tail handed on
        if (!c.last)
        for (int v = tid; v < NH; v += NT)
            c.carry_out[v] = ola_sum(c, TFH + v, (!c.first && TFH + v < NH) ? c.carry_in[TFH + v] : (T)0);
    }
    VHD static void gather_s(const InvCtx<T> &c, int tid) {
        switch (c.a.halo) {
            case 1: gather_s_t<1>(c, tid); break;
            case 3: gather_s_t<3>(c, tid); break;
            default: gather_s_t<-1>(c, tid); break;
        }
    }
};

}  // namespace p2
}  // namespace veils
