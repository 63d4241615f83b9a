// smooth_impl.cuh -- the mixed-radix ("smooth") kernel family: any mfft = product of up to four radices
// 2 .. 16 (every 13-smooth length up to the shared-memory limit), f32 and f64, all layouts, any window
// length / hop / roll.  It takes the non-power-of-two lengths that used to run the O(M^2) dft_generic
// kernels (e.g. Hann 400 / hop 160, or the reference's own 15-point golden case).
//
//   * one CTA = a tile of TF consecutive slices; lane = slice, the rest of the thread index walks the
//     butterfly units: work[element][slice] is interleaved complex, so every shared-memory access of a
//     warp row is one contiguous run (conflict free), twiddles are row-uniform;
//   * full-length complex transform by in-place decimation: pass p works inside blocks of B_p = M / (R_1
//     .. R_{p-1}) elements on the R_p elements S_p = B_p / R_p apart, multiplies by W_{B_p}^(j c) and
//     leaves bin k = c_1 + R_1 c_2 + R_1 R_2 c_3 + ... at position c_1 L_1 + c_2 L_2 + ... (L_p = M / (R_1
//     .. R_p)): digit arithmetic instead of a re-ordering pass;
//   * forward: load phase (frame extraction, zero padding, roll, window -- src/lib.rs:336-345, 612-652,
//     715-729), passes, store phase (mode epilogue :354-408, [f][p] order :735-746);
//   * inverse: load phase (mode prologue :414-487), passes, windowed rows (:499-503, 869-873), and an
//     owner-computes overlap-add gather in increasing q (:853-910) -- deterministic, no atomics.
// Phase functions compile for the device (smooth.cu) and for the host (tests/emu/smooth_emu.cu).
#pragma once
#include "smooth_dft.cuh"

namespace veils {
namespace sm {

// per-launch description of the transform
struct Shape {
    int M, N, H, F, ps, np;
    int R[kMaxPasses];       // radices
    int L[kMaxPasses];       // L_p = M / (R_1 .. R_p)
    int twoff[kMaxPasses];   // offset of pass p's twiddles W_{B_p}^(j c) at [c S_p + j] in the table
};
inline bool make_shape(int64_t M, int64_t N, int64_t H, int64_t F, int64_t ps, Shape *s) {
    Factors f;
    if (M > 32768 || !factorize(M, &f)) return false;
    s->M = (int)M; s->N = (int)N; s->H = (int)H; s->F = (int)F; s->ps = (int)ps; s->np = f.np;
    int l = (int)M, off = 0;
    for (int p = 0; p < kMaxPasses; ++p) {
        s->R[p] = f.R[p];
        s->twoff[p] = off;
        off += l;                                       // B_p entries
        l /= f.R[p];
        s->L[p] = l;
    }
    return true;
}
inline int64_t twiddle_entries(const Shape &s) {
    int64_t n = 0, b = s.M;
    for (int p = 0; p < s.np; ++p) { n += b; b /= s.R[p]; }
    return n;
}
// table of complex doubles: pass p at twoff[p]: entry [c S_p + j] = exp(-2 pi i j c / B_p)
inline void twiddle_fill(const Shape &s, double *re_im) {
    int64_t b = s.M;
    for (int p = 0; p < s.np; ++p) {
        const int64_t S = b / s.R[p];
        for (int64_t c = 0; c < s.R[p]; ++c)
            for (int64_t j = 0; j < S; ++j) p2::tw_exact(j * c, b, &re_im[2 * (s.twoff[p] + c * S + j)], &re_im[2 * (s.twoff[p] + c * S + j) + 1]);
        b = S;
    }
}
// position of bin k after the passes
VHD int bin_pos(const Shape &s, int k) {
    int pos = 0;
#pragma unroll
    for (int p = 0; p < kMaxPasses; ++p) {
        if (p < s.np) {
            const int c = k % s.R[p];
            k /= s.R[p];
            pos += c * s.L[p];
        }
    }
    return pos;
}

// one pass of radix R over the whole tile
template <typename T, int TF, int R, bool INV>
VHD void pass_r(cx<T> *work, const cx<T> *tw, int nblocks, int B, int tid, int nt) {
    const int lane = tid % TF, sub = tid / TF, nsub = nt / TF;
    const int S = B / R, units = nblocks * S;
    for (int u = sub; u < units; u += nsub) {
        const int blk = u / S, j = u - blk * S;
        cx<T> *p = work + (size_t)(blk * B + j) * TF + lane;
        cx<T> v[R];
#pragma unroll
        for (int a = 0; a < R; ++a) v[a] = p[(size_t)a * S * TF];
        Dft<R, INV, T>::run(v);
        p[0] = v[0];
        if (S == 1) {
#pragma unroll
            for (int c = 1; c < R; ++c) p[(size_t)c * TF] = v[c];
        } else {
#pragma unroll
            for (int c = 1; c < R; ++c) {
                const cx<T> w = tw[c * S + j];
                p[(size_t)c * S * TF] = INV ? p2::cmulconj(v[c], w) : p2::cmul(v[c], w);
            }
        }
    }
}
template <typename T, int TF, bool INV>
VHD void pass_any(int R, cx<T> *work, const cx<T> *tw, int nblocks, int B, int tid, int nt) {
    switch (R) {
        case 2: pass_r<T, TF, 2, INV>(work, tw, nblocks, B, tid, nt); break;
        case 3: pass_r<T, TF, 3, INV>(work, tw, nblocks, B, tid, nt); break;
        case 4: pass_r<T, TF, 4, INV>(work, tw, nblocks, B, tid, nt); break;
        case 5: pass_r<T, TF, 5, INV>(work, tw, nblocks, B, tid, nt); break;
        case 6: pass_r<T, TF, 6, INV>(work, tw, nblocks, B, tid, nt); break;
        case 7: pass_r<T, TF, 7, INV>(work, tw, nblocks, B, tid, nt); break;
        case 8: pass_r<T, TF, 8, INV>(work, tw, nblocks, B, tid, nt); break;
        case 9: pass_r<T, TF, 9, INV>(work, tw, nblocks, B, tid, nt); break;
        case 10: pass_r<T, TF, 10, INV>(work, tw, nblocks, B, tid, nt); break;
        case 11: pass_r<T, TF, 11, INV>(work, tw, nblocks, B, tid, nt); break;
        case 12: pass_r<T, TF, 12, INV>(work, tw, nblocks, B, tid, nt); break;
        case 13: pass_r<T, TF, 13, INV>(work, tw, nblocks, B, tid, nt); break;
        case 14: pass_r<T, TF, 14, INV>(work, tw, nblocks, B, tid, nt); break;
        case 15: pass_r<T, TF, 15, INV>(work, tw, nblocks, B, tid, nt); break;
        case 16: pass_r<T, TF, 16, INV>(work, tw, nblocks, B, tid, nt); break;
        default: break;
    }
}
// pass p of the plan (the caller synchronises the CTA between passes)
template <typename T, int TF, bool INV>
VHD void run_pass(const Shape &s, int p, cx<T> *work, const cx<T> *tw, int tid, int nt) {
    const int B = p == 0 ? s.M : s.L[p - 1];
    pass_any<T, TF, INV>(s.R[p], work, tw + s.twoff[p], s.M / B, B, tid, nt);
}

// ================================================================== FORWARD
template <typename T>
struct FwdArgs {
    const T *x;
    void *out;
    const T *win;            // [N]
    const cx<T> *tw;
    int64_t n, x_pitch, ldp, p0, P, k_offset, ptiles;
    int mid, mode, spectrogram, xpad;   // xpad: staged sample u lives at u + (u / H) xpad
    T fac2x;
    Shape s;
};
template <typename T>
struct FwdCtx {
    const FwdArgs<T> *a;      // the kernel's __grid_constant__ parameter (constant bank: no local copy)
    T *xin;
    cx<T> *work;
    int64_t b, pt;
};

template <typename T, int TF>
struct Fwd {
    VHD static int span(const Shape &s) { return (TF - 1) * s.H + s.N; }
    VHD static int xaddr(const FwdArgs<T> &a, int u) { return u + (u / a.s.H) * a.xpad; }
    // samples of the tile, zero outside [0, n) (src/lib.rs:631-645)
    VHD static void stage(const FwdCtx<T> &c, int tid, int nt) {
        const FwdArgs<T> &p = *c.a;
        const int64_t s0 = (p.p0 + c.pt * TF) * (int64_t)p.s.H + p.k_offset - p.mid;
        const T *xb = p.x + c.b * p.x_pitch;
        const int sp = span(p.s);
        for (int u = tid; u < sp; u += nt) {
            const int64_t g = s0 + u;
            c.xin[xaddr(p, u)] = (g >= 0 && g < p.n) ? p2::ldg(xb + g) : (T)0;
        }
    }
    // FFT slot i of slice s holds z[(i + ps) mod M]: the windowed sample, or zero padding (src/lib.rs:336-345)
    VHD static void load(const FwdCtx<T> &c, int tid, int nt) {
        const FwdArgs<T> &p = *c.a;
        const int lane = tid % TF, sub = tid / TF, nsub = nt / TF;
        for (int i = sub; i < p.s.M; i += nsub) {
            int j = i + p.s.ps;
            if (j >= p.s.M) j -= p.s.M;
            T v = 0;
            if (j < p.s.N) v = c.xin[xaddr(p, lane * p.s.H + j)] * p2::ldg(p.win + j);
            c.work[(size_t)i * TF + lane] = mk<T>(v, (T)0);
        }
    }
    VHD static void pass(const FwdCtx<T> &c, int pi, int tid, int nt) {
        run_pass<T, TF, false>(c.a->s, pi, c.work, c.a->tw, tid, nt);
    }
    // bins -> S rows with the mode epilogue (src/lib.rs:354-408): row r of the stored layout takes bin k(r)
    VHD static void store(const FwdCtx<T> &c, int tid, int nt) {
        const FwdArgs<T> &p = *c.a;
        const int lane = tid % TF, sub = tid / TF, nsub = nt / TF;
        const int64_t col = c.pt * TF + lane;
        if (col >= p.P) return;
        const int M = p.s.M, F = p.s.F;
        for (int r = sub; r < F; r += nsub) {
            int k = r;
            if (p.mode == 1) { k = r + (M + 1) / 2; if (k >= M) k -= M; }       // fftshift: S[r] = X[(r + ceil(M/2)) mod M]
            cx<T> v = c.work[(size_t)bin_pos(p.s, k) * TF + lane];
            if (p.mode >= 2) {
                const bool nyq = (M % 2 == 0) && k == M / 2;
                if (k == 0 || nyq) v.y = 0;                                       // :378-384
                else if (p.mode == 3) { v.x *= p.fac2x; v.y *= p.fac2x; }         // :397-403
            }
            if (p.spectrogram) {
                T *o = reinterpret_cast<T *>(p.out) + ((c.b * F + r) * p.ldp + col);
                *o = v.x * v.x + v.y * v.y;
            } else {
                cx<T> *o = reinterpret_cast<cx<T> *>(p.out) + ((c.b * F + r) * p.ldp + col);
                *o = v;
            }
        }
    }
};

// ================================================================== INVERSE
template <typename T>
struct InvArgs {
    const void *S;
    T *xo;
    const T *dual;           // [N] dual window / M (one-sided layouts) or / (2 M) (two-sided: Hermitian part)
    const cx<T> *tw;
    int64_t P, ldp, out_pitch, p_min, k0, k1, q0, q1, qa0, ntiles;
    int mid, mode, halo, ostride;
    T rfac2x;
    Shape s;
};
template <typename T>
struct InvCtx {
    const InvArgs<T> *a;
    cx<T> *work;
    T *ola;                  // [TF][ostride] windowed slices of the tile
    int64_t b, qa;
};

template <typename T, int TF>
struct Inv {
    VHD static cx<T> ldS(const InvArgs<T> &p, int64_t b, int row, int64_t col) {
        const T *q = reinterpret_cast<const T *>(p.S) + 2 * ((b * p.s.F + row) * p.ldp + col);
        return mk<T>(p2::ldg(q), p2::ldg(q + 1));
    }
    // full spectrum X[i], i in [0, M), of slice q from the stored layout (src/lib.rs:414-487); only the real
    // part of the inverse transform is kept (:903-907), so two-sided layouts enter Hermitian-symmetrised
    VHD static void load(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int lane = tid % TF, sub = tid / TF, nsub = nt / TF;
        const int64_t q = c.qa + lane, col = q - p.p_min;
        const bool live = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;        // :855-857
        const int M = p.s.M;
        for (int i = sub; i < M; i += nsub) {
            cx<T> v = mk<T>((T)0, (T)0);
            if (live) {
                if (p.mode >= 2) {
                    const int k = 2 * i <= M ? i : M - i;
                    v = ldS(p, c.b, k, col);
                    const bool nyq = (M % 2 == 0) && k == M / 2;
                    if (k == 0 || nyq) v.y = 0;                                    // irfft semantics
                    else if (p.mode == 3) { v.x *= p.rfac2x; v.y *= p.rfac2x; }
                    if (k != i) v.y = -v.y;                                        // X[M - k] = conj(X[k])
                } else {
                    // stored row of bin i: twosided r = i; centered r = (i - ceil(M/2)) mod M = (i + floor(M/2)) mod M
                    int ra = p.mode == 1 ? i + M / 2 : i;
                    if (ra >= M) ra -= M;
                    const int im = i == 0 ? 0 : M - i;
                    int rb = p.mode == 1 ? im + M / 2 : im;
                    if (rb >= M) rb -= M;
                    const cx<T> a = ldS(p, c.b, ra, col), bq = ldS(p, c.b, rb, col);
                    v = mk<T>(a.x + bq.x, a.y - bq.y);      // twice the Hermitian part: the 1/2 is folded into the dual table
                }
            }
            c.work[(size_t)i * TF + lane] = v;
        }
    }
    VHD static void pass(const InvCtx<T> &c, int pi, int tid, int nt) {
        run_pass<T, TF, true>(c.a->s, pi, c.work, c.a->tw, tid, nt);
    }
    // xs[j] = Re(y[(j - ps) mod M]) * dual[j] / M, j < N  (src/lib.rs:493-503, 869-873)
    VHD static void rows(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int lane = tid % TF, sub = tid / TF, nsub = nt / TF;
        const int M = p.s.M;
        for (int j = sub; j < p.s.N; j += nsub) {
            int i = j - p.s.ps;
            if (i < 0) i += M;
            const T y = c.work[(size_t)bin_pos(p.s, i) * TF + lane].x;
            c.ola[(size_t)lane * p.ostride + j] = y * p2::ldg(p.dual + j);
        }
    }
    // overlap-add as a gather: the tile owns the samples whose LATEST contributing slice is one of its
    // slices halo .. TF-1; sums run over increasing q (src/lib.rs:853)
    VHD static void gather(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int H = p.s.H, N = p.s.N;
        const int64_t base = c.qa * (int64_t)H - p.mid;                   // k = base + u
        int64_t ka = base + (int64_t)p.halo * H, kb = base + (int64_t)TF * H;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        T *xo = p.xo + c.b * p.out_pitch - p.k0;
        for (int64_t k = ka + tid; k < kb; k += nt) {
            const int u = (int)(k - base);
            const int rl = u / H;                                          // latest slice (tile index)
            int r = rl - p.halo;
            if (r < 0) r = 0;
            T acc = 0;
            for (; r <= rl; ++r) {
                const int j = u - r * H;
                if (j < N) acc += c.ola[(size_t)r * p.ostride + j];
            }
            xo[k] = acc;
        }
    }
};

}  // namespace sm
}  // namespace veils
