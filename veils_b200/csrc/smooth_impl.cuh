// smooth_impl.cuh -- the mixed-radix ("smooth") kernel family: any mfft = product of up to four radices
// 2 .. 16 (every 13-smooth length up to the shared-memory limit), f32 and f64, all layouts, any window
// length / hop / roll.  It takes the non-power-of-two lengths that used to run the O(M^2) dft_generic
// kernels (e.g. Hann 400 / hop 160, or the reference's own 15-point golden case).
//
//   * one CTA = a tile of TF consecutive slices; lane = slice, the rest of the thread index walks the
//     butterfly units: work[element][slice] is interleaved complex, so every shared-memory access of a
//     warp row is one contiguous run (conflict free), twiddles and table entries are row-uniform;
//   * even mfft: the real input goes through the HALF-length complex transform z[m] = x'[2m] + i x'[2m+1]
//     (LC = mfft / 2) and a split / merge step with W_M^k; odd mfft: full-length complex transform;
//   * in-place decimation: pass p works inside blocks of B_p = LC / (R_1 .. R_{p-1}) elements on the R_p
//     elements S_p = B_p / R_p apart, multiplies by W_{B_p}^(j c) and leaves index k = c_1 + R_1 c_2 + ...
//     at position c_1 L_1 + c_2 L_2 + ... (L_p = LC / (R_1 .. R_p)): digit arithmetic, no re-ordering pass;
//   * every index computation that is not a butterfly is a PLAN TABLE built once on the host (positions,
//     roll, zero padding, window, dual window, W_M^k): the kernels do no integer division in their loops
//     (the first version spent 60 % of its instructions on them: profiles/r2i_ncu_full_mixed_radix_400.txt);
//   * forward: stage the tile's samples; pass 1 reads them through the load table (frame extraction, zero
//     padding, roll, window -- src/lib.rs:336-345, 612-652, 715-729); store phase = real-FFT split + mode
//     epilogue (:354-408) in the reference's [f][p] order (:735-746);
//   * inverse: load phase (mode prologue :414-487 + Hermitian merge), passes, the last pass multiplies by
//     the dual window and writes the slice rows (:499-503, 869-873), then an owner-computes overlap-add
//     gather in increasing q (:853-910) -- deterministic, no atomics.
// Phase functions compile for the device (smooth.cu) and for the host (tests/emu/smooth_emu.cu).
#pragma once
#include <vector>

#include "smooth_dft.cuh"

namespace veils {
namespace sm {

// per-launch description of the transform
struct Shape {
    int M, LC, half, N, H, F, ps, np;
    int R[kMaxPasses];       // radices of the LC-point transform
    int L[kMaxPasses];       // L_p = LC / (R_1 .. R_p)
    int twoff[kMaxPasses];   // offset of pass p's twiddles W_{B_p}^(j c) at [c S_p + j] in the table
    int units[kMaxPasses];         // butterflies per slice in pass p: LC / R_p
    unsigned magicS[kMaxPasses];   // ceil(2^32 / S_p): v / S_p == umulhi(v, magic) for v, S_p < 2^16
    unsigned magicH;               // ceil(2^32 / H)
};
inline unsigned magic_of(int64_t d) { return d <= 1 ? 0u : (unsigned)((0x100000000ULL + (unsigned long long)d - 1) / (unsigned long long)d); }
// v / d through the magic number (d == 1: magic 0 stands for the identity)
VHD int mdiv(int v, unsigned magic) { return magic ? (int)p2::mulhi_u32((unsigned)v, magic) : v; }
inline bool make_shape(int64_t M, int64_t N, int64_t H, int64_t F, int64_t ps, Shape *s) {
    Factors f;
    if (M < 2 || M > 32768 || H < 1 || H >= 65536) return false;
    bool half = (M % 2 == 0) && M >= 4 && factorize(M / 2, &f);
    if (!half && !factorize(M, &f)) return false;
    s->M = (int)M; s->half = half ? 1 : 0; s->LC = (int)(half ? M / 2 : M);
    s->N = (int)N; s->H = (int)H; s->F = (int)F; s->ps = (int)ps; s->np = f.np;
    int l = s->LC, off = 0;
    for (int p = 0; p < kMaxPasses; ++p) {
        s->R[p] = f.R[p];
        s->twoff[p] = off;
        off += l;                                       // B_p entries
        l /= f.R[p];
        s->L[p] = l;
        s->magicS[p] = magic_of(l);
        s->units[p] = s->LC / f.R[p];
    }
    s->magicH = magic_of(H);
    return true;
}
inline int64_t twiddle_entries(const Shape &s) {
    int64_t n = 0, b = s.LC;
    for (int p = 0; p < s.np; ++p) { n += b; b /= s.R[p]; }
    return n;
}
// table of complex doubles: pass p at twoff[p]: entry [c S_p + j] = exp(-2 pi i j c / B_p)
inline void twiddle_fill(const Shape &s, double *re_im) {
    int64_t b = s.LC;
    for (int p = 0; p < s.np; ++p) {
        const int64_t S = b / s.R[p];
        for (int64_t c = 0; c < s.R[p]; ++c)
            for (int64_t j = 0; j < S; ++j) p2::tw_exact(j * c, b, &re_im[2 * (s.twoff[p] + c * S + j)], &re_im[2 * (s.twoff[p] + c * S + j) + 1]);
        b = S;
    }
}
// position of index k (0 <= k < LC) after the passes
inline int bin_pos(const Shape &s, int k) {
    int pos = 0;
    for (int p = 0; p < s.np; ++p) {
        const int c = k % s.R[p];
        k /= s.R[p];
        pos += c * s.L[p];
    }
    return pos;
}

// ---- plan tables (host-built, one blob per direction) ------------------------------------------------
template <typename T> struct LEnt { int off[2]; T w[2]; };   // pass-1 input e: staged-sample offsets + window
template <typename T> struct SEnt { int pa, pb; T wr, wi; }; // pair (k, LC - k): positions after the passes, W_M^k
template <typename T> struct REnt { int j[2]; T d[2]; };     // position q: slice-row sample indices (-1: none) + dual window
struct TabLayout {
    int tw_off, a_off, b_off, bytes;     // forward: a = LEnt[LC], b = SEnt[LC/2+1] (half) or int pos[M];
                                         // inverse: a = SEnt[LC/2+1] (half, else empty), b = REnt[LC]
};
inline int al16(int64_t v) { return (int)((v + 15) & ~(int64_t)15); }
template <typename T> inline TabLayout tab_layout(const Shape &s, bool inverse) {
    TabLayout l;
    l.tw_off = 0;
    l.a_off = al16((int64_t)sizeof(cx<T>) * twiddle_entries(s));
    if (!inverse) {
        l.b_off = l.a_off + al16((int64_t)sizeof(LEnt<T>) * s.LC);
        l.bytes = l.b_off + al16(s.half ? (int64_t)sizeof(SEnt<T>) * (s.LC / 2 + 1) : (int64_t)sizeof(int) * s.M);
    } else {
        l.b_off = l.a_off + al16(s.half ? (int64_t)sizeof(SEnt<T>) * (s.LC / 2 + 1) : 0);
        l.bytes = l.b_off + al16((int64_t)sizeof(REnt<T>) * s.LC);
    }
    return l;
}
inline int xpad_of(int64_t H) { return (H % 2 == 0) ? 1 : 0; }      // odd slice stride: lanes on distinct banks
// pitch of the inverse's slice rows: odd (lanes = slices on distinct banks); f32 with hop % 4 == 0: = 2 (mod 4), which
// keeps the 16 lanes of a unit on distinct (even) banks AND the rows 8-byte aligned for the four-sample gather
template <typename T> inline int ostride_of(int64_t N, int64_t H) {
    if (sizeof(T) == 4 && H % 4 == 0) { int64_t s = N; while (s % 4 != 2) ++s; return (int)s; }
    return (int)(N | 1);
}
// vec: forward = window [N]; inverse = dual window * 1/M (one-sided) or 1/(2M) (two-sided), [N]
template <typename T> inline void tab_fill(const Shape &s, bool inverse, const double *vec, unsigned char *out) {
    const TabLayout l = tab_layout<T>(s, inverse);
    for (int i = 0; i < l.bytes; ++i) out[i] = 0;
    {
        std::vector<double> tw((size_t)(2 * twiddle_entries(s)));
        twiddle_fill(s, tw.data());
        cx<T> *t = reinterpret_cast<cx<T> *>(out + l.tw_off);
        for (size_t i = 0; i < tw.size() / 2; ++i) t[i] = mk<T>((T)tw[2 * i], (T)tw[2 * i + 1]);
    }
    const int M = s.M, LC = s.LC, N = s.N, H = s.H, xp = xpad_of(s.H);
    SEnt<T> *st = reinterpret_cast<SEnt<T> *>(out + (inverse ? l.a_off : l.b_off));
    if (s.half)
        for (int k = 0; k <= LC / 2; ++k) {
            double wr, wi;
            p2::tw_exact(k, M, &wr, &wi);
            st[k].pa = bin_pos(s, k % LC);
            st[k].pb = bin_pos(s, (LC - k) % LC);
            st[k].wr = (T)wr;
            st[k].wi = (T)wi;
        }
    if (!inverse) {
        LEnt<T> *lt = reinterpret_cast<LEnt<T> *>(out + l.a_off);
        for (int e = 0; e < LC; ++e)
            for (int h = 0; h < 2; ++h) {
                lt[e].off[h] = 0;
                lt[e].w[h] = (T)0;
                if (h == 1 && !s.half) continue;
                const int i = s.half ? 2 * e + h : e;          // FFT slot i holds z[(i + ps) mod M] (src/lib.rs:336-345)
                const int j = (i + s.ps) % M;
                if (j < N) {
                    lt[e].off[h] = j + (j / H) * xp;
                    lt[e].w[h] = (T)(s.half ? 0.5 * vec[j] : vec[j]);   // the 1/2 of the real-FFT split (exact)
                }
            }
        if (!s.half) {
            int *pos = reinterpret_cast<int *>(out + l.b_off);
            for (int k = 0; k < M; ++k) pos[k] = bin_pos(s, k);
        }
    } else {
        REnt<T> *rt = reinterpret_cast<REnt<T> *>(out + l.b_off);
        for (int m = 0; m < LC; ++m) {
            REnt<T> &r = rt[bin_pos(s, m)];
            for (int h = 0; h < 2; ++h) {
                r.j[h] = -1;
                r.d[h] = (T)0;
                if (h == 1 && !s.half) continue;
                const int i = s.half ? 2 * m + h : m;          // y[i] lands on row sample (i + ps) mod M (:499-500)
                const int j = (i + s.ps) % M;
                if (j < N) {
                    r.j[h] = j;
                    r.d[h] = (T)vec[j];
                }
            }
        }
    }
}

// ---- compute vector: f64 keeps one slice per lane; f32 keeps TWO slices (l, l + LN) per lane in an f32x2 register,
// so every butterfly add / mul / fma is one FADD2 / FMUL2 / FFMA2 and all index work is shared by the two slices
template <typename T> struct Vec {
    using V = T;
    static constexpr int W = 1;
    VHD static V bc(T c) { return c; }
    VHD static V pk(T a, T) { return a; }
    VHD static T get(V v, int) { return v; }
    VHD static V neg(V v) { return -v; }
};
template <> struct Vec<float> {
    using V = p2::f2;
    static constexpr int W = 2;
    VHD static V bc(float c) { return p2::bc(c); }
    VHD static V pk(float a, float b) { return p2::mk2(a, b); }
    VHD static float get(V v, int h) { return h ? p2::hi(v) : p2::lo(v); }
    VHD static V neg(V v) { return p2::bc(0.0f) - v; }
};
template <typename T> using vec_t = typename Vec<T>::V;
// a * w / a * conj(w) with a table twiddle broadcast over the slices
template <bool CONJ, typename T> VHD cx<vec_t<T>> mul_tw(cx<vec_t<T>> a, T wr, T wi) {
    using VT = Vec<T>;
    return CONJ ? p2::cmulwc(a, VT::bc(wr), VT::bc(wi), VT::bc(-wi)) : p2::cmulw(a, VT::bc(wr), VT::bc(wi), VT::bc(-wi));
}

// ---- one pass of radix R over the whole tile -------------------------------------------------------------
// KIND 0: in place; 1: forward pass 1, inputs from the staged samples through the load table; 2: inverse last
// pass, outputs through the row table into the slice rows.  work is [LC][LN] of cx<V>, LN = TF / W lanes.
template <typename T> struct PassIO {
    cx<vec_t<T>> *work;
    const T *xin; const LEnt<T> *lt; int hp;     // KIND 1
    T *ola; const REnt<T> *rt; int ostride;      // KIND 2
    int half;
};
template <typename T, int TF, int R, bool INV, int KIND>
VHD void pass_r(const PassIO<T> &io, const cx<T> *tw, int S, unsigned magic, int units, int tid, int nt) {
    using VT = Vec<T>;
    using V = typename VT::V;
    constexpr int W = VT::W, LN = TF / W;
    const int lane = tid % LN, sub = tid / LN, nsub = nt / LN;
    const int B = S * R;
    int blk = mdiv(sub, magic), j = sub - blk * S;
    const int dq = mdiv(nsub, magic), dr = nsub - dq * S;
    for (int u = sub; u < units; u += nsub) {
        const int base = blk * B + j;
        cx<V> v[R];
        if (KIND == 1) {
            const T *xf0 = io.xin + lane * io.hp, *xf1 = xf0 + (W == 2 ? LN * io.hp : 0);
#pragma unroll
            for (int a = 0; a < R; ++a) {
                const LEnt<T> t = io.lt[base + a * S];
                v[a].x = VT::pk(xf0[t.off[0]], xf1[t.off[0]]) * VT::bc(t.w[0]);
                v[a].y = io.half ? VT::pk(xf0[t.off[1]], xf1[t.off[1]]) * VT::bc(t.w[1]) : VT::bc((T)0);
            }
        } else {
            const cx<V> *p = io.work + (size_t)base * LN + lane;
#pragma unroll
            for (int a = 0; a < R; ++a) v[a] = p[(size_t)a * S * LN];
        }
        Dft<R, INV, V>::run(v);
        if (KIND == 2) {                                  // S == 1: position base + c
            T *row = io.ola + (size_t)lane * io.ostride;
#pragma unroll
            for (int c = 0; c < R; ++c) {
                const REnt<T> t = io.rt[base + c];
                const V px = v[c].x * VT::bc(t.d[0]), py = v[c].y * VT::bc(t.d[1]);
#pragma unroll
                for (int h = 0; h < W; ++h) {
                    if (t.j[0] >= 0) row[(size_t)h * LN * io.ostride + t.j[0]] = VT::get(px, h);
                    if (io.half && t.j[1] >= 0) row[(size_t)h * LN * io.ostride + t.j[1]] = VT::get(py, h);
                }
            }
        } else {
            cx<V> *p = io.work + (size_t)base * LN + lane;
            p[0] = v[0];
            if (S == 1) {
#pragma unroll
                for (int c = 1; c < R; ++c) p[(size_t)c * LN] = v[c];
            } else {
#pragma unroll
                for (int c = 1; c < R; ++c) {
                    const cx<T> w = tw[c * S + j];
                    p[(size_t)c * S * LN] = mul_tw<INV, T>(v[c], w.x, w.y);
                }
            }
        }
        j += dr;
        blk += dq;
        if (j >= S) { j -= S; ++blk; }
    }
}
template <typename T, int TF, bool INV, int KIND>
VHD void pass_any(int R, const PassIO<T> &io, const cx<T> *tw, int S, unsigned magic, int units, int tid, int nt) {
    switch (R) {
        case 2: pass_r<T, TF, 2, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 3: pass_r<T, TF, 3, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 4: pass_r<T, TF, 4, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 5: pass_r<T, TF, 5, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 6: pass_r<T, TF, 6, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 7: pass_r<T, TF, 7, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 8: pass_r<T, TF, 8, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 9: pass_r<T, TF, 9, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 10: pass_r<T, TF, 10, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 11: pass_r<T, TF, 11, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 12: pass_r<T, TF, 12, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 13: pass_r<T, TF, 13, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 14: pass_r<T, TF, 14, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 15: pass_r<T, TF, 15, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        case 16: pass_r<T, TF, 16, INV, KIND>(io, tw, S, magic, units, tid, nt); break;
        default: break;
    }
}

// ================================================================== FORWARD
template <typename T>
struct FwdArgs {
    const T *x;
    void *out;
    const unsigned char *tab;  // plan blob (tab_layout<T>(s, false))
    TabLayout lay;
    int64_t n, x_pitch, ldp, p0, P, k_offset, ptiles;
    int mid, mode, spectrogram, xpad;   // xpad: staged sample u lives at u + (u / H) xpad
    T fac2x;
    Shape s;
};
template <typename T>
struct FwdCtx {
    const FwdArgs<T> *a;      // the kernel's __grid_constant__ parameter (constant bank: no local copy)
    const unsigned char *tab; // the blob: shared-memory copy when it fits, else the global one
    T *xin;
    cx<vec_t<T>> *work;       // [LC][LN]
    int64_t b, pt;
};

template <typename T, int TF>
struct Fwd {
    VHD static int span(const Shape &s) { return (TF - 1) * s.H + s.N; }
    // samples of the tile, zero outside [0, n) (src/lib.rs:631-645); sample u = r H + c lives at r (H + xpad) + c.
    // Asynchronous on the device (cp.async with zero fill): the caller commits and waits.
    VHD static void stage(const FwdCtx<T> &c, int tid, int nt) {
        const FwdArgs<T> &p = *c.a;
        const int64_t s0 = (p.p0 + c.pt * TF) * (int64_t)p.s.H + p.k_offset - p.mid;
        const T *xb = p.x + c.b * p.x_pitch;
        const int sp = span(p.s);
        const T *src = xb + s0;
        if (s0 >= 0 && s0 + sp <= p.n) {                 // the whole tile lies inside the signal: no per-sample checks
            for (int u = tid; u < sp; u += nt) {
                const int r = mdiv(u, p.s.magicH);
                p2::cp_async_zfill<(int)sizeof(T)>(c.xin + u + r * p.xpad, src + u, true);
            }
            return;
        }
        for (int u = tid; u < sp; u += nt) {
            const int r = mdiv(u, p.s.magicH);
            const bool ok = s0 + u >= 0 && s0 + u < p.n;
            p2::cp_async_zfill<(int)sizeof(T)>(c.xin + u + r * p.xpad, ok ? src + u : xb, ok);
        }
    }
    VHD static void pass(const FwdCtx<T> &c, int pi, int tid, int nt) {
        const FwdArgs<T> &p = *c.a;
        PassIO<T> io;
        io.work = c.work; io.xin = c.xin; io.hp = p.s.H + p.xpad; io.half = p.s.half;
        io.lt = reinterpret_cast<const LEnt<T> *>(c.tab + p.lay.a_off);
        io.ola = nullptr; io.rt = nullptr; io.ostride = 0;
        const cx<T> *tw = reinterpret_cast<const cx<T> *>(c.tab + p.lay.tw_off) + p.s.twoff[pi];
        const int S = p.s.L[pi], units = p.s.units[pi];
        if (pi == 0) pass_any<T, TF, false, 1>(p.s.R[pi], io, tw, S, p.s.magicS[pi], units, tid, nt);
        else pass_any<T, TF, false, 0>(p.s.R[pi], io, tw, S, p.s.magicS[pi], units, tid, nt);
    }
    // one bin -> its S row(s) with the mode epilogue (src/lib.rs:354-408); k in [0, M/2] for even M (x is real)
    VHD static void emit_real(const FwdArgs<T> &p, int64_t b, int64_t col, int k, cx<T> X) {
        const int M = p.s.M;
        const bool edge = k == 0 || 2 * k == M;
        if (edge) X.y = 0;                                                        // :378-384
        if (p.mode == 3 && !edge) { X.x *= p.fac2x; X.y *= p.fac2x; }             // :397-403
        int r0 = k, r1 = -1;
        if (p.mode < 2) {
            r1 = edge ? -1 : M - k;
            if (p.mode == 1) {                                                    // fftshift: row = (k + floor(M/2)) mod M
                r0 = k + M / 2; if (r0 >= M) r0 -= M;
                if (r1 >= 0) { r1 += M / 2; if (r1 >= M) r1 -= M; }
            }
        }
        if (p.spectrogram) {
            T *o = reinterpret_cast<T *>(p.out) + (b * p.s.F * p.ldp + col);
            const T v = X.x * X.x + X.y * X.y;
            o[(int64_t)r0 * p.ldp] = v;
            if (r1 >= 0) o[(int64_t)r1 * p.ldp] = v;
        } else {
            cx<T> *o = reinterpret_cast<cx<T> *>(p.out) + (b * p.s.F * p.ldp + col);
            o[(int64_t)r0 * p.ldp] = X;
            if (r1 >= 0) o[(int64_t)r1 * p.ldp] = mk<T>(X.x, -X.y);
        }
    }
    // split of the half-length spectrum: A = Z[k] / 2, Zp = Z[LC - k] / 2 -> X[k], X[LC - k]; ONES: one-sided layouts,
    // row = bin (SPEC: |X|^2 real output); else the general epilogue
    template <bool ONES, bool SPEC>
    VHD static void store_half(const FwdCtx<T> &c, int lane, int sub, int nsub) {
        using VT = Vec<T>;
        using V = typename VT::V;
        constexpr int W = VT::W, LN = TF / W;
        const FwdArgs<T> &p = *c.a;
        const int LC = p.s.LC;
        const SEnt<T> *st = reinterpret_cast<const SEnt<T> *>(c.tab + p.lay.b_off);
        const int64_t esz = SPEC ? sizeof(T) : sizeof(cx<T>), rb = p.ldp * esz;
        const int64_t col = c.pt * TF + lane;
        char *ob = reinterpret_cast<char *>(p.out) + (c.b * p.s.F * p.ldp + col) * esz;
        bool ok[W];
#pragma unroll
        for (int h = 0; h < W; ++h) ok[h] = col + h * LN < p.P;
        const V fac = VT::bc(p.fac2x);                                            // 1 unless onesided2X
        const cx<V> *wl = c.work + lane;
        for (int k = sub; k <= LC / 2; k += nsub) {
            const SEnt<T> t = st[k];
            const cx<V> A = wl[(size_t)t.pa * LN], Zp = wl[(size_t)t.pb * LN];
            const cx<V> Bc = mk<V>(Zp.x, VT::neg(Zp.y));
            const cx<V> E = A + Bc, Om = A - Bc;
            const cx<V> D = mul_tw<false, T>(Om, t.wi, -t.wr);                    // (-i W_M^k) (A - B)
            cx<V> X = E + D, Y = E - D;
            Y.y = VT::neg(Y.y);
            if (ONES) {
                if (k == 0) { X.y = VT::bc((T)0); Y.y = VT::bc((T)0); }           // DC and Nyquist (:378-384)
                else { X.x = X.x * fac; X.y = X.y * fac; Y.x = Y.x * fac; Y.y = Y.y * fac; }   // :397-403
                if (SPEC) {
                    const V px = p2::vfma(X.x, X.x, X.y * X.y), py = p2::vfma(Y.x, Y.x, Y.y * Y.y);
#pragma unroll
                    for (int h = 0; h < W; ++h) {
                        if (!ok[h]) continue;
                        *reinterpret_cast<T *>(ob + k * rb + h * LN * esz) = VT::get(px, h);
                        if (2 * k != LC) *reinterpret_cast<T *>(ob + (LC - k) * rb + h * LN * esz) = VT::get(py, h);
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < W; ++h) {
                        if (!ok[h]) continue;
                        *reinterpret_cast<cx<T> *>(ob + k * rb + h * LN * esz) = mk<T>(VT::get(X.x, h), VT::get(X.y, h));
                        if (2 * k != LC) *reinterpret_cast<cx<T> *>(ob + (LC - k) * rb + h * LN * esz) = mk<T>(VT::get(Y.x, h), VT::get(Y.y, h));
                    }
                }
            } else {
#pragma unroll
                for (int h = 0; h < W; ++h) {
                    if (!ok[h]) continue;
                    emit_real(p, c.b, col + h * LN, k, mk<T>(VT::get(X.x, h), VT::get(X.y, h)));
                    if (2 * k != LC) emit_real(p, c.b, col + h * LN, LC - k, mk<T>(VT::get(Y.x, h), VT::get(Y.y, h)));
                }
            }
        }
    }
    VHD static void store(const FwdCtx<T> &c, int tid, int nt) {
        using VT = Vec<T>;
        using V = typename VT::V;
        constexpr int W = VT::W, LN = TF / W;
        const FwdArgs<T> &p = *c.a;
        const int lane = tid % LN, sub = tid / LN, nsub = nt / LN;
        const int64_t col = c.pt * TF + lane;
        if (col >= p.P) return;
        const int M = p.s.M, F = p.s.F;
        if (p.s.half) {
            if (p.mode >= 2) {
                if (p.spectrogram) store_half<true, true>(c, lane, sub, nsub);
                else store_half<true, false>(c, lane, sub, nsub);
            } else {
                store_half<false, false>(c, lane, sub, nsub);
            }
            return;
        }
        // odd mfft: full spectrum in the work buffer; row r of the stored layout takes bin k(r)
        const int *pos = reinterpret_cast<const int *>(c.tab + p.lay.b_off);
        for (int r = sub; r < F; r += nsub) {
            int k = r;
            if (p.mode == 1) { k = r + (M + 1) / 2; if (k >= M) k -= M; }       // fftshift: S[r] = X[(r + ceil(M/2)) mod M]
            const cx<V> vv = c.work[(size_t)pos[k] * LN + lane];
#pragma unroll
            for (int h = 0; h < W; ++h) {
                if (col + h * LN >= p.P) continue;
                cx<T> v = mk<T>(VT::get(vv.x, h), VT::get(vv.y, h));
                if (p.mode >= 2) {
                    if (k == 0 || 2 * k == M) v.y = 0;                            // :378-384
                    else if (p.mode == 3) { v.x *= p.fac2x; v.y *= p.fac2x; }     // :397-403
                }
                if (p.spectrogram) {
                    T *o = reinterpret_cast<T *>(p.out) + ((c.b * F + r) * p.ldp + col + h * LN);
                    *o = v.x * v.x + v.y * v.y;
                } else {
                    cx<T> *o = reinterpret_cast<cx<T> *>(p.out) + ((c.b * F + r) * p.ldp + col + h * LN);
                    *o = v;
                }
            }
        }
    }
};

// ================================================================== INVERSE
template <typename T>
struct InvArgs {
    const void *S;
    T *xo;
    const unsigned char *tab;  // plan blob (tab_layout<T>(s, true)); dual window / M (one-sided layouts) or / (2 M)
    TabLayout lay;
    int64_t P, ldp, out_pitch, p_min, k0, k1, q0, q1, qa0, ntiles;
    int mid, mode, halo, ostride;
    T rfac2x;
    Shape s;
};
template <typename T>
struct InvCtx {
    const InvArgs<T> *a;
    const unsigned char *tab;
    cx<vec_t<T>> *work;      // [LC][LN]
    T *ola;                  // [TF][ostride] windowed slices of the tile
    cx<T> *nyq;              // [TF] landed Nyquist row (one-sided layouts, even mfft)
    int64_t b, qa;
};

template <typename T, int TF>
struct Inv {
    VHD static cx<T> ldS(const InvArgs<T> &p, int64_t b, int row, int64_t col) {
        const T *q = reinterpret_cast<const T *>(p.S) + 2 * ((b * p.s.F + row) * p.ldp + col);
        return p2::ldgc(q);
    }
    // Hermitian bin k (0 <= k <= M/2) of the real signal's spectrum from the stored layout (src/lib.rs:414-487);
    // two-sided layouts: twice the Hermitian part X[k] + conj(X[M-k]) (the 1/2 is folded into the dual table)
    VHD static cx<T> bin_real(const InvArgs<T> &p, int64_t b, int64_t col, int k) {
        const int M = p.s.M;
        const bool edge = k == 0 || 2 * k == M;
        cx<T> v;
        if (p.mode >= 2) {
            v = ldS(p, b, k, col);
            if (p.mode == 3 && !edge) { v.x *= p.rfac2x; v.y *= p.rfac2x; }
        } else {
            int ra = k, rb = k == 0 ? 0 : M - k;
            if (p.mode == 1) {                               // stored row of bin i: (i + floor(M/2)) mod M
                ra += M / 2; if (ra >= M) ra -= M;
                rb += M / 2; if (rb >= M) rb -= M;
            }
            const cx<T> a = ldS(p, b, ra, col), bq = ldS(p, b, rb, col);
            v = mk<T>(a.x + bq.x, a.y - bq.y);
        }
        if (edge) v.y = 0;                                   // irfft semantics
        return v;
    }
    // Z'[k] = 2 Z[k], Z'[LC - k] from X[k], X[LC - k] (ONES: one-sided layouts, row = bin)
    template <bool ONES>
    VHD static void load_half(const InvCtx<T> &c, int lane, int sub, int nsub) {
        using VT = Vec<T>;
        using V = typename VT::V;
        constexpr int W = VT::W, LN = TF / W;
        const InvArgs<T> &p = *c.a;
        const int LC = p.s.LC;
        const SEnt<T> *st = reinterpret_cast<const SEnt<T> *>(c.tab + p.lay.a_off);
        const cx<T> zero = mk<T>((T)0, (T)0);
        bool live[W];
        int64_t col[W];
        const T *sb[W];
#pragma unroll
        for (int h = 0; h < W; ++h) {
            const int64_t q = c.qa + lane + h * LN;
            col[h] = q - p.p_min;
            live[h] = q >= p.q0 && q < p.q1 && col[h] >= 0 && col[h] < p.P;    // :855-857
            sb[h] = reinterpret_cast<const T *>(p.S) + 2 * (c.b * p.s.F * p.ldp + (live[h] ? col[h] : 0));
        }
        const int64_t rb = 2 * p.ldp;
        const V rfac = VT::bc(p.rfac2x);                                          // 1 unless onesided2X
        cx<V> *wl = c.work + lane;
#pragma unroll 2
        for (int k = sub; k <= LC / 2; k += nsub) {
            cx<T> xk[W], xl[W];
#pragma unroll
            for (int h = 0; h < W; ++h) {
                xk[h] = xl[h] = zero;
                if (live[h]) {
                    if (ONES) {
                        xk[h] = p2::ldgc(sb[h] + k * rb);
                        xl[h] = p2::ldgc(sb[h] + (LC - k) * rb);
                    } else {
                        xk[h] = bin_real(p, c.b, col[h], k);
                        xl[h] = bin_real(p, c.b, col[h], LC - k);
                    }
                }
            }
            cx<V> Xk = mk<V>(VT::pk(xk[0].x, xk[W - 1].x), VT::pk(xk[0].y, xk[W - 1].y));
            cx<V> Xl = mk<V>(VT::pk(xl[0].x, xl[W - 1].x), VT::pk(xl[0].y, xl[W - 1].y));
            cx<V> za, zb;
            if (k == 0) {
                za = mk<V>(Xk.x + Xl.x, Xk.x - Xl.x);                             // DC pairs with Nyquist (imaginary parts ignored)
                zb = za;
            } else {
                if (ONES) { Xk.x = Xk.x * rfac; Xk.y = Xk.y * rfac; Xl.x = Xl.x * rfac; Xl.y = Xl.y * rfac; }
                const SEnt<T> t = st[k];
                const cx<V> Xm = mk<V>(Xl.x, VT::neg(Xl.y));
                const cx<V> E = Xk + Xm, Dm = Xk - Xm;
                const cx<V> O = mul_tw<true, T>(Dm, t.wr, t.wi);                  // conj(W_M^k) (Xk - conj Xl)
                const cx<V> iO = mk<V>(VT::neg(O.y), O.x);
                za = E + iO;
                const cx<V> Y = E - iO;
                zb = mk<V>(Y.x, VT::neg(Y.y));
            }
            wl[(size_t)k * LN] = za;
            if (k != 0 && 2 * k != LC) wl[(size_t)(LC - k) * LN] = zb;
        }
    }
    // ---- one-sided layouts, even mfft: the S tile LANDS in the work buffer through cp.async while the previous
    // tile's overlap-add runs (the direct loads left 25 % of the warp time waiting for S: ncu, long scoreboard
    // at the first add of the merge), and the Hermitian merge runs in place from shared memory.  Raw layout: row
    // k (bin) = TF complex slots (slice s at slot s) -- exactly the bytes of work row k; the Nyquist row has its own
    // buffer.  Slices outside [q0, q1) or S arrive as zeros.
    VHD static bool lands(const InvArgs<T> &p) { return p.s.half != 0 && p.mode >= 2; }
    VHD static void land(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int LC = p.s.LC;
        if (sizeof(cx<T>) == 8 && TF % 2 == 0 && ((c.qa - p.p_min) & 1) == 0 && (p.ldp & 1) == 0 &&
            (reinterpret_cast<size_t>(p.S) & 15) == 0) {
            // f32, even first column: two adjacent slices per 16-byte copy (both in range or both out of it; a
            // straddling pair falls back to two 8-byte copies)
            const int s2 = 2 * (tid % (TF / 2));
            const int64_t q = c.qa + s2, col = q - p.p_min;
            const bool l0 = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;
            const bool l1 = q + 1 >= p.q0 && q + 1 < p.q1 && col + 1 >= 0 && col + 1 < p.P;
            const cx<T> *src = reinterpret_cast<const cx<T> *>(p.S) + (c.b * p.s.F * p.ldp + ((l0 || l1) ? col : 0));
            cx<T> *raw = reinterpret_cast<cx<T> *>(c.work) + s2;
            for (int r = tid / (TF / 2); r <= LC; r += nt / (TF / 2)) {
                cx<T> *dst = r < LC ? raw + (size_t)r * TF : c.nyq + s2;
                const cx<T> *g = src + (int64_t)r * p.ldp;
                if (l0 == l1) {
                    p2::cp_async_zfill<16>(dst, g, l0);
                } else {
                    p2::cp_async_zfill<8>(dst, l0 ? g : g + 1, l0);
                    p2::cp_async_zfill<8>(dst + 1, g + 1, l1);
                }
            }
            return;
        }
        const int s = tid % TF;
        const int64_t q = c.qa + s, col = q - p.p_min;
        const bool live = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;        // :855-857
        const cx<T> *src = reinterpret_cast<const cx<T> *>(p.S) + (c.b * p.s.F * p.ldp + (live ? col : 0));
        cx<T> *raw = reinterpret_cast<cx<T> *>(c.work) + s;
        for (int r = tid / TF; r <= LC; r += nt / TF)
            p2::cp_async_zfill<(int)sizeof(cx<T>)>(r < LC ? raw + (size_t)r * TF : c.nyq + s, src + (int64_t)r * p.ldp, live);
    }
    struct MReg { cx<T> xk[Vec<T>::W], xl[Vec<T>::W]; };
    VHD static int merge_iters(const InvArgs<T> &p, int nt) { const int nsub = nt / (TF / Vec<T>::W); return (p.s.LC / 2 + nsub) / nsub; }
    // first half: the four raw bins of unit k = sub + it nsub (both slices of the lane, rows k and LC - k) into registers
    VHD static void merge_load(const InvCtx<T> &c, int tid, int it, int nt, MReg &r) {
        constexpr int W = Vec<T>::W, LN = TF / W;
        const int LC = c.a->s.LC, lane = tid % LN, k = tid / LN + it * (nt / LN);
        if (k > LC / 2) return;
        const cx<T> *raw = reinterpret_cast<const cx<T> *>(c.work);
#pragma unroll
        for (int h = 0; h < W; ++h) {
            r.xk[h] = raw[(size_t)k * TF + lane + h * LN];
            r.xl[h] = k == 0 ? c.nyq[lane + h * LN] : raw[(size_t)(LC - k) * TF + lane + h * LN];
        }
    }
    // second half (every lane of the unit has read both rows): Z'[k], Z'[LC - k] over the same rows
    VHD static void merge_store(const InvCtx<T> &c, int tid, int it, int nt, const MReg &r) {
        using VT = Vec<T>;
        using V = typename VT::V;
        constexpr int W = VT::W, LN = TF / W;
        const InvArgs<T> &p = *c.a;
        const int LC = p.s.LC, lane = tid % LN, k = tid / LN + it * (nt / LN);
        if (k > LC / 2) return;
        cx<V> Xk = mk<V>(VT::pk(r.xk[0].x, r.xk[W - 1].x), VT::pk(r.xk[0].y, r.xk[W - 1].y));
        cx<V> Xl = mk<V>(VT::pk(r.xl[0].x, r.xl[W - 1].x), VT::pk(r.xl[0].y, r.xl[W - 1].y));
        cx<V> za, zb;
        if (k == 0) {
            za = mk<V>(Xk.x + Xl.x, Xk.x - Xl.x);                                 // DC pairs with Nyquist (imaginary parts ignored)
            zb = za;
        } else {
            const V rfac = VT::bc(p.rfac2x);                                      // 1 unless onesided2X
            Xk.x = Xk.x * rfac; Xk.y = Xk.y * rfac; Xl.x = Xl.x * rfac; Xl.y = Xl.y * rfac;
            const SEnt<T> t = reinterpret_cast<const SEnt<T> *>(c.tab + p.lay.a_off)[k];
            const cx<V> Xm = mk<V>(Xl.x, VT::neg(Xl.y));
            const cx<V> E = Xk + Xm, Dm = Xk - Xm;
            const cx<V> O = mul_tw<true, T>(Dm, t.wr, t.wi);                      // conj(W_M^k) (Xk - conj Xl)
            const cx<V> iO = mk<V>(VT::neg(O.y), O.x);
            za = E + iO;
            const cx<V> Y = E - iO;
            zb = mk<V>(Y.x, VT::neg(Y.y));
        }
        cx<V> *wl = c.work + lane;
        wl[(size_t)k * LN] = za;
        if (k != 0 && 2 * k != LC) wl[(size_t)(LC - k) * LN] = zb;
    }
    // device flavour of the whole merge (the host emulation runs the two halves for all threads in turn)
    VHD static void merge(const InvCtx<T> &c, int tid, int nt) {
        const int iters = merge_iters(*c.a, nt);
        for (int it = 0; it < iters; ++it) {
            MReg r;
            merge_load(c, tid, it, nt, r);
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
            merge_store(c, tid, it, nt, r);
        }
    }
    // spectrum of slice q into the work buffer in natural order; only the real part of the inverse transform
    // is kept (:903-907), so two-sided layouts enter Hermitian-symmetrised
    VHD static void load(const InvCtx<T> &c, int tid, int nt) {
        using VT = Vec<T>;
        using V = typename VT::V;
        constexpr int W = VT::W, LN = TF / W;
        const InvArgs<T> &p = *c.a;
        const int lane = tid % LN, sub = tid / LN, nsub = nt / LN;
        const int M = p.s.M;
        if (p.s.half) {
            if (p.mode >= 2) load_half<true>(c, lane, sub, nsub);
            else load_half<false>(c, lane, sub, nsub);
            return;
        }
        for (int i = sub; i < M; i += nsub) {
            cx<T> vs[W];
#pragma unroll
            for (int h = 0; h < W; ++h) {
                const int64_t q = c.qa + lane + h * LN, col = q - p.p_min;
                const bool live = q >= p.q0 && q < p.q1 && col >= 0 && col < p.P;    // :855-857
                cx<T> v = mk<T>((T)0, (T)0);
                if (live) {
                    if (p.mode >= 2) {
                        const int k = 2 * i <= M ? i : M - i;
                        v = ldS(p, c.b, k, col);
                        if (k == 0 || 2 * k == M) v.y = 0;                         // irfft semantics
                        else if (p.mode == 3) { v.x *= p.rfac2x; v.y *= p.rfac2x; }
                        if (k != i) v.y = -v.y;                                    // X[M - k] = conj(X[k])
                    } else {
                        // stored row of bin i: twosided r = i; centered r = (i - ceil(M/2)) mod M = (i + floor(M/2)) mod M
                        int ra = p.mode == 1 ? i + M / 2 : i;
                        if (ra >= M) ra -= M;
                        const int im = i == 0 ? 0 : M - i;
                        int rb = p.mode == 1 ? im + M / 2 : im;
                        if (rb >= M) rb -= M;
                        const cx<T> a = ldS(p, c.b, ra, col), bq = ldS(p, c.b, rb, col);
                        v = mk<T>(a.x + bq.x, a.y - bq.y);  // twice the Hermitian part: the 1/2 is folded into the dual table
                    }
                }
                vs[h] = v;
            }
            c.work[(size_t)i * LN + lane] = mk<V>(VT::pk(vs[0].x, vs[W - 1].x), VT::pk(vs[0].y, vs[W - 1].y));
        }
    }
    // passes; the last one writes xs[j] = Re(y[(j - ps) mod M]) * dual[j] / M, j < N (src/lib.rs:493-503, 869-873)
    VHD static void pass(const InvCtx<T> &c, int pi, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        PassIO<T> io;
        io.work = c.work; io.xin = nullptr; io.lt = nullptr; io.hp = 0; io.half = p.s.half;
        io.ola = c.ola; io.ostride = p.ostride;
        io.rt = reinterpret_cast<const REnt<T> *>(c.tab + p.lay.b_off);
        const cx<T> *tw = reinterpret_cast<const cx<T> *>(c.tab + p.lay.tw_off) + p.s.twoff[pi];
        const int S = p.s.L[pi], units = p.s.units[pi];
        if (pi == p.s.np - 1) pass_any<T, TF, true, 2>(p.s.R[pi], io, tw, S, p.s.magicS[pi], units, tid, nt);
        else pass_any<T, TF, true, 0>(p.s.R[pi], io, tw, S, p.s.magicS[pi], units, tid, nt);
    }
    // overlap-add as a gather: the tile owns the samples whose LATEST contributing slice is one of its
    // slices halo .. TF-1; sums run over increasing q (src/lib.rs:853)
    template <int HALO>
    VHD static void gather_t(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int H = p.s.H, N = p.s.N, halo = HALO >= 0 ? HALO : p.halo;
        const int64_t base = c.qa * (int64_t)H - p.mid;                   // k = base + u
        int64_t ka = base + (int64_t)halo * H, kb = base + (int64_t)TF * H;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        T *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const int cnt = (TF - halo) * H, hH = halo * H, step = p.ostride - H;
        int rl = mdiv(tid, p.s.magicH), col = tid - rl * H;               // owned sample e = rl H + col, u = e + halo H
        const int dq = mdiv(nt, p.s.magicH), dr = nt - dq * H;
        for (int e = tid; e < cnt; e += nt) {
            const int u = e + hH;
            if (u >= ua && u < ub) {
                // contributing slices r = rl .. rl + halo (tile index), offset inside slice r: u - r H; only the
                // oldest one can fall beyond the window length (halo H < N)
                const T *q = c.ola + (size_t)rl * p.ostride + col + hH;
                T acc = (col + hH < N) ? q[0] : (T)0;
                if (HALO >= 0) {
#pragma unroll
                    for (int h = 1; h <= (HALO >= 0 ? HALO : 0); ++h) acc += q[h * step];
                } else {
                    for (int h = 1; h <= halo; ++h) acc += q[h * step];
                }
                xo[u] = acc;
            }
            col += dr;
            rl += dq;
            if (col >= H) { col -= H; ++rl; }
        }
    }
    // one warp per hop row of owned samples: lane + 32 i walks the row, so the slice-row reads of a warp are
    // consecutive words (conflict free), the stores 128-byte runs, and the only index work per sample is one add.
    // (The four-samples-per-thread variant reads the rows with a stride of four words -- 4-way bank conflicts -- but
    // stores 16 bytes per thread; it stays ahead in f32.)
    template <int HALO>
    VHD static void gather_rows_t(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int H = p.s.H, N = p.s.N;
        const int64_t base = c.qa * (int64_t)H - p.mid;                   // k = base + u
        int64_t ka = base + (int64_t)HALO * H, kb = base + (int64_t)TF * H;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        T *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const int hH = HALO * H, step = p.ostride - H;
        const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
        for (int rl = warp; rl < TF - HALO; rl += nwarp) {                // owned row rl: samples u = (rl + HALO) H + col
            const T *q = c.ola + (size_t)rl * p.ostride + hH;
            const int u0 = rl * H + hH;
            const bool inner = u0 >= ua && u0 + H <= ub;
            for (int col = lane; col < H; col += 32) {
                // contributing slices rl .. rl + HALO, oldest first; only the oldest can fall beyond the window length
                T acc = (col + hH < N) ? q[col] : (T)0;
#pragma unroll
                for (int h = 1; h <= HALO; ++h) acc += q[h * step + col];
                if (inner || (u0 + col >= ua && u0 + col < ub)) xo[u0 + col] = acc;
            }
        }
    }
    // four consecutive samples per step (hop % 4 == 0: they share their contributing slices): a quarter of the index
    // work per sample and one 16-byte store when the output is aligned (ncu: the scalar gather was 26 % of the
    // inverse's instructions)
    struct alignas(16) Quad { T v[4]; };
    template <int HALO>
    VHD static void gather4_t(const InvCtx<T> &c, int tid, int nt) {
        const InvArgs<T> &p = *c.a;
        const int H = p.s.H, N = p.s.N;
        const int64_t base = c.qa * (int64_t)H - p.mid;                   // k = base + u
        int64_t ka = base + (int64_t)HALO * H, kb = base + (int64_t)TF * H;
        if (ka < p.k0) ka = p.k0;
        if (kb > p.k1) kb = p.k1;
        T *xo = p.xo + c.b * p.out_pitch - p.k0 + base;
        const int ua = (int)(ka - base), ub = (int)(kb - base);
        const int cnt = (TF - HALO) * H, hH = HALO * H, step = p.ostride - H;
        const bool al = sizeof(T) == 4 && ((reinterpret_cast<size_t>(xo + hH) & 15) == 0);   // e4 and hH are multiples of 4
        for (int e = 4 * tid; e < cnt; e += 4 * nt) {
            const int u = e + hH;
            if (u + 3 < ua || u >= ub) continue;
            const int rl = mdiv(e, p.s.magicH), col = e - rl * H;
            const T *q = c.ola + (size_t)rl * p.ostride + col + hH;
            // rows are 8-byte aligned (ostride_of) and col, hH multiples of 4: two 8-byte reads per row
            struct alignas(8) Pair { T a, b; };
            Quad acc;
            {
                const Pair p0 = *reinterpret_cast<const Pair *>(q), p1 = *reinterpret_cast<const Pair *>(q + 2);
                acc.v[0] = (col + hH < N) ? p0.a : (T)0;
                acc.v[1] = (col + 1 + hH < N) ? p0.b : (T)0;
                acc.v[2] = (col + 2 + hH < N) ? p1.a : (T)0;
                acc.v[3] = (col + 3 + hH < N) ? p1.b : (T)0;
            }
#pragma unroll
            for (int h = 1; h <= HALO; ++h) {
                const Pair p0 = *reinterpret_cast<const Pair *>(q + h * step), p1 = *reinterpret_cast<const Pair *>(q + h * step + 2);
                acc.v[0] += p0.a; acc.v[1] += p0.b; acc.v[2] += p1.a; acc.v[3] += p1.b;
            }
            if (al && u >= ua && u + 3 < ub) {
                *reinterpret_cast<Quad *>(xo + u) = acc;
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (u + i >= ua && u + i < ub) xo[u + i] = acc.v[i];
            }
        }
    }
    VHD static void gather(const InvCtx<T> &c, int tid, int nt) {
        // measured (Hann 400/160 inverse): f32 four-sample 31.9 % / warp rows 30.5 %; f64 four-sample 27.9 % / warp rows 37.4 %
        if (sizeof(T) == 4 && c.a->s.H % 4 == 0 && c.a->halo >= 1 && c.a->halo <= 3) {
            switch (c.a->halo) {
                case 1: gather4_t<1>(c, tid, nt); break;
                case 2: gather4_t<2>(c, tid, nt); break;
                default: gather4_t<3>(c, tid, nt); break;
            }
            return;
        }
        if (nt >= 32 && nt % 32 == 0 && c.a->halo >= 1 && c.a->halo <= 3) {
            switch (c.a->halo) {
                case 1: gather_rows_t<1>(c, tid, nt); break;
                case 2: gather_rows_t<2>(c, tid, nt); break;
                default: gather_rows_t<3>(c, tid, nt); break;
            }
            return;
        }
        switch (c.a->halo) {
            case 0: gather_t<0>(c, tid, nt); break;
            case 1: gather_t<1>(c, tid, nt); break;
            case 2: gather_t<2>(c, tid, nt); break;
            case 3: gather_t<3>(c, tid, nt); break;
            default: gather_t<-1>(c, tid, nt); break;
        }
    }
};

}  // namespace sm
}  // namespace veils
