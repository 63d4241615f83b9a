// smooth_dft.cuh -- in-register DFTs of every length 2 .. 16 for the mixed-radix ("smooth") family,
// host + device.  Twiddles are compile-time constants (constexpr sine / cosine), so the fully
// unrolled butterflies multiply by immediates.
//   * 2, 4, 8, 16: the power-of-two butterflies of pow2_tile_impl.cuh;
//   * odd lengths 3, 5, 7, 11, 13: the symmetric form  X[k] = x0 + sum_n (x_n + x_{R-n}) cos(2 pi n k / R)
//     -/+ i (x_n - x_{R-n}) sin(2 pi n k / R), n = 1 .. (R-1)/2 -- half the multiplies of the direct sum;
//   * 6, 9, 10, 12, 14, 15: Cooley-Tukey composition A x B of the above.
// The reference plans any length through rustfft (src/lib.rs:209-211); its own golden vector uses
// mfft = 15 (examples/basic_usage.rs:150-160).
#pragma once
#include "pow2_cfg.hpp"

namespace veils {
namespace sm {

using p2::cx;
using p2::mk;

// ---- constexpr sine / cosine of 2 pi m / n: reduction to [-pi/2, pi/2], 16-term Taylor series
// (error < 1e-18), exact values on the axes
#if defined(__CUDACC__)
#define VCX __host__ __device__ constexpr
#else
#define VCX constexpr
#endif
constexpr double kPi = 3.14159265358979323846264338327950288;
VCX double sin_series(double x) {
    const double x2 = x * x;
    double term = x, sum = x;
    for (int i = 1; i < 16; ++i) {
        term *= -x2 / ((2 * i) * (2 * i + 1));
        sum += term;
    }
    return sum;
}
VCX double cos_series(double x) {
    const double x2 = x * x;
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < 16; ++i) {
        term *= -x2 / ((2 * i - 1) * (2 * i));
        sum += term;
    }
    return sum;
}
VCX int mod_pos(int m, int n) { return ((m % n) + n) % n; }
VCX double cos2pi(int m, int n) {
    m = mod_pos(m, n);
    if ((4 * m) % n == 0) { const int q = 4 * m / n; return q == 0 ? 1.0 : (q == 2 ? -1.0 : 0.0); }
    if (2 * m > n) m = n - m;                               // cos is even around pi
    if (4 * m > n) return -cos_series(kPi - 2.0 * kPi * m / n);   // second quadrant
    return cos_series(2.0 * kPi * m / n);
}
VCX double sin2pi(int m, int n) {
    m = mod_pos(m, n);
    if ((4 * m) % n == 0) { const int q = 4 * m / n; return q == 1 ? 1.0 : (q == 3 ? -1.0 : 0.0); }
    if (2 * m > n) return -sin2pi(n - m, n);                // sin is odd around pi
    if (4 * m > n) return sin_series(kPi - 2.0 * kPi * m / n);
    return sin_series(2.0 * kPi * m / n);
}

// cos / sin of 2 pi m / R for m = 0 .. R-1 as a compile-time table: a `constexpr` local object of this
// type is evaluated by the front end (host and device alike), and after unrolling every index is a
// literal -- the butterflies multiply by immediates
template <int R> struct Trig {
    double c[R], s[R];
};
template <int R> VCX Trig<R> make_trig() {
    Trig<R> t{};
    for (int m = 0; m < R; ++m) {
        t.c[m] = cos2pi(m, R);
        t.s[m] = sin2pi(m, R);
    }
    return t;
}

// a * (c - i s) forward, a * (c + i s) inverse, c / s compile-time; T = float, double or the packed pair f2
template <bool INV, typename T> VHD cx<T> mulw(cx<T> a, double c, double s) {
    const T cc = p2::vconst<T>(c), ss = p2::vconst<T>(INV ? -s : s), ns = p2::vconst<T>(INV ? s : -s);
    return mk<T>(p2::vfma(a.y, ss, a.x * cc), p2::vfma(a.x, ns, a.y * cc));
}

template <int R, bool INV, typename T> struct Dft;
template <bool INV, typename T> struct Dft<1, INV, T> { VHD static void run(cx<T> (&)[1]) {} };
template <bool INV, typename T> struct Dft<2, INV, T> { VHD static void run(cx<T> (&v)[2]) { p2::Dft<2, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<4, INV, T> { VHD static void run(cx<T> (&v)[4]) { p2::Dft<4, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<8, INV, T> { VHD static void run(cx<T> (&v)[8]) { p2::Dft<8, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<16, INV, T> { VHD static void run(cx<T> (&v)[16]) { p2::Dft<16, INV, T>::run(v); } };

// odd length, symmetric form
template <int R, bool INV, typename T> struct DftOdd {
    VHD static void run(cx<T> (&v)[R]) {
        constexpr int H = (R - 1) / 2;
        constexpr Trig<R> tt = make_trig<R>();
        cx<T> s[H], d[H];
#pragma unroll
        for (int n = 1; n <= H; ++n) {
            s[n - 1] = v[n] + v[R - n];
            d[n - 1] = v[n] - v[R - n];
        }
        cx<T> o[R];
        o[0] = v[0];
#pragma unroll
        for (int n = 0; n < H; ++n) o[0] = o[0] + s[n];
#pragma unroll
        for (int k = 1; k <= H; ++k) {
            T ar = v[0].x, ai = v[0].y;
            T br = d[0].x * p2::vconst<T>(tt.s[k % R]), bi = d[0].y * p2::vconst<T>(tt.s[k % R]);
            ar = p2::vfma(s[0].x, p2::vconst<T>(tt.c[k % R]), ar);
            ai = p2::vfma(s[0].y, p2::vconst<T>(tt.c[k % R]), ai);
#pragma unroll
            for (int n = 2; n <= H; ++n) {
                const T c = p2::vconst<T>(tt.c[(n * k) % R]), sn = p2::vconst<T>(tt.s[(n * k) % R]);
                ar = p2::vfma(s[n - 1].x, c, ar);
                ai = p2::vfma(s[n - 1].y, c, ai);
                br = p2::vfma(d[n - 1].x, sn, br);
                bi = p2::vfma(d[n - 1].y, sn, bi);
            }
            // forward: X[k] = a - i b, X[R-k] = a + i b  (b = sum d sin);  inverse: the other way round
            const cx<T> lo = INV ? mk<T>(ar - bi, ai + br) : mk<T>(ar + bi, ai - br);
            const cx<T> hi = INV ? mk<T>(ar + bi, ai - br) : mk<T>(ar - bi, ai + br);
            o[k] = lo;
            o[R - k] = hi;
        }
#pragma unroll
        for (int i = 0; i < R; ++i) v[i] = o[i];
    }
};
template <bool INV, typename T> struct Dft<3, INV, T> { VHD static void run(cx<T> (&v)[3]) { DftOdd<3, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<5, INV, T> { VHD static void run(cx<T> (&v)[5]) { DftOdd<5, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<7, INV, T> { VHD static void run(cx<T> (&v)[7]) { DftOdd<7, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<11, INV, T> { VHD static void run(cx<T> (&v)[11]) { DftOdd<11, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<13, INV, T> { VHD static void run(cx<T> (&v)[13]) { DftOdd<13, INV, T>::run(v); } };

// Cooley-Tukey A x B: n = B n1 + n2, k = k1 + A k2
template <int A, int B, bool INV, typename T> struct DftComp {
    VHD static void run(cx<T> (&v)[A * B]) {
        constexpr Trig<A * B> tt = make_trig<A * B>();
        cx<T> t[B][A];
#pragma unroll
        for (int n2 = 0; n2 < B; ++n2) {
            cx<T> u[A];
#pragma unroll
            for (int n1 = 0; n1 < A; ++n1) u[n1] = v[B * n1 + n2];
            Dft<A, INV, T>::run(u);
#pragma unroll
            for (int k1 = 0; k1 < A; ++k1)
                t[n2][k1] = (n2 * k1 == 0) ? u[k1] : mulw<INV>(u[k1], tt.c[(n2 * k1) % (A * B)], tt.s[(n2 * k1) % (A * B)]);
        }
#pragma unroll
        for (int k1 = 0; k1 < A; ++k1) {
            cx<T> u[B];
#pragma unroll
            for (int n2 = 0; n2 < B; ++n2) u[n2] = t[n2][k1];
            Dft<B, INV, T>::run(u);
#pragma unroll
            for (int k2 = 0; k2 < B; ++k2) v[k1 + A * k2] = u[k2];
        }
    }
};
template <bool INV, typename T> struct Dft<6, INV, T> { VHD static void run(cx<T> (&v)[6]) { DftComp<2, 3, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<9, INV, T> { VHD static void run(cx<T> (&v)[9]) { DftComp<3, 3, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<10, INV, T> { VHD static void run(cx<T> (&v)[10]) { DftComp<2, 5, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<12, INV, T> { VHD static void run(cx<T> (&v)[12]) { DftComp<4, 3, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<14, INV, T> { VHD static void run(cx<T> (&v)[14]) { DftComp<2, 7, INV, T>::run(v); } };
template <bool INV, typename T> struct Dft<15, INV, T> { VHD static void run(cx<T> (&v)[15]) { DftComp<3, 5, INV, T>::run(v); } };

// ---- host side: factorisation of mfft into <= 4 passes of radix 2 .. 16
constexpr int kMaxPasses = 4;
struct Factors {
    int np = 0;
    int R[kMaxPasses] = {1, 1, 1, 1};
};
// fewest passes first, then the most balanced radices; false when mfft has a prime factor > 13 or needs
// more than four passes
inline bool factorize(int64_t M, Factors *out) {
    if (M < 2) return false;
    int best_np = kMaxPasses + 1, best_max = 1 << 30;
    Factors best, cur;
    // depth-first over non-increasing radices
    struct Rec {
        static void go(int64_t rem, int depth, int maxr, Factors &cur, Factors &best, int &best_np, int &best_max) {
            if (rem == 1) {
                int mx = 0;
                for (int i = 0; i < depth; ++i) mx = cur.R[i] > mx ? cur.R[i] : mx;
                if (depth < best_np || (depth == best_np && mx < best_max)) {
                    best = cur; best.np = depth; best_np = depth; best_max = mx;
                    for (int i = depth; i < kMaxPasses; ++i) best.R[i] = 1;
                }
                return;
            }
            if (depth == kMaxPasses) return;
            for (int r = maxr; r >= 2; --r) {
                if (rem % r) continue;
                cur.R[depth] = r;
                go(rem / r, depth + 1, r, cur, best, best_np, best_max);
            }
        }
    };
    Rec::go(M, 0, 16, cur, best, best_np, best_max);
    if (best_np > kMaxPasses) return false;
    *out = best;
    return true;
}

}  // namespace sm
}  // namespace veils
