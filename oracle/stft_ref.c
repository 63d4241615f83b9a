/* stft_ref.c -- plain-C restatement of the reference's CPU path.  TEST INFRASTRUCTURE / CPU
 * BASELINE, NOT PRODUCT: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  Parity status: pinned through tests/test_oracle_c.py, which
 * holds it to the numpy oracle (itself pinned to the reference's golden vectors).
 *
 * It keeps the per-slice STRUCTURE of ultrasaw/veils src/lib.rs so that its timing is a fair
 * stand-in for the Rust crate (which cannot be built here: no cargo, rustfft not vendored):
 *   - x_slices materialises every frame as complex f64          (src/lib.rs:612-652)
 *   - window multiply into a fresh buffer per frame             (src/lib.rs:715-729)
 *   - zero-pad to mfft, rotate_left(p_s), FULL complex FFT,
 *     mode epilogue                                             (src/lib.rs:336-411)
 *   - transpose into [f_pts][P]                                 (src/lib.rs:735-746)
 *   - istft: strided column gather, mode prologue, full complex
 *     inverse FFT, 1/mfft, rotate_right(p_s), truncate, dual
 *     window, overlap-add in increasing q                       (src/lib.rs:414-505, 853-911)
 * The FFT itself (rustfft 6.4.1 in the reference, not vendored) is restated as an iterative
 * radix-2 transform for power-of-two lengths and Bluestein's algorithm otherwise.
 * Rows of a batch are independent; they are distributed over OpenMP threads (the reference is
 * single threaded -- the thread count used is reported next to every number).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cpx;

typedef struct {
    int64_t n;          /* transform length */
    int pow2;
    cpx *tw;            /* radix-2: exp(-2 pi i k / n), k < n/2 */
    /* Bluestein */
    int64_t m;          /* padded power-of-two length >= 2n-1 */
    cpx *chirp;         /* exp(-i pi k^2 / n) */
    cpx *bfft;          /* FFT_m of the conjugate chirp */
    cpx *tw_m;
} fft_plan;

static void fft_pow2(cpx *a, int64_t n, const cpx *tw, int inverse) {
    for (int64_t i = 1, j = 0; i < n; ++i) {          /* bit reversal */
        int64_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { cpx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int64_t len = 2; len <= n; len <<= 1) {
        const int64_t half = len >> 1, step = n / len;
        for (int64_t i = 0; i < n; i += len)
            for (int64_t k = 0; k < half; ++k) {
                cpx w = tw[k * step];
                if (inverse) w.im = -w.im;
                cpx *u = &a[i + k], *v = &a[i + k + half];
                const double tr = v->re * w.re - v->im * w.im, ti = v->re * w.im + v->im * w.re;
                v->re = u->re - tr; v->im = u->im - ti;
                u->re += tr; u->im += ti;
            }
    }
}

static cpx *make_tw(int64_t n) {
    cpx *tw = (cpx *)malloc(sizeof(cpx) * (size_t)(n / 2 + 1));
    for (int64_t k = 0; k < n / 2 + 1; ++k) {
        const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)k / (long double)n;
        tw[k].re = (double)cosl(a); tw[k].im = (double)sinl(a);
    }
    return tw;
}

static fft_plan *plan_make(int64_t n) {
    fft_plan *p = (fft_plan *)calloc(1, sizeof(fft_plan));
    p->n = n;
    p->pow2 = (n & (n - 1)) == 0;
    if (p->pow2) { p->tw = make_tw(n); return p; }
    int64_t m = 1;
    while (m < 2 * n - 1) m <<= 1;
    p->m = m;
    p->tw_m = make_tw(m);
    p->chirp = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    p->bfft = (cpx *)calloc((size_t)m, sizeof(cpx));
    for (int64_t k = 0; k < n; ++k) {
        const int64_t k2 = (k * k) % (2 * n);
        const long double a = -3.14159265358979323846264338327950288L * (long double)k2 / (long double)n;
        p->chirp[k].re = (double)cosl(a); p->chirp[k].im = (double)sinl(a);
    }
    p->bfft[0].re = p->chirp[0].re; p->bfft[0].im = -p->chirp[0].im;
    for (int64_t k = 1; k < n; ++k) {
        p->bfft[k].re = p->bfft[m - k].re = p->chirp[k].re;
        p->bfft[k].im = p->bfft[m - k].im = -p->chirp[k].im;
    }
    fft_pow2(p->bfft, m, p->tw_m, 0);
    return p;
}

static void plan_free(fft_plan *p) {
    if (!p) return;
    free(p->tw); free(p->chirp); free(p->bfft); free(p->tw_m); free(p);
}

/* unnormalised DFT of length p->n, in place; scratch must hold p->m elements (Bluestein) */
static void fft_any(const fft_plan *p, cpx *a, int inverse, cpx *scratch) {
    const int64_t n = p->n;
    if (p->pow2) { fft_pow2(a, n, p->tw, inverse); return; }
    if (inverse) for (int64_t k = 0; k < n; ++k) a[k].im = -a[k].im;     /* IDFT = conj(DFT(conj)) */
    const int64_t m = p->m;
    for (int64_t k = 0; k < n; ++k) {
        scratch[k].re = a[k].re * p->chirp[k].re - a[k].im * p->chirp[k].im;
        scratch[k].im = a[k].re * p->chirp[k].im + a[k].im * p->chirp[k].re;
    }
    memset(scratch + n, 0, sizeof(cpx) * (size_t)(m - n));
    fft_pow2(scratch, m, p->tw_m, 0);
    for (int64_t k = 0; k < m; ++k) {
        const double r = scratch[k].re * p->bfft[k].re - scratch[k].im * p->bfft[k].im;
        const double i = scratch[k].re * p->bfft[k].im + scratch[k].im * p->bfft[k].re;
        scratch[k].re = r; scratch[k].im = i;
    }
    fft_pow2(scratch, m, p->tw_m, 1);
    const double s = 1.0 / (double)m;
    for (int64_t k = 0; k < n; ++k) {
        const double r = scratch[k].re * s, i = scratch[k].im * s;
        a[k].re = r * p->chirp[k].re - i * p->chirp[k].im;
        a[k].im = r * p->chirp[k].im + i * p->chirp[k].re;
        if (inverse) a[k].im = -a[k].im;
    }
}

static void rotate_left(cpx *a, int64_t n, int64_t k, cpx *tmp) {
    if (n == 0) return;
    k %= n;
    if (k == 0) return;
    memcpy(tmp, a, sizeof(cpx) * (size_t)k);
    memmove(a, a + k, sizeof(cpx) * (size_t)(n - k));
    memcpy(a + n - k, tmp, sizeof(cpx) * (size_t)k);
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* mode: 0 twosided, 1 centered, 2 onesided, 3 onesided2X.  S: [batch][F][P] complex f64.
 * Geometry (p0, p1, ps, F) is supplied by the caller (host planner / oracle). */
int ref_stft(const double *x, int64_t n, int64_t batch, const double *win, int64_t N, int64_t H,
             int64_t M, int mode, int64_t ps, double fac2x, int64_t p0, int64_t p1, int64_t k_offset,
             double *S_out, int nthreads) {
    const int64_t P = p1 - p0, mid = N / 2;
    const int64_t F = (mode >= 2) ? M / 2 + 1 : M;
    fft_plan *plan = plan_make(M);
    int rc = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < batch; ++b) {
        const double *xb = x + b * n;
        cpx *S = (cpx *)S_out + b * F * P;
        /* x_slices: materialise all frames (src/lib.rs:612-652) */
        cpx *frames = (cpx *)calloc((size_t)(P * N), sizeof(cpx));
        cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)M);
        cpx *tmp = (cpx *)malloc(sizeof(cpx) * (size_t)(plan->pow2 ? M : plan->m));
        cpx *slices = (cpx *)malloc(sizeof(cpx) * (size_t)(P * F));
        if (!frames || !buf || !tmp || !slices) { rc = 1; free(frames); free(buf); free(tmp); free(slices); continue; }
        for (int64_t p = p0; p < p1; ++p) {
            const int64_t ks = p * H + k_offset - mid;
            cpx *fr = frames + (p - p0) * N;
            for (int64_t i = 0; i < N; ++i) {
                const int64_t s = ks + i;
                if (s >= 0 && s < n) fr[i].re = xb[s];
            }
        }
        for (int64_t t = 0; t < P; ++t) {
            const cpx *fr = frames + t * N;
            for (int64_t i = 0; i < N; ++i) { buf[i].re = fr[i].re * win[i]; buf[i].im = 0.0; }
            for (int64_t i = N; i < M; ++i) { buf[i].re = 0.0; buf[i].im = 0.0; }      /* resize */
            rotate_left(buf, M, ps, tmp);                                            /* :343-345 */
            fft_any(plan, buf, 0, tmp);
            cpx *o = slices + t * F;
            if (mode == 0) memcpy(o, buf, sizeof(cpx) * (size_t)M);
            else if (mode == 1) { rotate_left(buf, M, (M + 1) / 2, tmp); memcpy(o, buf, sizeof(cpx) * (size_t)M); }
            else {
                memcpy(o, buf, sizeof(cpx) * (size_t)F);
                o[0].im = 0.0;
                if (M % 2 == 0 && F > 1) o[F - 1].im = 0.0;
                if (mode == 3) {
                    const int64_t last = (M % 2 == 0) ? F - 2 : F - 1;
                    for (int64_t k = 1; k <= last; ++k) { o[k].re *= fac2x; o[k].im *= fac2x; }
                }
            }
        }
        for (int64_t t = 0; t < P; ++t)                 /* transpose (src/lib.rs:735-746) */
            for (int64_t f = 0; f < F; ++f) S[f * P + t] = slices[t * F + f];
        free(frames); free(buf); free(tmp); free(slices);
    }
    plan_free(plan);
    return rc;
}

/* istft: S [batch][F][P] -> y [batch][k1-k0]; q0, q1, p_min from the planner (src/lib.rs:842-847) */
int ref_istft(const double *S_in, int64_t P, int64_t batch, const double *dual, int64_t N, int64_t H,
              int64_t M, int mode, int64_t ps, double fac2x, int64_t p_min, int64_t k0, int64_t k1,
              int64_t q0, int64_t q1, double *y, int nthreads) {
    const int64_t mid = N / 2, len = k1 - k0;
    const int64_t F = (mode >= 2) ? M / 2 + 1 : M;
    fft_plan *plan = plan_make(M);
    int rc = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t b = 0; b < batch; ++b) {
        const cpx *S = (const cpx *)S_in + b * F * P;
        double *yb = y + b * len;
        cpx *buf = (cpx *)malloc(sizeof(cpx) * (size_t)M);
        cpx *tmp = (cpx *)malloc(sizeof(cpx) * (size_t)(plan->pow2 ? M : plan->m));
        if (!buf || !tmp) { rc = 1; free(buf); free(tmp); continue; }
        memset(yb, 0, sizeof(double) * (size_t)len);
        for (int64_t q = q0; q < q1; ++q) {
            const int64_t col = q - p_min;
            if (col < 0 || col >= P) continue;
            for (int64_t f = 0; f < F; ++f) buf[f] = S[f * P + col];          /* column gather */
            if (mode == 1) rotate_left(buf, M, M / 2, tmp);                   /* ifftshift */
            if (mode >= 2) {
                if (mode == 3) {
                    const int64_t last = (M % 2 == 0) ? F - 2 : F - 1;
                    for (int64_t k = 1; k <= last; ++k) { buf[k].re /= fac2x; buf[k].im /= fac2x; }
                }
                for (int64_t k = F; k < M; ++k) { buf[k].re = 0; buf[k].im = 0; }
                const int64_t me = (M / 2 < F - 1) ? M / 2 : F - 1;
                for (int64_t i = 1; i <= me; ++i) { buf[M - i].re = buf[i].re; buf[M - i].im = -buf[i].im; }
            }
            fft_any(plan, buf, 1, tmp);
            const double s = 1.0 / (double)M;
            for (int64_t i = 0; i < M; ++i) { buf[i].re *= s; buf[i].im *= s; }
            rotate_left(buf, M, (M - ps % M) % M, tmp);                      /* rotate_right(ps) */
            const int64_t i0 = q * H - mid;
            int64_t i1 = i0 + N;
            if (i1 > k0 + len) i1 = k0 + len;
            int64_t j0 = 0, a0 = i0;
            if (i0 < k0) { j0 = k0 - i0; a0 = k0; }
            for (int64_t i = a0; i < i1; ++i) {
                const int64_t j = j0 + (i - a0);
                yb[i - k0] += buf[j].re * dual[j];
            }
        }
        free(buf); free(tmp);
    }
    plan_free(plan);
    return rc;
}
