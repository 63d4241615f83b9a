// smooth_emu.cu -- HOST EMULATION of the mixed-radix kernel family (test infrastructure only): the phase
// functions of veils_b200/csrc/smooth_impl.cuh run on the CPU, thread by thread, phase by phase, in the
// order of the CUDA kernels in smooth.cu.  Not linked into libveils_cuda.so; not a fallback.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../veils_b200/csrc/smooth_impl.cuh"

using namespace veils;
using namespace veils::sm;
namespace veils { std::atomic<unsigned long long> g_launch_count{0}; }

template <typename T, int TF>
static int emu_fwd(const T *x, int64_t n, int64_t batch, const double *win, int64_t M, int64_t N, int64_t H, int64_t F,
                   int64_t ps, int64_t mid, int64_t p0, int64_t P, int64_t k_offset, int mode, double fac2x, int spec,
                   void *out, int64_t ldp) {
    FwdArgs<T> a;
    if (!make_shape(M, N, H, F, ps, &a.s)) return 2;
    a.lay = tab_layout<T>(a.s, false);
    std::vector<unsigned char> blob((size_t)a.lay.bytes);
    tab_fill<T>(a.s, false, win, blob.data());
    a.x = x; a.out = out; a.tab = blob.data();
    a.n = n; a.x_pitch = n; a.ldp = ldp; a.p0 = p0; a.P = P; a.k_offset = k_offset; a.ptiles = (P + TF - 1) / TF;
    a.mid = (int)mid; a.mode = mode; a.spectrogram = spec; a.xpad = xpad_of(H); a.fac2x = (T)fac2x;
    const int NT = 64;
    const int64_t span = (int64_t)(TF - 1) * H + N;
    std::vector<cx<vec_t<T>>> work((size_t)a.s.LC * TF / Vec<T>::W);
    std::vector<T> xin((size_t)(((span + H - 1) / H) * (H + a.xpad) + 4));
    FwdCtx<T> c;
    c.a = &a; c.tab = blob.data(); c.work = work.data(); c.xin = xin.data();
    for (int64_t blk = 0; blk < a.ptiles * batch; ++blk) {
        c.b = blk / a.ptiles; c.pt = blk % a.ptiles;
        for (auto &v : work) v = mk<vec_t<T>>(Vec<T>::bc((T)NAN), Vec<T>::bc((T)NAN));
        for (auto &v : xin) v = (T)NAN;
        for (int t = 0; t < NT; ++t) Fwd<T, TF>::stage(c, t, NT);
        for (int p = 0; p < a.s.np; ++p)
            for (int t = 0; t < NT; ++t) Fwd<T, TF>::pass(c, p, t, NT);
        for (int t = 0; t < NT; ++t) Fwd<T, TF>::store(c, t, NT);
    }
    return 0;
}

template <typename T, int TF>
static int emu_inv(const void *S, int64_t P, int64_t ldp, int64_t batch, const double *dual, int64_t M, int64_t N,
                   int64_t H, int64_t F, int64_t ps, int64_t mid, int64_t p_min, int64_t k0, int64_t k1, int64_t q0,
                   int64_t q1, int mode, double fac2x, T *xo, int64_t out_pitch) {
    InvArgs<T> a;
    if (!make_shape(M, N, H, F, ps, &a.s)) return 2;
    a.halo = (int)((N + H - 1) / H) - 1;
    if (a.halo >= TF) return 2;
    std::vector<double> d(N);
    for (int64_t i = 0; i < N; ++i) d[i] = dual[i] / ((double)M * (mode >= 2 ? 1.0 : 2.0));
    a.lay = tab_layout<T>(a.s, true);
    std::vector<unsigned char> blob((size_t)a.lay.bytes);
    tab_fill<T>(a.s, true, d.data(), blob.data());
    a.S = S; a.xo = xo; a.tab = blob.data();
    a.P = P; a.ldp = ldp; a.out_pitch = out_pitch; a.p_min = p_min; a.k0 = k0; a.k1 = k1; a.q0 = q0; a.q1 = q1;
    a.mid = (int)mid; a.mode = mode; a.ostride = ostride_of<T>(N, H); a.rfac2x = (T)(1.0 / fac2x);
    const int own = TF - a.halo;
    const int64_t qh0 = p2::fdiv64(k0 + mid, H), qh1 = p2::fdiv64(k1 - 1 + mid, H);
    a.ntiles = (qh1 - qh0 + 1 + own - 1) / own;
    a.qa0 = qh0 - a.halo;
    const int NT = 64;
    std::vector<cx<vec_t<T>>> work((size_t)a.s.LC * TF / Vec<T>::W);
    std::vector<T> ola((size_t)TF * a.ostride);
    std::vector<cx<T>> nyq((size_t)TF);
    InvCtx<T> c;
    c.a = &a; c.tab = blob.data(); c.work = work.data(); c.ola = ola.data(); c.nyq = nyq.data();
    for (int64_t blk = 0; blk < a.ntiles * batch; ++blk) {
        c.b = blk / a.ntiles; c.qa = a.qa0 + (blk % a.ntiles) * own;
        for (auto &v : work) v = mk<vec_t<T>>(Vec<T>::bc((T)NAN), Vec<T>::bc((T)NAN));
        for (auto &v : ola) v = (T)NAN;
        if (Inv<T, TF>::lands(a)) {
            for (auto &v : nyq) v = mk<T>((T)NAN, (T)NAN);
            for (int t = 0; t < NT; ++t) Inv<T, TF>::land(c, t, NT);
            for (int it = 0; it < Inv<T, TF>::merge_iters(a, NT); ++it) {
                std::vector<typename Inv<T, TF>::MReg> regs(NT);
                for (int t = 0; t < NT; ++t) Inv<T, TF>::merge_load(c, t, it, NT, regs[t]);
                for (int t = 0; t < NT; ++t) Inv<T, TF>::merge_store(c, t, it, NT, regs[t]);
            }
        } else {
            for (int t = 0; t < NT; ++t) Inv<T, TF>::load(c, t, NT);
        }
        for (int p = 0; p < a.s.np; ++p)
            for (int t = 0; t < NT; ++t) Inv<T, TF>::pass(c, p, t, NT);
        for (int t = 0; t < NT; ++t) Inv<T, TF>::gather(c, t, NT);
    }
    return 0;
}

extern "C" {
int smooth_emu_factors(int64_t M, int *np, int *R) {
    Factors f;
    if (!factorize(M, &f)) return 0;
    *np = f.np;
    for (int i = 0; i < kMaxPasses; ++i) R[i] = f.R[i];
    return 1;
}
// in-register DFT of length R (2 .. 16) on interleaved complex doubles
int smooth_emu_dft(int R, int inverse, double *re_im) {
    cx<double> w[16 * 1];
    if (R < 2 || R > 16) return 0;
    for (int i = 0; i < R; ++i) w[i] = mk<double>(re_im[2 * i], re_im[2 * i + 1]);
    PassIO<double> io;
    io.work = w; io.xin = nullptr; io.lt = nullptr; io.hp = 0; io.ola = nullptr; io.rt = nullptr; io.ostride = 0; io.half = 0;
    if (inverse) pass_any<double, 1, true, 0>(R, io, nullptr, 1, 0u, 1, 0, 1);
    else pass_any<double, 1, false, 0>(R, io, nullptr, 1, 0u, 1, 0, 1);
    for (int i = 0; i < R; ++i) { re_im[2 * i] = w[i].x; re_im[2 * i + 1] = w[i].y; }
    return 1;
}
int smooth_emu_forward(int prec, const void *x, int64_t n, int64_t batch, const double *win, int64_t M, int64_t N,
                       int64_t H, int64_t F, int64_t ps, int64_t mid, int64_t p0, int64_t P, int64_t k_offset, int mode,
                       double fac2x, int spec, void *out, int64_t ldp) {
    if (prec == 0) return emu_fwd<float, 8>((const float *)x, n, batch, win, M, N, H, F, ps, mid, p0, P, k_offset, mode, fac2x, spec, out, ldp);
    return emu_fwd<double, 4>((const double *)x, n, batch, win, M, N, H, F, ps, mid, p0, P, k_offset, mode, fac2x, spec, out, ldp);
}
int smooth_emu_inverse(int prec, const void *S, int64_t P, int64_t ldp, int64_t batch, const double *dual, int64_t M,
                       int64_t N, int64_t H, int64_t F, int64_t ps, int64_t mid, int64_t p_min, int64_t k0, int64_t k1,
                       int64_t q0, int64_t q1, int mode, double fac2x, void *xo, int64_t out_pitch) {
    if (prec == 0) return emu_inv<float, 8>(S, P, ldp, batch, dual, M, N, H, F, ps, mid, p_min, k0, k1, q0, q1, mode, fac2x, (float *)xo, out_pitch);
    return emu_inv<double, 8>(S, P, ldp, batch, dual, M, N, H, F, ps, mid, p_min, k0, k1, q0, q1, mode, fac2x, (double *)xo, out_pitch);
}
}
