// pow2_emu.cu -- HOST EMULATION of the pow2 kernel family (test infrastructure only).
// Runs the very phase functions of veils_b200/csrc/pow2_tile_impl.cuh on the CPU, looping over
// blocks and threads with __syncthreads() replaced by phase boundaries, so that the index
// algebra of the CUDA kernels can be checked against the oracle in the CPU test suite.
// It is NOT linked into libveils_cuda.so and is not a fallback of the product.
#include <cmath>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../veils_b200/csrc/plan.hpp"
#include "../../veils_b200/csrc/pow2_cfg.hpp"
#include "../../veils_b200/csrc/pow2_pk_cfg.hpp"

using namespace veils;
using namespace veils::p2;

namespace veils { std::atomic<unsigned long long> g_launch_count{0}; }

static thread_local std::string e_err;
static int64_t emu_grid_hint = 2;   // small -> several tiles per chunk even for short test signals

template <typename T>
static std::vector<T> conv(const std::vector<double> &s, double scale = 1.0) {
    std::vector<T> r(s.size());
    for (size_t i = 0; i < s.size(); ++i) r[i] = (T)(s[i] * scale);
    return r;
}

struct EmuFwd {
    const void *x; void *out; FwdParams prm; const Plan *pl;
    template <typename T, int LOGMC, int TF, int NSUB>
    int run() {
        using K = Fwd<T, LOGMC, TF, NSUB>;
        std::vector<double> twd((size_t)twiddle_count(prm.M));
        twiddle_fill(prm.M, twd.data());
        std::vector<T> tw = conv<T>(twd), win = conv<T>(pl->win, 0.5);
        FwdCtx<T> c;
        c.a = make_fwd_args<T>((const T *)x, out, prm, win.data(), tw.data(), TF);
        std::vector<cx<T>> work((size_t)K::MCP * TF);
        std::vector<T> xin((size_t)fwd_xin_elems(prm.N, prm.H, TF, c.a.lay));
        c.work = work.data(); c.xin = xin.data(); c.tw = tw.data(); c.win = win.data();
        for (int64_t blk = 0; blk < c.a.ptiles * prm.batch; ++blk) {
            c.b = blk / c.a.ptiles; c.pt = blk % c.a.ptiles;
            for (int t = 0; t < K::NT; ++t) K::fill(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass1(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass_mid(c, t);
            if (K::UNPAIRED) {
                for (int t = 0; t < K::NT; ++t) K::last_a(c, t);
                for (int t = 0; t < K::NT; ++t) K::last_b(c, t);
            } else {
                for (int t = 0; t < K::NT; ++t) K::pass_last(c, t);
            }
        }
        return 0;
    }
};

struct EmuInv {
    const void *S; void *xo; InvParams prm; const Plan *pl;
    template <typename T, int LOGMC, int TF, int NSUB>
    int run() {
        using K = Inv<T, LOGMC, TF, NSUB>;
        std::vector<double> twd((size_t)twiddle_count(prm.M));
        twiddle_fill(prm.M, twd.data());
        const double sc = 1.0 / ((double)prm.M * (prm.mode >= 2 ? 1.0 : 2.0));
        std::vector<T> tw = conv<T>(twd), dual = conv<T>(pl->dual, sc);
        InvCtx<T> c;
        c.a = make_inv_args<T>(S, (T *)xo, prm, dual.data(), tw.data(), TF);
        size_t elems = (size_t)K::MCP * TF * 2;
        if ((size_t)TF * c.a.ostride > elems) elems = (size_t)TF * c.a.ostride;
        std::vector<T> buf(elems + 4);
        c.work = (cx<T> *)buf.data(); c.ola = buf.data(); c.tw = tw.data(); c.dual = dual.data();
        std::vector<cx<T>> regs((size_t)K::NT * K::GPT * K::RL);
        tile_inv_chunks<T>(c.a, prm, TF, prm.batch, emu_grid_hint);
        std::vector<T> carry0(inv_carry_elems(prm.N, prm.H) + 1, (T)NAN), carry1(inv_carry_elems(prm.N, prm.H) + 1, (T)NAN);
        const int64_t lc = (int64_t)TF * c.a.chunk_tiles - c.a.halo;
        for (int64_t ch = 0; ch < c.a.cps * prm.batch; ++ch) {
            c.b = ch / c.a.cps;
            const int64_t a0 = c.a.qh0 + (ch - c.b * c.a.cps) * lc;
            c.kown = a0 * (int64_t)c.a.H - c.a.mid;
            int flip = 0;
            for (int tl = 0; tl < c.a.chunk_tiles; ++tl) {
                c.qa = a0 - c.a.halo + (int64_t)tl * TF;
                if (c.qa > c.a.qh1) break;
                c.first = tl == 0;
                c.last = tl + 1 >= c.a.chunk_tiles || c.qa + TF > c.a.qh1;
                c.carry_in = flip ? carry1.data() : carry0.data();
                c.carry_out = flip ? carry0.data() : carry1.data();
                flip ^= 1;
                for (int t = 0; t < K::NT; ++t) K::pass1(c, t);
                for (int t = 0; t < K::NT; ++t) K::pass_mid(c, t);
                for (int t = 0; t < K::NT; ++t) {
                    cx<T> r[K::GPT][K::RL];
                    K::last_load(c, t, r);
                    std::memcpy(&regs[(size_t)t * K::GPT * K::RL], r, sizeof(r));
                }
                for (int t = 0; t < K::NT; ++t) {
                    cx<T> r[K::GPT][K::RL];
                    std::memcpy(r, &regs[(size_t)t * K::GPT * K::RL], sizeof(r));
                    K::last_store(c, t, r);
                }
                for (int t = 0; t < K::NT; ++t) K::gather_s(c, t);
            }
        }
        return 0;
    }
};

struct EmuPkFwd {
    const void *x; void *out; FwdParams prm; const Plan *pl;
    template <int LOGMC, int TF, int NSUB>
    int run() {
        using K = PkFwd<LOGMC, TF, NSUB>;
        std::vector<double> twd((size_t)p2::pk_twiddle_count(prm.M));
        p2::pk_twiddle_fill(prm.M, false, twd.data());
        std::vector<float> tw = conv<float>(twd);
        tw.resize(tw.size() + 8);
        std::vector<PTab> ptab((size_t)K::MC);
        pk_ptab_fill(prm.M, prm.N, prm.H, prm.ps, pl->win.data(), TF, ptab.data());
        PkFwdCtx c;
        c.a = make_fwd_args<float>((const float *)x, out, prm, nullptr, tw.data(), TF);
        std::vector<fl4> work((size_t)K::MC * 2 * TF / 4 + TF / 2 + 4);      // + staged Nyquist row
        std::vector<fl4> xin((size_t)fwd_xin_elems(prm.N, prm.H, TF, c.a.lay) / 4 + 4);
        c.work = (float *)work.data(); c.xin = (float *)xin.data(); c.tw = tw.data(); c.ptab = ptab.data();
        const int zo = pk_zero_off(prm.N, prm.H, TF);
        std::vector<C2> regs((size_t)K::NT * K::RL);
        for (int64_t blk = 0; blk < c.a.ptiles * prm.batch; ++blk) {
            c.b = blk / c.a.ptiles; c.pt = blk % c.a.ptiles;
            // poison the tile (NaN) except the zero pair: padded entries must never read it
            for (size_t i = 0; i < xin.size() * 4; ++i) c.xin[i] = std::nanf("");
            c.xin[zo] = c.xin[zo + 1] = 0.f;
            for (int t = 0; t < K::NT; ++t) K::fill(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass1(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass_mid(c, t);
            if (prm.mode >= 2) {
                // staged path of pk_fwd_kernel (one-sided layouts): paired last pass writes the finished
                // bins over its own work rows; the TMA boxes are replayed here as clipped row copies
                const int esz = prm.spectrogram ? 4 : 8;
                for (int i = 0; i < K::WPT2; ++i) {
                    for (int t = 0; t < K::NT; ++t) {
                        C2 u[K::RL];
                        K::last_dft2(c, t, i, u);
                        std::memcpy(&regs[(size_t)t * K::RL], u, sizeof(u));
                    }
                    for (int t = 0; t < K::NT; ++t) {
                        C2 u[K::RL];
                        std::memcpy(u, &regs[(size_t)t * K::RL], sizeof(u));
                        if (prm.spectrogram) K::template last_emit2<true>(c, t, i, u);
                        else K::template last_emit2<false>(c, t, i, u);
                    }
                    for (int wg = i * K::NW2; wg < (i + 1) * K::NW2 && wg < K::G / 2; ++wg) {
                        for (int h = 0; h < 2; ++h) {
                            const int g = K::group_of(wg, h);
                            for (int d = 0; d < K::RL + (g == 0 ? 1 : 0); ++d) {
                                const int f = g + K::G * d;
                                const char *src = (const char *)c.work + (d == K::RL ? (size_t)K::MC * TF * 8
                                                  : (size_t)K::GEO::goff(g) * TF * 8 + (size_t)d * TF * esz);
                                for (int sl = 0; sl < TF; ++sl) {
                                    const int64_t pcol = c.pt * TF + sl;
                                    if (pcol >= prm.P) continue;
                                    char *dst = (char *)out + (((size_t)c.b * prm.F + f) * prm.ldp + pcol) * esz;
                                    std::memcpy(dst, src + (size_t)sl * esz, esz);
                                }
                            }
                        }
                    }
                }
            } else
            for (int i = 0; i < K::WPT; ++i) {
                for (int t = 0; t < K::NT; ++t) {
                    C2 u[K::RL];
                    K::last_dft(c, t, i, u);
                    std::memcpy(&regs[(size_t)t * K::RL], u, sizeof(u));
                }
                for (int t = 0; t < K::NT; ++t) {
                    C2 u[K::RL];
                    std::memcpy(u, &regs[(size_t)t * K::RL], sizeof(u));
                    K::last_emit(c, t, i, u, &regs[(size_t)(t ^ K::LN) * K::RL]);
                }
            }
        }
        return 0;
    }
};

struct EmuPkInv {
    const void *S; void *xo; InvParams prm; const Plan *pl;
    template <int LOGMC, int TF, int NSUB>
    int run() {
        using K = PkInv<LOGMC, TF, NSUB>;
        std::vector<double> twd((size_t)p2::pk_twiddle_count(prm.M));
        p2::pk_twiddle_fill(prm.M, true, twd.data());
        const double sc = 1.0 / ((double)prm.M * (prm.mode >= 2 ? 1.0 : 2.0));
        std::vector<float> tw = conv<float>(twd);
        std::vector<double> ds(pl->dual);
        for (double &v : ds) v *= sc;
        std::vector<ITab> itab((size_t)K::MC);
        pk_itab_fill(prm.M, prm.N, prm.ps, ds.data(), itab.data());
        PkInvCtx c;
        c.a = make_inv_args<float>(S, (float *)xo, prm, nullptr, tw.data(), TF);
        c.a.ostride = pk_ostride(prm.N, prm.ps);
        c.a.pair = pk_pair(prm.N, prm.ps) ? 1 : 0;
        size_t elems = (size_t)K::MC * 2 * TF;
        if ((size_t)TF * c.a.ostride > elems) elems = (size_t)TF * c.a.ostride;
        std::vector<fl4> buf(elems / 4 + 4);
        c.work = (float *)buf.data(); c.ola = (float *)buf.data(); c.tw = tw.data(); c.itab = itab.data();
        std::vector<C2> regs((size_t)K::NT * K::GPT * K::RL);
        std::vector<Raw> raws((size_t)K::NT * (K::R1 + 1));
        std::vector<float> carry0((size_t)prm.N + 4), carry1((size_t)prm.N + 4);
        // same conditions as PkInvLaunch::run (pow2_pk.cu) apart from the TMA / alignment ones
        const bool acc_path = prm.mode >= 2 && c.a.halo == 3 && prm.N == prm.M && 4 * prm.H == prm.N &&
                              prm.ps == prm.N / 2 && c.a.pair && prm.k0 % 2 == 0 && K::WPT2 == 1 &&
                              (LOGMC == 7 || LOGMC == 8 || LOGMC == 10 || LOGMC == 11);
        std::vector<float> acc((size_t)K::ACC_FLOATS + 8);
        pk_inv_chunks(c.a, prm, TF, prm.batch, emu_grid_hint);
        const int64_t lc = (int64_t)TF * c.a.chunk_tiles - c.a.halo;
        for (int64_t ch = 0; ch < c.a.cps * prm.batch; ++ch) {
            c.b = ch / c.a.cps;
            const int64_t a0 = c.a.qh0 + (ch - c.b * c.a.cps) * lc;
            int flip = 0;
            // poison the carries: a chunk must never read what another chunk left behind
            for (auto &v : carry0) v = std::nanf("");
            for (auto &v : carry1) v = std::nanf("");
            for (int t = 0; t < c.a.chunk_tiles; ++t) {
                c.qa = a0 - c.a.halo + (int64_t)t * TF;
                if (c.qa > c.a.qh1) break;
                c.first = t == 0;
                c.carry_in = flip ? carry1.data() : carry0.data();
                c.carry_out = flip ? carry0.data() : carry1.data();
                flip ^= 1;
                if (prm.mode >= 2) {           // paired pass 1 of pk_inv_kernel (one-sided layouts)
                    for (int i = 0; i < K::WPT2; ++i)
                        for (int tt = 0; tt < K::NT; ++tt) K::p1p(c, tt, i);
                } else
                for (int i = 0; i < K::WPT; ++i) {
                    for (int tt = 0; tt < K::NT; ++tt) {
                        Raw xr[K::R1]; Raw xn;
                        std::memset(xr, 0, sizeof(xr)); std::memset(&xn, 0, sizeof(xn));
                        K::p1_load(c, tt, i, xr, xn);
                        std::memcpy(&raws[(size_t)tt * (K::R1 + 1)], xr, sizeof(xr));
                        raws[(size_t)tt * (K::R1 + 1) + K::R1] = xn;
                    }
                    for (int tt = 0; tt < K::NT; ++tt) {
                        Raw xr[K::R1];
                        std::memcpy(xr, &raws[(size_t)tt * (K::R1 + 1)], sizeof(xr));
                        K::p1_rest(c, tt, i, xr, raws[(size_t)tt * (K::R1 + 1) + K::R1],
                                   &raws[(size_t)(tt ^ K::LN) * (K::R1 + 1)]);
                    }
                }
                for (int tt = 0; tt < K::NT; ++tt) K::pass_mid(c, tt);
                for (int tt = 0; tt < K::NT; ++tt) {
                    C2 r[K::GPT][K::RL];
                    K::last_load(c, tt, r);
                    std::memcpy(&regs[(size_t)tt * K::GPT * K::RL], r, sizeof(r));
                }
                if (acc_path) {
                    // accumulator overlap-add of pk_inv_kernel (75 % overlap, default roll): four quarter
                    // passes, output, tail -> head.  The accumulator is NaN-poisoned at every chunk start:
                    // a tile must only ever read what this chunk has written.
                    if (c.first) {
                        for (auto &v : acc) v = std::nanf("");
                        for (int tt = 0; tt < K::NT; ++tt) K::ola_zero_head(tt, acc.data());
                    }
                    auto pass = [&](auto q) {
                        for (int tt = 0; tt < K::NT; ++tt) {
                            C2 r[K::GPT][K::RL];
                            std::memcpy(r, &regs[(size_t)tt * K::GPT * K::RL], sizeof(r));
                            K::template ola_acc<decltype(q)::value>(c, tt, r, acc.data());
                        }
                    };
                    pass(std::integral_constant<int, 3>());
                    pass(std::integral_constant<int, 2>());
                    pass(std::integral_constant<int, 1>());
                    pass(std::integral_constant<int, 0>());
                    std::vector<typename K::Tail> tails((size_t)K::NT);
                    for (int tt = 0; tt < K::NT; ++tt) {
                        K::ola_out(c, tt, acc.data());
                        tails[(size_t)tt] = K::ola_take_tail(tt, acc.data());
                    }
                    for (int tt = 0; tt < K::NT; ++tt) K::ola_keep_tail(tt, acc.data(), tails[(size_t)tt]);
                    continue;
                }
                for (int tt = 0; tt < K::NT; ++tt) {
                    C2 r[K::GPT][K::RL];
                    std::memcpy(r, &regs[(size_t)tt * K::GPT * K::RL], sizeof(r));
                    K::last_store(c, tt, r);
                }
                for (int tt = 0; tt < K::NT; ++tt) K::gather(c, tt);
            }
        }
        return 0;
    }
};

extern "C" {

const char *emu_last_error() { return e_err.c_str(); }

// returns 0 ok, 1 invalid args (message), 2 configuration not supported by the pow2 family
int emu_stft(const double *win, int64_t m_num, int64_t hop, double fs, const char *mode, int64_t mfft,
             const char *scale_to, int64_t phase_shift, int precision, const void *x, int64_t n,
             int64_t batch, const int64_t *p0, const int64_t *p1, int64_t k_offset, void *out,
             int spectrogram, int family) {
    Plan pl;
    std::string e = plan_init(pl, win, m_num, hop, fs, mode, mfft, nullptr, scale_to, phase_shift, precision, 0);
    if (!e.empty()) { e_err = e; return 1; }
    int64_t a, b;
    e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) { e_err = e; return 1; }
    if (family == 1 ? !(precision == 0 && p2::pk_supported(pl.M, pl.N, pl.H, pl.ps, false))
                    : !supported(pl.M, pl.N, pl.H, pl.ps, precision, false)) return 2;
    FwdParams prm;
    prm.n = n; prm.x_pitch = n; prm.batch = batch; prm.N = pl.N; prm.H = pl.H; prm.M = pl.M;
    prm.F = pl.F; prm.mid = pl.mid; prm.ps = pl.ps; prm.p0 = a; prm.P = b - a; prm.k_offset = k_offset;
    prm.ldp = b - a; prm.mode = pl.mode; prm.fac2x = pl.fac2x; prm.spectrogram = spectrogram;
    Cfg c;
    cfg_for(pl.M, precision, &c);
    if (family == 1) {
        EmuPkFwd f{x, out, prm, &pl};
        return pk_dispatch(c.logmc, f);
    }
    EmuFwd f{x, out, prm, &pl};
    return dispatch(precision, c.logmc, f);
}

int emu_istft(const double *win, int64_t m_num, int64_t hop, double fs, const char *mode, int64_t mfft,
              const char *scale_to, int64_t phase_shift, int precision, const void *S, int64_t P,
              int64_t batch, const int64_t *k0, const int64_t *k1, void *xo, int family) {
    Plan pl;
    std::string e = plan_init(pl, win, m_num, hop, fs, mode, mfft, nullptr, scale_to, phase_shift, precision, 0);
    if (!e.empty()) { e_err = e; return 1; }
    IstftRange r;
    e = plan_istft_range(pl, P, k0, k1, &r);
    if (e.empty()) e = plan_dual(pl);
    if (!e.empty()) { e_err = e; return 1; }
    if (family == 1 ? !(precision == 0 && p2::pk_supported(pl.M, pl.N, pl.H, pl.ps, true))
                    : !supported(pl.M, pl.N, pl.H, pl.ps, precision, true)) return 2;
    InvParams prm;
    prm.P = P; prm.ldp = P; prm.batch = batch; prm.N = pl.N; prm.H = pl.H; prm.M = pl.M; prm.F = pl.F;
    prm.mid = pl.mid; prm.ps = pl.ps; prm.p_min = pl.p_min; prm.k0 = r.k0; prm.k1 = r.k1; prm.q0 = r.q0;
    prm.q1 = r.q1; prm.out_pitch = r.k1 - r.k0; prm.mode = pl.mode; prm.fac2x = pl.fac2x;
    Cfg c;
    cfg_for(pl.M, precision, &c);
    if (family == 1) {
        EmuPkInv f{S, xo, prm, &pl};
        return pk_dispatch(c.logmc, f);
    }
    EmuInv f{S, xo, prm, &pl};
    return dispatch(precision, c.logmc, f);
}
}
