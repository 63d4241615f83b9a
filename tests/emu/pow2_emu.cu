// pow2_emu.cu -- HOST EMULATION of the pow2 kernel family (test infrastructure only).
// Runs the very phase functions of veils_b200/csrc/pow2_tile_impl.cuh on the CPU, looping over
// blocks and threads with __syncthreads() replaced by phase boundaries, so that the index
// algebra of the CUDA kernels can be checked against the oracle in the CPU test suite.
// It is NOT linked into libveils_cuda.so and is not a fallback of the product.
#include <cstring>
#include <string>
#include <vector>

#include "../../veils_b200/csrc/plan.hpp"
#include "../../veils_b200/csrc/pow2_cfg.hpp"

using namespace veils;
using namespace veils::p2;

namespace veils { std::atomic<unsigned long long> g_launch_count{0}; }

static thread_local std::string e_err;

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
        const bool vec = fwd_vec(prm.ps, prm.H, prm.N);
        const TileLayout lay = tile_layout(prm.H, vec);
        std::vector<cx<T>> work((size_t)K::MC * TF);
        std::vector<T> xin((size_t)fwd_xin_elems(prm.N, prm.H, TF, lay));
        std::vector<double> twd((size_t)twiddle_count(prm.M));
        twiddle_fill(prm.M, twd.data());
        std::vector<T> tw = conv<T>(twd), win = conv<T>(pl->win, 0.5);
        FwdCtx<T> c;
        c.x = (const T *)x; c.out = out; c.prm = prm; c.win = win.data(); c.tw = tw.data();
        c.work = work.data(); c.xin = xin.data(); c.lay = lay; c.vec = vec;
        const int64_t ptiles = (prm.P + TF - 1) / TF;
        for (int64_t blk = 0; blk < ptiles * prm.batch; ++blk) {
            c.b = blk / ptiles; c.pt = blk % ptiles;
            for (int t = 0; t < K::NT; ++t) K::fill(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass1(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass_mid(c, t);
            for (int t = 0; t < K::NT; ++t) K::pass_last(c, t);
        }
        return 0;
    }
};

struct EmuInv {
    const void *S; void *xo; InvParams prm; const Plan *pl;
    template <typename T, int LOGMC, int TF, int NSUB>
    int run() {
        using K = Inv<T, LOGMC, TF, NSUB>;
        const int halo = inv_halo(prm.N, prm.H), own = TF - halo;
        const int ostride = inv_ostride(prm.N);
        size_t elems = (size_t)K::MC * TF * 2;
        if ((size_t)TF * ostride > elems) elems = (size_t)TF * ostride;
        std::vector<T> buf(elems);
        std::vector<double> twd((size_t)twiddle_count(prm.M));
        twiddle_fill(prm.M, twd.data());
        const double sc = 1.0 / ((double)prm.M * (prm.mode >= 2 ? 1.0 : 2.0));
        std::vector<T> tw = conv<T>(twd), dual = conv<T>(pl->dual, sc);
        InvCtx<T> c;
        c.S = S; c.xo = (T *)xo; c.prm = prm; c.dual = dual.data(); c.tw = tw.data();
        c.work = (cx<T> *)buf.data(); c.ola = buf.data(); c.halo = halo; c.ostride = ostride;
        const int64_t qh0 = floordiv(prm.k0 + prm.mid, prm.H), qh1 = floordiv(prm.k1 - 1 + prm.mid, prm.H);
        const int64_t ntiles = (qh1 - qh0 + 1 + own - 1) / own;
        std::vector<cx<T>> regs((size_t)K::NT * K::GPT * K::RL);
        for (int64_t blk = 0; blk < ntiles * prm.batch; ++blk) {
            c.b = blk / ntiles; c.qa = qh0 - halo + (blk % ntiles) * own;
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
            for (int t = 0; t < K::NT; ++t) K::gather(c, t);
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
             int spectrogram) {
    Plan pl;
    std::string e = plan_init(pl, win, m_num, hop, fs, mode, mfft, nullptr, scale_to, phase_shift, precision, 0);
    if (!e.empty()) { e_err = e; return 1; }
    int64_t a, b;
    e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) { e_err = e; return 1; }
    if (!supported(pl.M, pl.N, pl.H, pl.ps, precision, false)) return 2;
    FwdParams prm;
    prm.n = n; prm.x_pitch = n; prm.batch = batch; prm.N = pl.N; prm.H = pl.H; prm.M = pl.M;
    prm.F = pl.F; prm.mid = pl.mid; prm.ps = pl.ps; prm.p0 = a; prm.P = b - a; prm.k_offset = k_offset;
    prm.ldp = b - a; prm.mode = pl.mode; prm.fac2x = pl.fac2x; prm.spectrogram = spectrogram;
    Cfg c;
    cfg_for(pl.M, precision, &c);
    EmuFwd f{x, out, prm, &pl};
    return dispatch(precision, c.logmc, f);
}

int emu_istft(const double *win, int64_t m_num, int64_t hop, double fs, const char *mode, int64_t mfft,
              const char *scale_to, int64_t phase_shift, int precision, const void *S, int64_t P,
              int64_t batch, const int64_t *k0, const int64_t *k1, void *xo) {
    Plan pl;
    std::string e = plan_init(pl, win, m_num, hop, fs, mode, mfft, nullptr, scale_to, phase_shift, precision, 0);
    if (!e.empty()) { e_err = e; return 1; }
    IstftRange r;
    e = plan_istft_range(pl, P, k0, k1, &r);
    if (e.empty()) e = plan_dual(pl);
    if (!e.empty()) { e_err = e; return 1; }
    if (!supported(pl.M, pl.N, pl.H, pl.ps, precision, true)) return 2;
    InvParams prm;
    prm.P = P; prm.ldp = P; prm.batch = batch; prm.N = pl.N; prm.H = pl.H; prm.M = pl.M; prm.F = pl.F;
    prm.mid = pl.mid; prm.ps = pl.ps; prm.p_min = pl.p_min; prm.k0 = r.k0; prm.k1 = r.k1; prm.q0 = r.q0;
    prm.q1 = r.q1; prm.out_pitch = r.k1 - r.k0; prm.mode = pl.mode; prm.fac2x = pl.fac2x;
    Cfg c;
    cfg_for(pl.M, precision, &c);
    EmuInv f{S, xo, prm, &pl};
    return dispatch(precision, c.logmc, f);
}
}
