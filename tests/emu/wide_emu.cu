// wide_emu.cu -- HOST EMULATION of the "wide" kernel family (test infrastructure only).
// Runs the phase functions of veils_b200/csrc/wide_impl.cuh on the CPU, thread by thread with
// __syncthreads() replaced by phase boundaries, in the order the CUDA kernels of wide.cu run them,
// so the index algebra is checked against the oracle without a GPU.  Every shared-memory access is
// traced and replayed against the 32-bank model (8-byte accesses: one wavefront per half-warp,
// 16-byte: per quarter-warp), which gives the bank-conflict count per phase.
// Not linked into libveils_cuda.so; not a fallback of the product.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#define VEILS_WIDE_TRACE 1
#include "../../veils_b200/csrc/wide_impl.cuh"

using namespace veils;
using namespace veils::wd;

namespace veils { std::atomic<unsigned long long> g_launch_count{0}; }

namespace {
struct Acc { int tid, seq, kind, bytes; uintptr_t addr; };
std::vector<Acc> g_trace;
int g_tid = 0, g_seq = 0;
bool g_tracing = false;
const char *g_lo = nullptr, *g_hi = nullptr;     // shared-memory image: addresses are relative to g_lo
long long g_wavefronts[8], g_ideal[8];            // per phase id
int g_phase = 0;

void analyse() {
    std::map<std::pair<int, int>, std::vector<const Acc *>> byinst;
    for (const Acc &a : g_trace) byinst[{a.tid / 32, a.seq}].push_back(&a);
    for (auto &kv : byinst) {
        const int bytes = kv.second[0]->bytes;
        const int group = bytes == 16 ? 8 : (bytes == 8 ? 16 : 32);   // lanes per wavefront
        for (int g0 = 0; g0 < 32; g0 += group) {
            std::map<int, std::vector<uintptr_t>> banks;
            bool any = false;
            for (const Acc *a : kv.second) {
                const int lane = a->tid % 32;
                if (lane < g0 || lane >= g0 + group) continue;
                any = true;
                for (int w = 0; w < a->bytes / 4; ++w) {
                    const uintptr_t word = a->addr / 4 + w;
                    auto &v = banks[(int)(word % 32)];
                    if (std::find(v.begin(), v.end(), word) == v.end()) v.push_back(word);
                }
            }
            if (!any) continue;
            size_t worst = 1;
            for (auto &b : banks) worst = std::max(worst, b.second.size());
            g_wavefronts[g_phase] += (long long)worst;
            g_ideal[g_phase] += 1;
        }
    }
    g_trace.clear();
}
}  // namespace

namespace veils { namespace wd {
void wide_trace(int kind, const void *p, int bytes) {
    if (!g_tracing) return;
    const char *q = (const char *)p;
    if (q < g_lo || q >= g_hi) return;
    g_trace.push_back(Acc{g_tid, g_seq++, kind, bytes, (uintptr_t)(q - g_lo)});
}
}}

template <typename F> static void phase(int id, int nt, F f) {
    g_phase = id;
    for (int t = 0; t < nt; ++t) { g_tid = t; g_seq = 0; f(t); }
    if (g_tracing) analyse();
}

// pass 2 is in place with a permutation inside warp-owned groups: __syncwarp() separates the loads of
// all lanes from the stores; here: all threads load, then all threads store, per iteration
template <typename K, bool INV, int RA, int RB, int LB>
static void emu_pass2(int id, float *work, const fl2 *tw2) {
    std::vector<C2> regs((size_t)K::NT * RB);
    for (int it = 0; it < K::GEO::template pass2_iters<RA, LB>(); ++it) {
        g_phase = id;
        for (int t = 0; t < K::NT; ++t) {
            g_tid = t; g_seq = 0;
            const int u = t / K::LN + it * K::NSUB;
            if (u >= RA * LB) continue;
            C2 v[RB];
            K::GEO::template pass2_load<INV, RA, RB, LB>(work, u, t % K::LN, v);
            std::memcpy(&regs[(size_t)t * RB], v, sizeof(v));
        }
        if (g_tracing) analyse();
        for (int t = 0; t < K::NT; ++t) {
            g_tid = t; g_seq = 0;
            const int u = t / K::LN + it * K::NSUB;
            if (u >= RA * LB) continue;
            C2 v[RB];
            std::memcpy(v, &regs[(size_t)t * RB], sizeof(v));
            K::GEO::template pass2_store<INV, RA, RB, LB>(work, tw2, u, t % K::LN, v);
        }
        if (g_tracing) analyse();
    }
}

static std::vector<fl2> make_tables(int64_t M, int r1, int r2, const double *vec, double scale) {
    std::vector<double> d((size_t)(2 * wide_table_entries(M)));
    wide_table_fill(M, r1, r2, vec, scale, d.data());
    std::vector<fl2> t((size_t)wide_table_entries(M));
    for (size_t i = 0; i < t.size(); ++i) { t[i].x = (float)d[2 * i]; t[i].y = (float)d[2 * i + 1]; }
    return t;
}

struct EmuFwd {
    const float *x; void *out; const double *win;
    int64_t n, batch, x_pitch, p0, P, k_offset, ldp, F;
    int mode, spectrogram; double fac2x;
    template <int LOGMC, int TF>
    int run() {
        using K = WFwd<LOGMC, TF>;
        std::vector<fl2> tab = make_tables(K::M, K::R1, K::R2, win, 0.5);
        const WTabOff o = wide_table_offsets(K::M);
        WFwdCtx c;
        c.a.x = x; c.a.out = out;
        c.a.tab = tab.data();
        c.wtab = tab.data() + o.xtab; c.tw1 = tab.data() + o.tw1; c.tw2 = tab.data() + o.tw2; c.twp = tab.data() + o.twp;
        c.a.n = n; c.a.x_pitch = x_pitch; c.a.ldp = ldp; c.a.p0 = p0; c.a.P = P; c.a.k_offset = k_offset;
        c.a.ptiles = (P + TF - 1) / TF; c.a.F = F; c.a.mode = mode; c.a.spectrogram = spectrogram;
        const int esz = spectrogram ? 4 : 8;
        c.a.vec16 = ((ldp * esz) % 16 == 0 && ((uintptr_t)out % 16) == 0) ? 1 : 0;
        c.a.fac2x = (float)fac2x;
        c.a.no_bulk = 1;
        std::vector<fl4> smem((wide_work_floats<LOGMC, TF>() + K::XIN_FLOATS) / 4 + 4);
        c.work = (float *)smem.data();
        c.xin = c.work + wide_work_floats<LOGMC, TF>();
        g_lo = (const char *)smem.data(); g_hi = g_lo + smem.size() * 16;
        for (int64_t blk = 0; blk < c.a.ptiles * batch; ++blk) {
            c.b = blk / c.a.ptiles; c.pt = blk % c.a.ptiles;
            for (size_t i = 0; i < smem.size() * 4; ++i) c.work[i] = std::nanf("");    // poison
            g_tracing = blk == 0;
            for (int t = 0; t < K::NT; ++t) K::fill(c, t);
            phase(0, K::NT, [&](int t) { K::pass1(c, t); });
            emu_pass2<K, false, K::R1, K::R2, K::L2>(1, c.work, c.tw2);
            phase(2, K::NT, [&](int t) { K::pass3(c, t); });
        }
        g_tracing = false;
        return 0;
    }
};

struct EmuInv {
    const void *S; float *xo; const double *dual;
    int64_t P, ldp, batch, out_pitch, p_min, k0, k1, q0, q1, F;
    int mode, chunk_tiles; double fac2x; int use_land;
    template <int LOGMC, int TF>
    int run() {
        using K = WInv<LOGMC, TF>;
        const double sc = 1.0 / ((double)K::M * (mode >= 2 ? 1.0 : 2.0));
        std::vector<fl2> tab = make_tables(K::M, K::R1, K::R2, dual, sc);
        const WTabOff o = wide_table_offsets(K::M);
        WInvCtx c;
        WInvArgs &a = c.a;
        a.S = S; a.xo = xo;
        a.tab = tab.data();
        c.dtab = tab.data() + o.xtab; c.tw1 = tab.data() + o.tw1; c.tw2 = tab.data() + o.tw2; c.twp = tab.data() + o.twp;
        a.P = P; a.ldp = ldp; a.out_pitch = out_pitch; a.p_min = p_min; a.k0 = k0; a.k1 = k1; a.q0 = q0; a.q1 = q1; a.F = F;
        a.mode = mode; a.rfac2x = (float)(1.0 / fac2x); a.prefetch = 0;
        a.vec16 = (ldp % 2 == 0 && ((uintptr_t)S % 16) == 0) ? 1 : 0;
        a.vec_out = ((uintptr_t)xo % 16) == 0 ? 1 : 0;
        a.qh0 = fdiv64(k0 + K::MID, K::H);
        a.qh1 = fdiv64(k1 - 1 + K::MID, K::H);
        a.chunk_tiles = chunk_tiles;
        const int64_t lc = (int64_t)TF * chunk_tiles - 4, Q = a.qh1 - a.qh0 + 1;
        a.cps = (Q + lc - 1) / lc;
        std::vector<fl4> smem((wide_work_floats<LOGMC, TF>() + K::XIN_FLOATS) / 4 + 4);
        c.work = (float *)smem.data();
        c.acc = c.work + wide_work_floats<LOGMC, TF>();
        g_lo = (const char *)smem.data(); g_hi = g_lo + smem.size() * 16;
        std::vector<typename K::Regs> regs((size_t)K::NT);
        std::vector<typename K::Tail> tails((size_t)K::NT);
        for (int64_t ch = 0; ch < a.cps * batch; ++ch) {
            const WChunk wc = wide_chunk(a, ch, TF, K::H, K::MID);
            c.b = wc.b; c.halo = wc.halo; c.kend = wc.kend;
            for (size_t i = 0; i < smem.size() * 4; ++i) c.work[i] = std::nanf("");    // poison
            for (int t = 0; t < chunk_tiles; ++t) {
                c.qa = wc.qa0 + (int64_t)t * TF;
                if (c.qa > a.qh1) break;
                if (((c.qa - p_min) & 1) != 0) return -7;                              // tiles start on even columns
                c.first = t == 0;
                c.more = 0;
                g_tracing = ch == 0 && t == 1;
                if (use_land && mode >= 2 && a.vec16) {
                    // what the TMA engine does (wide_inv_kernel): NH column blocks [Mc rows][8 slices] + the
                    // Nyquist rows land over the work buffer, columns outside [0, P) as zeros
                    const int64_t col0 = c.qa - p_min;
                    const float *Sf = (const float *)S;
                    for (int h = 0; h < K::NH; ++h)
                        for (int k = 0; k <= K::MC; ++k)
                            for (int i = 0; i < 8; ++i) {
                                const int64_t col = col0 + 8 * h + i;
                                float *dst = k < K::MC ? c.work + h * K::LAND_BLOCK + 16 * k + 2 * i
                                                       : c.work + K::NH * K::LAND_BLOCK + K::NYQ_STRIDE * h + 2 * i;
                                const bool ok = col >= 0 && col < P;
                                const float *src = Sf + 2 * ((c.b * F + k) * ldp + col);
                                dst[0] = ok ? src[0] : 0.f;
                                dst[1] = ok ? src[1] : 0.f;
                            }
                    std::vector<typename K::P1Regs> pr((size_t)K::NT);
                    for (int it = 0; it < K::P1_ITERS; ++it) {
                        phase(3, K::NT, [&](int tt) {
                            if (mode == 3) K::template land_load<3>(c, tt, it, pr[tt]);
                            else K::template land_load<2>(c, tt, it, pr[tt]);
                        });
                        phase(3, K::NT, [&](int tt) { K::land_rest(c, tt, it, pr[tt]); });
                    }
                } else
                phase(3, K::NT, [&](int tt) { K::pass1(c, tt); });
                emu_pass2<K, true, K::R1, K::R2, K::L2>(4, c.work, c.tw2);
                phase(5, K::NT, [&](int tt) { K::pass3_load(c, tt, regs[tt]); });
                if (c.first) for (int tt = 0; tt < K::NT; ++tt) K::zero_head(c, tt);
                phase(6, K::NT, [&](int tt) { K::template ola<3>(c, tt, regs[tt]); });
                phase(6, K::NT, [&](int tt) { K::template ola<2>(c, tt, regs[tt]); });
                phase(6, K::NT, [&](int tt) { K::template ola<1>(c, tt, regs[tt]); });
                phase(6, K::NT, [&](int tt) { K::template ola<0>(c, tt, regs[tt]); });
                phase(7, K::NT, [&](int tt) { K::out(c, tt); });
                for (int tt = 0; tt < K::NT; ++tt) K::take_tail(c, tt, tails[tt]);
                for (int tt = 0; tt < K::NT; ++tt) K::keep_tail(c, tt, tails[tt]);
            }
        }
        g_tracing = false;
        return 0;
    }
};

extern "C" {
int wide_emu_supported(int64_t M, int64_t N, int64_t H, int64_t ps) { return wide_geometry_ok(M, N, H, ps) ? 1 : 0; }

int wide_emu_forward(int64_t M, int tf, const double *win, const float *x, int64_t n, int64_t batch, int64_t x_pitch,
                     int64_t p0, int64_t P, int64_t k_offset, int mode, double fac2x, int spectrogram, void *out,
                     int64_t ldp, int64_t F) {
    WCfg c;
    if (!wide_cfg_for(M, tf, &c)) return -1;
    EmuFwd e{x, out, win, n, batch, x_pitch, p0, P, k_offset, ldp, F, mode, spectrogram, fac2x};
    return wide_dispatch(c, e);
}
int wide_emu_inverse(int64_t M, int tf, const double *dual, const void *S, int64_t P, int64_t ldp, int64_t batch,
                     int64_t F, int64_t p_min, int64_t k0, int64_t k1, int64_t q0, int64_t q1, int mode, double fac2x,
                     int chunk_tiles, float *xo, int64_t out_pitch, int use_land) {
    WCfg c;
    if (!wide_cfg_for(M, tf, &c)) return -1;
    EmuInv e{S, xo, dual, P, ldp, batch, out_pitch, p_min, k0, k1, q0, q1, F, mode, chunk_tiles, fac2x, use_land};
    return wide_dispatch(c, e);
}
void wide_emu_conflicts(long long *wavefronts, long long *ideal) {
    for (int i = 0; i < 8; ++i) { wavefronts[i] = g_wavefronts[i]; ideal[i] = g_ideal[i]; }
}
void wide_emu_reset(void) {
    for (int i = 0; i < 8; ++i) g_wavefronts[i] = g_ideal[i] = 0;
}
}
