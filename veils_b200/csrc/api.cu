// api.cu -- the C ABI of include/veils_cuda.h: plan objects, geometry queries, dispatch of the
// compute calls to the kernel families, device-buffer / stream helpers.
#include "../../include/veils_cuda.h"

#include <algorithm>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.hpp"

using namespace veils;

namespace veils {
// device copies of the plan's tables, element type = plan precision
struct DeviceTables {
    void *win = nullptr;          // generic path
    void *dual = nullptr;
    double2 *tw_full = nullptr;
    void *win_half = nullptr;     // pow2 path (exact power-of-two rescalings, see pow2_tile_impl.cuh)
    void *dual_scaled = nullptr;
    void *tw_pow2 = nullptr;
    float *tw_pk = nullptr;
    float *tw_pk_inv = nullptr;
    void *ptab_pk = nullptr;
    void *itab_pk = nullptr;
    void *smooth_fwd = nullptr;   // mixed-radix plan blobs (smooth_impl.cuh)
    void *smooth_inv = nullptr;
    float *wide_fwd = nullptr;    // wide family tables (wide_impl.cuh)
    float *wide_inv = nullptr;
    void *detr_w = nullptr;       // [2][F] complex: fft_func(win * 1), fft_func(win * i)  (detrending)
};
}  // namespace veils

// grow-only device staging buffers + two streams for the host-pointer entry points
struct HostStage {
    void *buf[4] = {nullptr, nullptr, nullptr, nullptr};   // [2 * slot + {0: input chunk, 1: output chunk}]
    size_t cap[4] = {0, 0, 0, 0};
    cudaStream_t st[2] = {nullptr, nullptr};
    std::mutex mu;                                         // host-pointer calls on one plan serialise
    int *nan_flag = nullptr;                               // page-locked, written by nan_dc_check_kernel
    cudaError_t reserve(int i, size_t bytes) {
        if (cap[i] >= bytes) return cudaSuccess;
        if (buf[i]) cudaFree(buf[i]);
        buf[i] = nullptr; cap[i] = 0;
        cudaError_t e = cudaMalloc(&buf[i], bytes);
        if (e == cudaSuccess) cap[i] = bytes;
        return e;
    }
    cudaError_t streams() {
        for (int i = 0; i < 2; ++i)
            if (!st[i]) { cudaError_t e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking); if (e != cudaSuccess) return e; }
        if (!nan_flag) { cudaError_t e = cudaMallocHost((void **)&nan_flag, sizeof(int)); if (e != cudaSuccess) return e; }
        return cudaSuccess;
    }
    void release() {
        for (int i = 0; i < 4; ++i) { if (buf[i]) cudaFree(buf[i]); buf[i] = nullptr; cap[i] = 0; }
        for (int i = 0; i < 2; ++i) { if (st[i]) cudaStreamDestroy(st[i]); st[i] = nullptr; }
        if (nan_flag) cudaFreeHost(nan_flag);
        nan_flag = nullptr;
    }
};

struct veils_plan {
    Plan pl;
    HostStage stage;
};

static thread_local std::string t_err;

static int32_t fail(int32_t code, const std::string &msg) {
    t_err = msg;
    return code;
}
static int32_t cuda_fail(cudaError_t e, const char *what) {
    return fail(VEILS_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                          \
    do {                                                  \
        cudaError_t e_ = (call);                          \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

// exp(-2 pi i m / M): long-double evaluation (error ~1e-19), exact values on the axes
static void twiddle(int64_t m, int64_t M, double *re, double *im) {
    const long double two_pi = 6.283185307179586476925286766559005768L;
    const int64_t mm = ((m % M) + M) % M;
    if ((4 * mm) % M == 0) {                       // multiples of a quarter turn
        static const double c4[4] = {1, 0, -1, 0}, s4[4] = {0, 1, 0, -1};
        const int q = (int)(4 * mm / M);
        *re = c4[q];
        *im = -s4[q];
        return;
    }
    const long double a = two_pi * (long double)mm / (long double)M;
    *re = (double)cosl(a);
    *im = (double)(-sinl(a));
}

template <typename T>
static cudaError_t upload_as(const std::vector<double> &src, void **dst, double scale = 1.0) {
    std::vector<T> tmp(src.size());
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = (T)(src[i] * scale);
    cudaError_t e = cudaMalloc(dst, tmp.size() * sizeof(T));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, tmp.data(), tmp.size() * sizeof(T), cudaMemcpyHostToDevice);
}
static cudaError_t upload_prec(const Plan &pl, const std::vector<double> &src, void **dst,
                               double scale = 1.0) {
    return pl.precision == F32 ? upload_as<float>(src, dst, scale) : upload_as<double>(src, dst, scale);
}

// Create the device tables on first use.  need_dual: also resolve + upload the dual window.
static int32_t ensure_device(Plan &pl, bool need_dual) {
    const int want = need_dual ? 3 : 1;
    if ((pl.dev_ready.load(std::memory_order_acquire) & want) == want) return VEILS_OK;   // no lock, no CUDA call
    if (need_dual) {
        std::string e = plan_dual(pl);
        if (!e.empty()) return fail(VEILS_ERR_NOT_INVERTIBLE, e);
    }
    std::lock_guard<std::mutex> g(pl.mu);
    if (pl.dev && (!need_dual || pl.dev->dual)) return VEILS_OK;      // another thread was faster
    // first use: build the tables on the plan's device without disturbing the caller's current device
    struct DeviceGuard {
        int prev = -1;
        ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    } guard;
    cudaGetDevice(&guard.prev);
    CU(cudaSetDevice(pl.device));
    if (!pl.dev) {
        DeviceTables *d = new (std::nothrow) DeviceTables();
        if (!d) return fail(VEILS_ERR_CUDA, "out of host memory");
        cudaError_t e = upload_prec(pl, pl.win, &d->win);
        if (e == cudaSuccess) e = upload_prec(pl, pl.win, &d->win_half, 0.5);
        if (e == cudaSuccess && (pl.pow2_fwd || pl.pow2_inv)) {
            std::vector<double> tw((size_t)pow2_twiddle_count(pl.M));
            pow2_twiddle_fill(pl.M, tw.data());
            e = upload_prec(pl, tw, &d->tw_pow2);
        }
        if (e == cudaSuccess && (pl.pk_fwd || pl.pk_inv)) {
            std::vector<double> tw((size_t)pk_twiddle_count(pl.M));
            pk_twiddle_fill(pl.M, false, tw.data());
            e = upload_as<float>(tw, (void **)&d->tw_pk);
            if (e == cudaSuccess) {
                pk_twiddle_fill(pl.M, true, tw.data());
                e = upload_as<float>(tw, (void **)&d->tw_pk_inv);
            }
            if (e == cudaSuccess && pl.pk_fwd) {
                std::vector<unsigned char> pt((size_t)(pl.M / 2) * 16);
                pk_ptab_fill_host(pl.M, pl.N, pl.H, pl.ps, pl.win.data(), pt.data());
                e = cudaMalloc(&d->ptab_pk, pt.size());
                if (e == cudaSuccess) e = cudaMemcpy(d->ptab_pk, pt.data(), pt.size(), cudaMemcpyHostToDevice);
            }
        }
        if (e == cudaSuccess && pl.smooth_fwd) {
            std::vector<unsigned char> blob((size_t)smooth_table_bytes(pl.M, pl.N, pl.H, pl.ps, pl.precision == F32 ? 0 : 1, false));
            smooth_table_fill(pl.M, pl.N, pl.H, pl.ps, pl.precision == F32 ? 0 : 1, false, pl.win.data(), blob.data());
            e = cudaMalloc(&d->smooth_fwd, blob.size());
            if (e == cudaSuccess) e = cudaMemcpy(d->smooth_fwd, blob.data(), blob.size(), cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess && pl.wide) {
            std::vector<double> tw((size_t)wide_table_floats(pl.M));
            wide_table_fill_host(pl.M, pl.wide_tf, false, pl.win.data(), 0.5, tw.data());
            e = upload_as<float>(tw, (void **)&d->wide_fwd);
        }
        if (e == cudaSuccess && !((pl.pow2_fwd || pl.pk_fwd || pl.smooth_fwd) && (pl.pow2_inv || pl.pk_inv || pl.smooth_inv))) {
            std::vector<double> tw((size_t)pl.M * 2);
            for (int64_t m = 0; m < pl.M; ++m) twiddle(m, pl.M, &tw[2 * m], &tw[2 * m + 1]);
            e = cudaMalloc((void **)&d->tw_full, tw.size() * sizeof(double));
            if (e == cudaSuccess)
                e = cudaMemcpy(d->tw_full, tw.data(), tw.size() * sizeof(double), cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) {
            cudaFree(d->win); cudaFree(d->win_half); cudaFree(d->tw_pow2); cudaFree(d->tw_full); cudaFree(d->tw_pk); cudaFree(d->tw_pk_inv); cudaFree(d->ptab_pk); cudaFree(d->wide_fwd); cudaFree(d->smooth_fwd);
            delete d;
            return cuda_fail(e, "upload plan tables");
        }
        pl.dev = d;
        pl.dev_ready.fetch_or(1, std::memory_order_release);
    }
    if (need_dual && !pl.dev->dual) {
        // pow2 inverse folds 1/M (and the 1/2 of the Hermitian part for two-sided layouts) here.
        // Everything is uploaded into locals and published only when all of it succeeded.
        const double sc = 1.0 / ((double)pl.M * (pl.mode >= ONESIDED ? 1.0 : 2.0));
        void *dual = nullptr, *dual_scaled = nullptr, *itab = nullptr;
        float *wide_inv = nullptr;
        cudaError_t e = upload_prec(pl, pl.dual, &dual_scaled, sc);
        if (e == cudaSuccess) e = upload_prec(pl, pl.dual, &dual);
        if (e == cudaSuccess && pl.pk_inv) {
            std::vector<double> ds(pl.dual);
            for (double &v : ds) v *= sc;
            std::vector<unsigned char> it((size_t)(pl.M / 2) * 16);
            pk_itab_fill_host(pl.M, pl.N, pl.ps, ds.data(), it.data());
            e = cudaMalloc(&itab, it.size());
            if (e == cudaSuccess) e = cudaMemcpy(itab, it.data(), it.size(), cudaMemcpyHostToDevice);
        }
        if (e == cudaSuccess && pl.wide_inv) {
            std::vector<double> tw((size_t)wide_table_floats(pl.M));
            wide_table_fill_host(pl.M, pl.wide_tf, true, pl.dual.data(), sc, tw.data());
            e = upload_as<float>(tw, (void **)&wide_inv);
        }
        void *smooth_inv = nullptr;
        if (e == cudaSuccess && pl.smooth_inv) {
            std::vector<double> ds(pl.dual);
            for (double &v : ds) v *= sc;
            std::vector<unsigned char> blob((size_t)smooth_table_bytes(pl.M, pl.N, pl.H, pl.ps, pl.precision == F32 ? 0 : 1, true));
            smooth_table_fill(pl.M, pl.N, pl.H, pl.ps, pl.precision == F32 ? 0 : 1, true, ds.data(), blob.data());
            e = cudaMalloc(&smooth_inv, blob.size());
            if (e == cudaSuccess) e = cudaMemcpy(smooth_inv, blob.data(), blob.size(), cudaMemcpyHostToDevice);
        }
        if (e != cudaSuccess) {
            cudaFree(dual); cudaFree(dual_scaled); cudaFree(itab); cudaFree(wide_inv); cudaFree(smooth_inv);
            return cuda_fail(e, "upload dual window");
        }
        pl.dev->smooth_inv = smooth_inv;
        pl.dev->dual_scaled = dual_scaled;
        pl.dev->itab_pk = itab;
        pl.dev->wide_inv = wide_inv;
        pl.dev->dual = dual;
        pl.dev_ready.fetch_or(2, std::memory_order_release);
    }
    return VEILS_OK;
}

// kernels are launched on the plan's device; the caller's current device is restored afterwards
struct PlanDeviceScope {
    int prev = -1;
    explicit PlanDeviceScope(int dev) {
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != dev) { prev = cur; cudaSetDevice(dev); }
    }
    ~PlanDeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
static Tables<T> tables_of(const Plan &pl) {
    Tables<T> tb;
    tb.win = (const T *)pl.dev->win;
    tb.dual = (const T *)pl.dev->dual;
    tb.tw_full = pl.dev->tw_full;
    tb.win_half = (const T *)pl.dev->win_half;
    tb.dual_scaled = (const T *)pl.dev->dual_scaled;
    tb.tw_pow2 = (const T *)pl.dev->tw_pow2;
    tb.tw_pk = pl.dev->tw_pk;
    tb.tw_pk_inv = pl.dev->tw_pk_inv;
    tb.ptab_pk = pl.dev->ptab_pk;
    tb.itab_pk = pl.dev->itab_pk;
    tb.smooth_fwd = pl.dev->smooth_fwd;
    tb.smooth_inv = pl.dev->smooth_inv;
    tb.wide_fwd = pl.dev->wide_fwd;
    tb.wide_inv = pl.dev->wide_inv;
    return tb;
}

extern "C" {

int32_t veils_abi_version(void) { return VEILS_ABI_VERSION; }
const char *veils_last_error(void) { return t_err.c_str(); }

int32_t veils_plan_create(const veils_plan_desc *d, veils_plan **out) {
    if (!d || !out) return fail(VEILS_ERR_INVALID, "null argument");
    *out = nullptr;
    veils_plan *p = new (std::nothrow) veils_plan();
    if (!p) return fail(VEILS_ERR_INVALID, "out of host memory");
    std::string e = plan_init(p->pl, d->win, d->m_num, d->hop, d->fs, d->fft_mode, d->mfft,
                              d->dual_win, d->scale_to, d->phase_shift, d->precision, d->device);
    if (!e.empty()) {
        delete p;
        return fail(VEILS_ERR_INVALID, e);
    }
    {
        Plan &pl = p->pl;
        const bool force_generic = d->precision >= 0 && getenv("VEILS_FORCE_GENERIC") != nullptr;
        pl.pow2_fwd = !force_generic && pow2_supported(pl.M, pl.N, pl.H, pl.ps, pl.precision, false);
        pl.pow2_inv = !force_generic && pow2_supported(pl.M, pl.N, pl.H, pl.ps, pl.precision, true);
        const bool no_pk = getenv("VEILS_NO_PACKED") != nullptr;
        // measured on B200 (profiles/r1_config_sweep.txt): the packed forward wins up to mfft = 1024;
        // beyond, its 255-register 3-pass build loses to the scalar family (and both to the wide family)
        pl.pk_fwd = !force_generic && !no_pk && pl.precision == F32 && (pl.M <= 1024 || !pl.pow2_fwd || getenv("VEILS_PK_FWD_ALL") != nullptr) &&
                    pk_supported(pl.M, pl.N, pl.H, pl.ps, false);
        pl.pk_inv = !force_generic && !no_pk && pl.precision == F32 && pk_supported(pl.M, pl.N, pl.H, pl.ps, true);
        // mfft 1024 / 2048 / 4096 with the default geometry: the wide family (conflict-free swizzled tiles)
        pl.wide = !force_generic && pl.precision == F32 && getenv("VEILS_NO_WIDE") == nullptr &&
                  wide_supported(pl.M, pl.N, pl.H, pl.ps);
        pl.wide_tf = pl.wide ? wide_tf(pl.M) : 0;
        // lengths the power-of-two families do not take: mixed radix (prime factors <= 13, <= 4 passes) before
        // the O(M^2) kernels
        const bool no_smooth = getenv("VEILS_NO_SMOOTH") != nullptr;
        // VEILS_PREFER_SMOOTH: 1 = every precision, 2 = f64 only (measurement switch: the mixed-radix family also
        // takes powers of two)
        const char *pref = getenv("VEILS_PREFER_SMOOTH");
        if (!force_generic && !no_smooth && pref && (pref[0] == '1' || (pref[0] == '2' && pl.precision != F32))) {
            if (smooth_supported(pl.M, pl.N, pl.H, pl.precision, false)) { pl.pow2_fwd = pl.pk_fwd = pl.wide = false; pl.wide_tf = 0; }
            if (smooth_supported(pl.M, pl.N, pl.H, pl.precision, true)) { pl.pow2_inv = pl.pk_inv = false; }
        }
        pl.smooth_fwd = !force_generic && !no_smooth && !pl.pow2_fwd && !pl.pk_fwd && !pl.wide &&
                        smooth_supported(pl.M, pl.N, pl.H, pl.precision, false);
        pl.smooth_inv = !force_generic && !no_smooth && !pl.pow2_inv && !pl.pk_inv && !pl.wide_inv &&
                        smooth_supported(pl.M, pl.N, pl.H, pl.precision, true);
        // two-sided inverse: two rows per bin, no tensor-map landing yet -> the packed family is faster there
        // (measured c4: 49 % vs 27 % of the roofline); VEILS_WIDE_INV_ALL=1 forces the wide kernels
        pl.wide_inv = pl.wide && !pl.smooth_inv && (pl.mode >= ONESIDED || !pl.pk_inv || getenv("VEILS_WIDE_INV_ALL") != nullptr);
    }
    *out = p;
    return VEILS_OK;
}

void veils_plan_destroy(veils_plan *p) {
    if (!p) return;
    if (p->pl.dev) {
        cudaSetDevice(p->pl.device);
        cudaFree(p->pl.dev->win);
        cudaFree(p->pl.dev->dual);
        cudaFree(p->pl.dev->tw_full);
        cudaFree(p->pl.dev->win_half);
        cudaFree(p->pl.dev->dual_scaled);
        cudaFree(p->pl.dev->tw_pow2);
        cudaFree(p->pl.dev->tw_pk);
        cudaFree(p->pl.dev->tw_pk_inv);
        cudaFree(p->pl.dev->detr_w);
        cudaFree(p->pl.dev->ptab_pk);
        cudaFree(p->pl.dev->itab_pk);
        cudaFree(p->pl.dev->wide_fwd);
        cudaFree(p->pl.dev->wide_inv);
        cudaFree(p->pl.dev->smooth_fwd);
        cudaFree(p->pl.dev->smooth_inv);
        delete p->pl.dev;
    }
    cudaSetDevice(p->pl.device);
    p->stage.release();
    delete p;
}

int64_t veils_m_num(const veils_plan *p) { return p->pl.N; }
int64_t veils_m_num_mid(const veils_plan *p) { return p->pl.mid; }
int64_t veils_hop(const veils_plan *p) { return p->pl.H; }
int64_t veils_mfft(const veils_plan *p) { return p->pl.M; }
int64_t veils_f_pts(const veils_plan *p) { return p->pl.F; }
int64_t veils_phase_shift(const veils_plan *p) { return p->pl.phase_shift; }
double veils_fs(const veils_plan *p) { return p->pl.fs; }
double veils_delta_t(const veils_plan *p) { return (double)p->pl.H * (1.0 / p->pl.fs); }
double veils_delta_f(const veils_plan *p) { return 1.0 / ((double)p->pl.M * (1.0 / p->pl.fs)); }
int32_t veils_fft_mode(const veils_plan *p) { return p->pl.mode; }
int32_t veils_onesided_fft(const veils_plan *p) { return p->pl.mode >= ONESIDED; }
int32_t veils_scaling(const veils_plan *p) { return p->pl.scaling; }
int32_t veils_precision(const veils_plan *p) { return p->pl.precision; }
int32_t veils_device(const veils_plan *p) { return p->pl.device; }

int32_t veils_win(const veils_plan *p, double *out) {
    if (!p || !out) return fail(VEILS_ERR_INVALID, "null argument");
    std::memcpy(out, p->pl.win.data(), p->pl.win.size() * sizeof(double));
    return VEILS_OK;
}

int32_t veils_dual_win(veils_plan *p, double *out) {
    if (!p || !out) return fail(VEILS_ERR_INVALID, "null argument");
    std::string e = plan_dual(p->pl);
    if (!e.empty()) return fail(VEILS_ERR_NOT_INVERTIBLE, e);
    std::memcpy(out, p->pl.dual.data(), p->pl.dual.size() * sizeof(double));
    return VEILS_OK;
}

int32_t veils_invertible(veils_plan *p) { return p && plan_dual(p->pl).empty() ? 1 : 0; }

int64_t veils_p_min(const veils_plan *p) { return p->pl.p_min; }
int64_t veils_k_min(const veils_plan *p) { return p->pl.k_min; }

int32_t veils_p_max(const veils_plan *p, int64_t n, int64_t *out) {
    std::string e = plan_post_padding(p->pl, n, nullptr, out);
    return e.empty() ? VEILS_OK : fail(VEILS_ERR_INVALID, e);
}
int32_t veils_k_max(const veils_plan *p, int64_t n, int64_t *out) {
    std::string e = plan_post_padding(p->pl, n, out, nullptr);
    return e.empty() ? VEILS_OK : fail(VEILS_ERR_INVALID, e);
}
int32_t veils_p_num(const veils_plan *p, int64_t n, int64_t *out) {
    int64_t pm = 0;
    std::string e = plan_post_padding(p->pl, n, nullptr, &pm);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    *out = pm - p->pl.p_min;
    return VEILS_OK;
}
int32_t veils_p_range(const veils_plan *p, int64_t n, const int64_t *p0, const int64_t *p1,
                      int64_t *o0, int64_t *o1) {
    std::string e = plan_p_range(p->pl, n, p0, p1, o0, o1);
    return e.empty() ? VEILS_OK : fail(VEILS_ERR_INVALID, e);
}
int32_t veils_istft_range(veils_plan *p, int64_t P, const int64_t *k0, const int64_t *k1,
                          int64_t *o0, int64_t *o1) {
    IstftRange r;
    std::string e = plan_istft_range(p->pl, P, k0, k1, &r);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    *o0 = r.k0;
    *o1 = r.k1;
    return VEILS_OK;
}

int32_t veils_t(const veils_plan *p, int64_t n, const int64_t *p0, const int64_t *p1,
                int64_t k_offset, double *out) {
    int64_t a, b;
    std::string e = plan_p_range(p->pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    const double T = 1.0 / p->pl.fs;
    for (int64_t q = a; q < b; ++q) out[q - a] = (double)(q * p->pl.H + k_offset) * T;
    return VEILS_OK;
}

int32_t veils_f(const veils_plan *p, double *out) {
    const Plan &pl = p->pl;
    const double T = 1.0 / pl.fs;
    const int64_t M = pl.M;
    if (pl.mode >= ONESIDED) {
        for (int64_t i = 0; i < pl.F; ++i) out[i] = (double)i / ((double)M * T);
        return VEILS_OK;
    }
    // fftfreq order (src/lib.rs:936-945); numpy puts i >= ceil(M/2) on the negative side
    std::vector<double> fr((size_t)M);
    for (int64_t i = 0; i < M; ++i) {
        if (pl.mode == CENTERED) {
            const int64_t s = (i < (M + 1) / 2) ? i : i - M;          // numpy.fft.fftfreq
            fr[i] = (double)s / ((double)M * T);
        } else {
            double f = (double)i / ((double)M * T);
            fr[i] = (i > M / 2) ? f - 1.0 / T : f;
        }
    }
    if (pl.mode == CENTERED) {                                         // fftshift (scipy)
        const int64_t sh = (M + 1) / 2;
        for (int64_t i = 0; i < M; ++i) out[i] = fr[(i + sh) % M];
    } else {
        std::memcpy(out, fr.data(), (size_t)M * sizeof(double));
    }
    return VEILS_OK;
}

// ---- scipy.signal.ShortTimeFFT axis helpers (scipy _short_time_fft.py: lower_border_end, upper_border_begin,
// extent, nearest_k_p); host integer logic on the plan's window
static int64_t floordiv64(int64_t a, int64_t b) {
    int64_t q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
int32_t veils_lower_border_end(const veils_plan *p, int64_t *k, int64_t *pp) {
    if (!p || !k || !pp) return fail(VEILS_ERR_INVALID, "null argument");
    const Plan &pl = p->pl;
    int64_t m0 = 0;
    while (m0 < pl.N && pl.win[(size_t)m0] == 0.0) ++m0;              // first non-zero window sample
    int64_t q = 0;
    for (int64_t kk = -pl.mid + m0; kk < pl.H + 1; kk += pl.H, ++q)
        if (kk + pl.H >= 0) { *k = kk + pl.N; *pp = q + 1; return VEILS_OK; }
    *k = 0;
    *pp = pl.p_min > 0 ? pl.p_min : 0;
    return VEILS_OK;
}
int32_t veils_upper_border_begin(const veils_plan *p, int64_t n, int64_t *k, int64_t *pp) {
    if (!p || !k || !pp) return fail(VEILS_ERR_INVALID, "null argument");
    const Plan &pl = p->pl;
    const int64_t m2p = pl.N - pl.mid;
    if (n < m2p) return fail(VEILS_ERR_INVALID, "Parameter n must be >= ceil(m_num/2) = " + std::to_string((long long)m2p) + "!");
    const int64_t q2 = n / pl.H + 1;
    int64_t q1 = floordiv64(n - pl.N, pl.H) - 1;
    if (q1 < -1) q1 = -1;
    for (int64_t q = q2; q > q1; --q) {
        const int64_t kk = q * pl.H + (pl.N - pl.mid);
        bool ok = kk <= n;
        if (!ok) {                                                     // all(w2[n - kk:] == 0), numpy negative slicing
            int64_t from = n - kk;                                     // < 0: counts from the end
            if (from < -pl.N) from = -pl.N;
            ok = true;
            for (int64_t i = pl.N + from; i < pl.N; ++i)
                if (pl.win[(size_t)i] != 0.0) { ok = false; break; }
        }
        if (ok) { *k = (q + 1) * pl.H - pl.mid; *pp = q + 1; return VEILS_OK; }
    }
    return fail(VEILS_ERR_INVALID, "upper_border_begin: no slice found");
}
int32_t veils_extent(const veils_plan *p, int64_t n, int32_t ft_order, int32_t center_bins, double *out4) {
    if (!p || !out4) return fail(VEILS_ERR_INVALID, "null argument");
    const Plan &pl = p->pl;
    int64_t q0, q1;
    if (pl.mode >= ONESIDED) { q0 = 0; q1 = pl.F; }
    else if (pl.mode == CENTERED) { q0 = -(pl.M / 2); q1 = pl.M % 2 == 0 ? pl.M / 2 : pl.M / 2 + 1; }
    else return fail(VEILS_ERR_INVALID, "Attribute fft_mode=twosided must be in ['centered', 'onesided', 'onesided2X']");
    int64_t p1;
    int32_t rc = veils_p_max(p, n, &p1);
    if (rc != VEILS_OK) return rc;
    const double o = center_bins ? 0.5 : 0.0, dt = (double)pl.H / pl.fs, df = pl.fs / (double)pl.M;
    const double t0 = dt * ((double)pl.p_min - o), t1 = dt * ((double)p1 - o);
    const double f0 = df * ((double)q0 - o), f1 = df * ((double)q1 - o);
    if (ft_order) { out4[0] = f0; out4[1] = f1; out4[2] = t0; out4[3] = t1; }
    else { out4[0] = t0; out4[1] = t1; out4[2] = f0; out4[3] = f1; }
    return VEILS_OK;
}
int64_t veils_nearest_k_p(const veils_plan *p, int64_t k, int32_t left) {
    const int64_t H = p->pl.H, q = floordiv64(k, H);
    if (q * H == k) return k;
    return left ? q * H : (q + 1) * H;
}

// StandaloneSTFT::debug_ifft_func (src/lib.rs:952-954 -> ifft_func :414-505) of ONE slice, on the host:
// mode prologue, direct O(M^2) inverse DFT in f64 (a debug helper, not a compute path), 1/M, roll right by
// p_s over mfft, first m_num samples.  X: f_pts complex {re, im}; out: m_num complex.  One-sided layouts
// follow irfft semantics (imaginary parts of DC / Nyquist ignored), so the result is real.
int32_t veils_debug_ifft_func(const veils_plan *p, const double *X, double *out) {
    if (!p || !X || !out) return fail(VEILS_ERR_INVALID, "null argument");
    const Plan &pl = p->pl;
    const int64_t M = pl.M, N = pl.N, F = pl.F;
    std::vector<double> fr((size_t)M), fi((size_t)M);
    if (pl.mode >= ONESIDED) {
        const double c = pl.mode == ONESIDED2X ? 1.0 / pl.fac2x : 1.0;
        for (int64_t k = 0; k < F; ++k) {
            const bool paired = k >= 1 && (M % 2 == 0 ? k <= F - 2 : k <= F - 1);
            const double sc = paired ? c : 1.0;
            fr[k] = X[2 * k] * sc;
            fi[k] = (k == 0 || (M % 2 == 0 && k == M / 2)) ? 0.0 : X[2 * k + 1] * sc;
        }
        for (int64_t k = F; k < M; ++k) { fr[k] = fr[M - k]; fi[k] = -fi[M - k]; }
    } else {
        const int64_t sh = pl.mode == CENTERED ? M / 2 : 0;          // ifftshift = rotate_left(floor(M/2)) :516-521
        for (int64_t k = 0; k < M; ++k) { fr[k] = X[2 * ((k + sh) % M)]; fi[k] = X[2 * ((k + sh) % M) + 1]; }
    }
    std::vector<double> yr((size_t)M), yi((size_t)M);
    for (int64_t nn = 0; nn < M; ++nn) {
        long double ar = 0, ai = 0;
        for (int64_t k = 0; k < M; ++k) {
            double wr, wi;
            twiddle(-(k * nn), M, &wr, &wi);                          // exp(+2 pi i k n / M)
            ar += (long double)fr[k] * wr - (long double)fi[k] * wi;
            ai += (long double)fr[k] * wi + (long double)fi[k] * wr;
        }
        yr[nn] = (double)(ar / M);
        yi[nn] = (double)(ai / M);
    }
    for (int64_t j = 0; j < N; ++j) {                                 // roll right by ps: out[j] = y[(j - ps) mod M]
        const int64_t i = ((j - pl.ps) % M + M) % M;
        out[2 * j] = yr[i];
        out[2 * j + 1] = yi[i];
    }
    return VEILS_OK;
}

// ---- compute -----------------------------------------------------------------------------
static int32_t forward_impl(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                            const int64_t *p0, const int64_t *p1, int64_t k_offset, void *out,
                            int64_t ldp, void *stream, int spectrogram) {
    if (!p || !x || !out) return fail(VEILS_ERR_INVALID, "null argument");
    Plan &pl = p->pl;
    if (batch < 0) return fail(VEILS_ERR_INVALID, "batch must be >= 0");
    const int64_t m2p = pl.N - pl.mid;
    if (n < m2p)                                              // src/lib.rs:700-707
        return fail(VEILS_ERR_INVALID, "Input length " + std::to_string((long long)n) +
                                           " must be >= ceil(m_num/2) = " + std::to_string((long long)m2p));
    int64_t a, b;
    std::string e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    if (x_pitch < n) return fail(VEILS_ERR_INVALID, "x_pitch must be >= n");
    if (ldp < b - a) return fail(VEILS_ERR_INVALID, "ldp must be >= p1 - p0");
    if (batch == 0) return VEILS_OK;
    int32_t rc = ensure_device(pl, false);
    if (rc != VEILS_OK) return rc;
    PlanDeviceScope scope(pl.device);
    FwdParams prm;
    prm.n = n; prm.x_pitch = x_pitch; prm.batch = batch;
    prm.N = pl.N; prm.H = pl.H; prm.M = pl.M; prm.F = pl.F; prm.mid = pl.mid; prm.ps = pl.ps;
    prm.p0 = a; prm.P = b - a; prm.k_offset = k_offset; prm.ldp = ldp;
    prm.mode = pl.mode; prm.fac2x = pl.fac2x; prm.spectrogram = spectrogram;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce;
    if (pl.precision == F32) {
        Tables<float> tb = tables_of<float>(pl);
        ce = pl.wide ? wide_forward((const float *)x, out, prm, tb.wide_fwd, pl.wide_tf, st)
             : pl.pk_fwd ? pk_forward((const float *)x, out, prm, tb, st)
             : pl.pow2_fwd ? pow2_forward<float>((const float *)x, out, prm, tb, st)
             : pl.smooth_fwd ? smooth_forward<float>((const float *)x, out, prm, tb.smooth_fwd, st)
                           : dft_generic_forward<float>((const float *)x, out, prm, tb, st);
    } else {
        Tables<double> tb = tables_of<double>(pl);
        ce = pl.pow2_fwd ? pow2_forward<double>((const double *)x, out, prm, tb, st)
             : pl.smooth_fwd ? smooth_forward<double>((const double *)x, out, prm, tb.smooth_fwd, st)
                          : dft_generic_forward<double>((const double *)x, out, prm, tb, st);
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "stft launch");
    return VEILS_OK;
}

int32_t veils_stft_dev(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                       const int64_t *p0, const int64_t *p1, int64_t k_offset, void *S, int64_t ldp,
                       void *stream) {
    return forward_impl(p, x, n, batch, x_pitch, p0, p1, k_offset, S, ldp, stream, 0);
}

int32_t veils_spectrogram_dev(veils_plan *p, const void *x, int64_t n, int64_t batch,
                              int64_t x_pitch, const int64_t *p0, const int64_t *p1,
                              int64_t k_offset, void *Sxx, int64_t ldp, void *stream) {
    return forward_impl(p, x, n, batch, x_pitch, p0, p1, k_offset, Sxx, ldp, stream, 1);
}

static int32_t nan_check_launch(const Plan &pl, const void *x, int64_t n, int64_t batch, int64_t x_pitch, int64_t a,
                                int64_t P, int64_t k_offset, const void *S, int64_t ldp, int spectrogram, int *flag,
                                cudaStream_t st);
// device-pointer flavour with the reference's NaN screen (src/lib.rs:696-698): *nan_flag (device memory or
// page-locked host memory, zeroed by the caller) becomes 1 when a NaN sample went into a one-sided transform
int32_t veils_stft_dev_checked(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                               const int64_t *p0, const int64_t *p1, int64_t k_offset, void *S, int64_t ldp,
                               void *stream, int32_t spectrogram, int32_t *nan_flag) {
    int32_t rc = forward_impl(p, x, n, batch, x_pitch, p0, p1, k_offset, S, ldp, stream, spectrogram ? 1 : 0);
    if (rc != VEILS_OK || !nan_flag) return rc;
    int64_t a, b;
    std::string e = plan_p_range(p->pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    return nan_check_launch(p->pl, x, n, batch, x_pitch, a, b - a, k_offset, S, ldp, spectrogram ? 1 : 0, nan_flag,
                            (cudaStream_t)stream);
}

// Host-side cost of one veils_stft_dev call: `reps` back-to-back calls, wall clock / reps, measured inside the
// library (no binding overhead).  The stream is not synchronised: this is what the caller's thread pays per launch.
int32_t veils_debug_launch_cost_us(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch, void *S,
                                   int64_t ldp, void *stream, int32_t reps, double *us_per_call) {
    if (!us_per_call || reps <= 0) return fail(VEILS_ERR_INVALID, "null argument");
    int32_t rc = forward_impl(p, x, n, batch, x_pitch, nullptr, nullptr, 0, S, ldp, stream, 0);    // warm: tables, caches
    if (rc != VEILS_OK) return rc;
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int32_t i = 0; i < reps && rc == VEILS_OK; ++i)
        rc = forward_impl(p, x, n, batch, x_pitch, nullptr, nullptr, 0, S, ldp, stream, 0);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *us_per_call = ((double)(t1.tv_sec - t0.tv_sec) * 1e6 + (double)(t1.tv_nsec - t0.tv_nsec) * 1e-3) / reps;
    return rc;
}

static int padding_mode(const char *padding) {
    if (!padding || !strcmp(padding, "zeros")) return 0;
    if (!strcmp(padding, "edge")) return 1;
    if (!strcmp(padding, "even")) return 2;
    if (!strcmp(padding, "odd")) return 3;
    return -1;
}

int32_t veils_pad_extent(const veils_plan *p, int64_t n, const int64_t *p0, const int64_t *p1, int64_t k_offset,
                         int64_t *left, int64_t *n_ext) {
    if (!p || !left || !n_ext) return fail(VEILS_ERR_INVALID, "null argument");
    const Plan &pl = p->pl;
    int64_t a, b;
    std::string e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    const int64_t lo = a * pl.H + k_offset - pl.mid, hi = (b - 1) * pl.H + k_offset - pl.mid + pl.N;
    const int64_t l = lo < 0 ? -lo : 0, r = hi > n ? hi - n : 0;
    *left = l;
    *n_ext = l + n + r;
    return VEILS_OK;
}

int32_t veils_stft_dev_padded(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                              const int64_t *p0, const int64_t *p1, int64_t k_offset, const char *padding,
                              void *x_ext, int64_t ext_pitch, void *S, int64_t ldp, void *stream, int32_t spectrogram) {
    if (!p || !x || !S) return fail(VEILS_ERR_INVALID, "null argument");
    const int mode = padding_mode(padding);
    if (mode < 0) return fail(VEILS_ERR_INVALID, std::string("Parameter padding='") + padding + "' not in ['zeros', 'edge', 'even', 'odd']!");
    if (mode == 0) return forward_impl(p, x, n, batch, x_pitch, p0, p1, k_offset, S, ldp, stream, spectrogram ? 1 : 0);
    Plan &pl = p->pl;
    const int64_t m2p = pl.N - pl.mid;
    if (n < m2p)
        return fail(VEILS_ERR_INVALID, "Input length " + std::to_string((long long)n) +
                                           " must be >= ceil(m_num/2) = " + std::to_string((long long)m2p));
    int64_t a, b, left, n_ext;
    std::string e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    int32_t rc = veils_pad_extent(p, n, p0, p1, k_offset, &left, &n_ext);
    if (rc != VEILS_OK) return rc;
    if (!x_ext) return fail(VEILS_ERR_INVALID, "x_ext workspace required for padding != 'zeros' (size: veils_pad_extent)");
    if (ext_pitch < n_ext) return fail(VEILS_ERR_INVALID, "ext_pitch must be >= n_ext");
    if (x_pitch < n) return fail(VEILS_ERR_INVALID, "x_pitch must be >= n");
    if (batch <= 0) return batch == 0 ? VEILS_OK : fail(VEILS_ERR_INVALID, "batch must be >= 0");
    rc = ensure_device(pl, false);
    if (rc != VEILS_OK) return rc;
    PlanDeviceScope scope(pl.device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = pl.precision == F32
                         ? pad_signal<float>((const float *)x, (float *)x_ext, n, batch, x_pitch, left, n_ext, ext_pitch, mode, st)
                         : pad_signal<double>((const double *)x, (double *)x_ext, n, batch, x_pitch, left, n_ext, ext_pitch, mode, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "pad launch");
    return forward_impl(p, x_ext, n_ext, batch, ext_pitch, &a, &b, k_offset + left, S, ldp, stream, spectrogram ? 1 : 0);
}

// W0 = fft_func(win * 1), W1 = fft_func(win * i): one slice that covers [0, m_num) of a ones / ramp signal,
// through the plan's own forward path (same roll, mode epilogue and scaling as S)
static int32_t ensure_detrend_tables(veils_plan *p, void *stream) {
    Plan &pl = p->pl;
    int32_t rc = ensure_device(pl, false);
    if (rc != VEILS_OK) return rc;
    if (pl.dev->detr_w) return VEILS_OK;
    const size_t esz = pl.precision == F32 ? 4 : 8;
    std::vector<double> sig((size_t)(2 * pl.N));
    for (int64_t i = 0; i < pl.N; ++i) { sig[(size_t)i] = 1.0; sig[(size_t)(pl.N + i)] = (double)i; }
    void *xs = nullptr, *w = nullptr;
    cudaError_t e = upload_prec(pl, sig, &xs);
    if (e == cudaSuccess) e = cudaMalloc(&w, (size_t)(2 * pl.F) * 2 * esz);
    if (e != cudaSuccess) { cudaFree(xs); cudaFree(w); return cuda_fail(e, "detrend tables"); }
    const int64_t zero = 0, one = 1;
    rc = forward_impl(p, xs, pl.N, 2, pl.N, &zero, &one, pl.mid, w, 1, stream, 0);
    if (rc == VEILS_OK) {
        e = cudaStreamSynchronize((cudaStream_t)stream);
        if (e != cudaSuccess) rc = cuda_fail(e, "detrend tables");
    }
    cudaFree(xs);
    if (rc != VEILS_OK) { cudaFree(w); return rc; }
    std::lock_guard<std::mutex> g(pl.mu);
    if (pl.dev->detr_w) cudaFree(w);
    else pl.dev->detr_w = w;
    return VEILS_OK;
}

int32_t veils_stft_detrend_dev(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                               const int64_t *p0, const int64_t *p1, int64_t k_offset, const char *padding,
                               const char *detr, void *x_ext, int64_t ext_pitch, void *coef_ws, void *S,
                               int64_t ldp, void *stream) {
    if (!p || !x || !S) return fail(VEILS_ERR_INVALID, "null argument");
    if (!detr) return veils_stft_dev_padded(p, x, n, batch, x_pitch, p0, p1, k_offset, padding, x_ext, ext_pitch, S, ldp, stream, 0);
    const int linear = !strcmp(detr, "linear") ? 1 : (!strcmp(detr, "constant") ? 0 : -1);
    if (linear < 0) return fail(VEILS_ERR_INVALID, std::string("Unknown detrend type: ") + detr);
    if (!coef_ws) return fail(VEILS_ERR_INVALID, "coef_ws workspace required (16 bytes per slice and signal)");
    const int mode = padding_mode(padding);
    if (mode < 0) return fail(VEILS_ERR_INVALID, std::string("Parameter padding='") + padding + "' not in ['zeros', 'edge', 'even', 'odd']!");
    Plan &pl = p->pl;
    int32_t rc = veils_stft_dev_padded(p, x, n, batch, x_pitch, p0, p1, k_offset, padding, x_ext, ext_pitch, S, ldp, stream, 0);
    if (rc != VEILS_OK || batch == 0) return rc;
    int64_t a, b, left = 0, n_ext = n;
    std::string e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    if (mode != 0) {
        rc = veils_pad_extent(p, n, p0, p1, k_offset, &left, &n_ext);
        if (rc != VEILS_OK) return rc;
    }
    rc = ensure_detrend_tables(p, stream);
    if (rc != VEILS_OK) return rc;
    const void *xs = mode != 0 ? x_ext : x;
    const int64_t pitch = mode != 0 ? ext_pitch : x_pitch;
    const int64_t start0 = k_offset + left - pl.mid;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = pl.precision == F32
        ? detrend_correct<float>((const float *)xs, n_ext, batch, pitch, a, b - a, pl.H, pl.N, start0, linear,
                                 (const float *)pl.dev->detr_w, (double2 *)coef_ws, (float *)S, pl.F, ldp, st)
        : detrend_correct<double>((const double *)xs, n_ext, batch, pitch, a, b - a, pl.H, pl.N, start0, linear,
                                  (const double *)pl.dev->detr_w, (double2 *)coef_ws, (double *)S, pl.F, ldp, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "detrend launch");
    return VEILS_OK;
}

int32_t veils_cross_multiply_dev(veils_plan *p, void *Sx, const void *Sy, int64_t P, int64_t batch, int64_t ldx,
                                 int64_t ldy, void *stream) {
    if (!p || !Sx || !Sy) return fail(VEILS_ERR_INVALID, "null argument");
    if (P < 0 || batch < 0 || ldx < P || ldy < P) return fail(VEILS_ERR_INVALID, "ldx, ldy must be >= P >= 0");
    Plan &pl = p->pl;
    int32_t rc = ensure_device(pl, false);
    if (rc != VEILS_OK) return rc;
    PlanDeviceScope scope(pl.device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = pl.precision == F32 ? cross_multiply<float>((float *)Sx, (const float *)Sy, batch * pl.F, P, ldx, ldy, st)
                                         : cross_multiply<double>((double *)Sx, (const double *)Sy, batch * pl.F, P, ldx, ldy, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "cross-spectrogram launch");
    return VEILS_OK;
}

int32_t veils_complex_op_dev(veils_plan *p, int32_t op, void *A, const void *B, int64_t P, int64_t batch, int64_t lda,
                             int64_t ldb, void *stream) {
    if (!p || !A || !B) return fail(VEILS_ERR_INVALID, "null argument");
    if (op != 1 && op != 2) return fail(VEILS_ERR_INVALID, "op must be 1 (A += i B) or 2 (A = -i B)");
    if (P < 0 || batch < 0 || lda < P || ldb < P) return fail(VEILS_ERR_INVALID, "lda, ldb must be >= P >= 0");
    Plan &pl = p->pl;
    int32_t rc = ensure_device(pl, false);
    if (rc != VEILS_OK) return rc;
    PlanDeviceScope scope(pl.device);
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce = pl.precision == F32 ? complex_op<float>((float *)A, (const float *)B, batch * pl.F, P, lda, ldb, op, st)
                                         : complex_op<double>((double *)A, (const double *)B, batch * pl.F, P, lda, ldb, op, st);
    if (ce != cudaSuccess) return cuda_fail(ce, "complex op launch");
    return VEILS_OK;
}

int32_t veils_istft_dev(veils_plan *p, const void *S, int64_t P, int64_t batch, int64_t ldp,
                        const int64_t *k0, const int64_t *k1, void *x_out, int64_t out_pitch,
                        void *stream) {
    if (!p || !S || !x_out) return fail(VEILS_ERR_INVALID, "null argument");
    Plan &pl = p->pl;
    if (batch < 0) return fail(VEILS_ERR_INVALID, "batch must be >= 0");
    IstftRange r;
    std::string e = plan_istft_range(pl, P, k0, k1, &r);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    if (ldp < P) return fail(VEILS_ERR_INVALID, "ldp must be >= P");
    if (out_pitch < r.k1 - r.k0) return fail(VEILS_ERR_INVALID, "out_pitch must be >= k1 - k0");
    if (batch == 0) return VEILS_OK;
    int32_t rc = ensure_device(pl, true);                    // dual_win()? (src/lib.rs:849)
    if (rc != VEILS_OK) return rc;
    PlanDeviceScope scope(pl.device);
    InvParams prm;
    prm.P = P; prm.ldp = ldp; prm.batch = batch;
    prm.N = pl.N; prm.H = pl.H; prm.M = pl.M; prm.F = pl.F; prm.mid = pl.mid; prm.ps = pl.ps;
    prm.p_min = pl.p_min; prm.k0 = r.k0; prm.k1 = r.k1; prm.q0 = r.q0; prm.q1 = r.q1;
    prm.out_pitch = out_pitch; prm.mode = pl.mode; prm.fac2x = pl.fac2x;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t ce;
    if (pl.precision == F32) {
        Tables<float> tb = tables_of<float>(pl);
        ce = pl.wide_inv ? wide_inverse(S, (float *)x_out, prm, tb.wide_inv, pl.wide_tf, st)
             : pl.pk_inv ? pk_inverse(S, (float *)x_out, prm, tb, st)
             : pl.pow2_inv ? pow2_inverse<float>(S, (float *)x_out, prm, tb, st)
             : pl.smooth_inv ? smooth_inverse<float>(S, (float *)x_out, prm, tb.smooth_inv, st)
                           : dft_generic_inverse<float>(S, (float *)x_out, prm, tb, st);
    } else {
        Tables<double> tb = tables_of<double>(pl);
        ce = pl.pow2_inv ? pow2_inverse<double>(S, (double *)x_out, prm, tb, st)
             : pl.smooth_inv ? smooth_inverse<double>(S, (double *)x_out, prm, tb.smooth_inv, st)
                          : dft_generic_inverse<double>(S, (double *)x_out, prm, tb, st);
    }
    if (ce != cudaSuccess) return cuda_fail(ce, "istft launch");
    return VEILS_OK;
}

// ---- NaN screen (src/lib.rs:696-698 rejects NaN input in the one-sided modes) ----------------
// A NaN sample makes EVERY bin of the slices it touches NaN, and bin 0 is a plain sum of the
// windowed samples (no twiddle can hide it).  So the screen reads only the DC row of the finished
// S -- 1 / f_pts of the output, < 1 % of the transform's traffic, any kernel family -- and where
// that value is not finite (NaN or Inf input, or overflow) a thread re-reads the slice's samples and
// tests them exactly for NaN.  The flag may live in device memory or in page-locked host memory.
extern "C++" {
template <typename T>
__global__ void nan_dc_check_kernel(const T *__restrict__ x, int64_t n, int64_t x_pitch, int64_t batch, int64_t p0,
                                    int64_t P, int64_t H, int64_t N, int64_t mid, int64_t k_offset,
                                    const T *__restrict__ S, int64_t row_elems, int64_t sig_elems, int esz_mul,
                                    int *flag) {
    const int64_t total = batch * P;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / P, p = i - b * P;
        const T v = S[b * sig_elems + p * esz_mul];                  // Re(S[b][0][p]) (or |X|^2 of the spectrogram)
        (void)row_elems;
        if (v - v == (T)0) continue;                                 // finite: no NaN went into this slice
        const int64_t s0 = (p0 + p) * H + k_offset - mid;
        const T *xb = x + b * x_pitch;
        bool bad = false;
        for (int64_t j = 0; j < N; ++j) {
            const int64_t s = s0 + j;
            if (s >= 0 && s < n) { const T w = xb[s]; bad |= (w != w); }
        }
        if (bad) *flag = 1;
    }
}
}  // extern "C++"
// enqueue the screen behind a forward transform (one-sided modes only; no-op otherwise)
static int32_t nan_check_launch(const Plan &pl, const void *x, int64_t n, int64_t batch, int64_t x_pitch, int64_t a,
                                int64_t P, int64_t k_offset, const void *S, int64_t ldp, int spectrogram, int *flag,
                                cudaStream_t st) {
    if (pl.mode < ONESIDED || !flag || batch <= 0 || P <= 0) return VEILS_OK;
    PlanDeviceScope scope(pl.device);
    const int64_t total = batch * P;
    const int grid = (int)std::min<int64_t>(148 * 8, (total + 255) / 256);
    const int mul = spectrogram ? 1 : 2;
    if (pl.precision == F32)
        nan_dc_check_kernel<float><<<grid, 256, 0, st>>>((const float *)x, n, x_pitch, batch, a, P, pl.H, pl.N, pl.mid,
                                                         k_offset, (const float *)S, ldp * mul, pl.F * ldp * mul, mul, flag);
    else
        nan_dc_check_kernel<double><<<grid, 256, 0, st>>>((const double *)x, n, x_pitch, batch, a, P, pl.H, pl.N, pl.mid,
                                                          k_offset, (const double *)S, ldp * mul, pl.F * ldp * mul, mul, flag);
    ++g_launch_count;
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) return cuda_fail(ce, "NaN screen launch");
    return VEILS_OK;
}

// ---- host-pointer flavour ------------------------------------------------------------------
static size_t esz(const Plan &pl) { return pl.precision == F32 ? 4 : 8; }
// device row pitch of S in elements: 128-byte rows, so that the TMA paths of the kernels apply
static int64_t host_ldp(int64_t P, size_t elem_bytes) {
    // 128-byte rows on the device (pitched 2-D copies measured as fast as 1-D ones over PCIe), so that the
    // host flavour runs the SAME kernels as the pitch-aligned device path and returns the same bits
    static const bool packed = [] { const char *e = getenv("VEILS_HOST_LDP_PACKED"); return e && atoi(e) != 0; }();
    if (packed) return P;
    const int64_t al = (int64_t)(128 / elem_bytes);
    return (P + al - 1) / al * al;
}
// rows per staged chunk: about 32 MB of S (enough to fill the GPU, small enough to pipeline)
static int64_t host_chunk_rows(int64_t batch, size_t s_row_bytes) {
    const size_t target = (size_t)32 << 20;
    int64_t rows = (int64_t)std::max<size_t>(1, target / std::max<size_t>(1, s_row_bytes));
    return std::min(rows, batch);
}

static int32_t forward_host(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                            const int64_t *p0, const int64_t *p1, int64_t k_offset, void *out,
                            int spectrogram) {
    if (!p || !x || !out) return fail(VEILS_ERR_INVALID, "null argument");
    Plan &pl = p->pl;
    if (x_pitch < n) return fail(VEILS_ERR_INVALID, "x_pitch must be >= n");
    const int64_t m2p = pl.N - pl.mid;
    if (n < m2p)
        return fail(VEILS_ERR_INVALID, "Input length " + std::to_string((long long)n) +
                                           " must be >= ceil(m_num/2) = " + std::to_string((long long)m2p));
    int64_t a, b;
    std::string e = plan_p_range(pl, n, p0, p1, &a, &b);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    if (batch <= 0) return VEILS_OK;
    const int64_t P = b - a;
    const size_t es = esz(pl);
    const size_t oes = es * (spectrogram ? 1 : 2);                       // output element
    const int64_t ldp = host_ldp(P, oes);
    const size_t in_row = (size_t)n * es, out_row = (size_t)pl.F * (size_t)ldp * oes;
    const int64_t rows = host_chunk_rows(batch, out_row);
    CU(cudaSetDevice(pl.device));
    HostStage &hs = p->stage;
    std::lock_guard<std::mutex> lock(hs.mu);
    CU(hs.streams());
    *hs.nan_flag = 0;
    const int slots = rows < batch ? 2 : 1;
    for (int sl = 0; sl < slots; ++sl) {
        CU(hs.reserve(2 * sl, (size_t)rows * in_row));
        CU(hs.reserve(2 * sl + 1, (size_t)rows * out_row));
    }
    // rows [r0, r0 + nr) per chunk, alternating between two streams: the upload of one chunk, the
    // transform of the previous and the download of the one before overlap (when the host buffers
    // are page-locked; pageable copies are staged by the driver and serialise)
    int32_t rc = VEILS_OK;
    int64_t c = 0;
    for (int64_t r0 = 0; r0 < batch && rc == VEILS_OK; r0 += rows, ++c) {
        const int64_t nr = std::min(rows, batch - r0);
        const int sl = (int)(c & 1) % slots;
        cudaStream_t st = hs.st[sl];
        void *dx = hs.buf[2 * sl], *ds = hs.buf[2 * sl + 1];
        cudaError_t ce = cudaMemcpy2DAsync(dx, in_row, (const char *)x + (size_t)r0 * (size_t)x_pitch * es,
                                           (size_t)x_pitch * es, in_row, (size_t)nr, cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "H2D copy"); break; }
        rc = forward_impl(p, dx, n, nr, n, &a, &b, k_offset, ds, ldp, st, spectrogram);
        if (rc != VEILS_OK) break;
        rc = nan_check_launch(pl, dx, n, nr, n, a, P, k_offset, ds, ldp, spectrogram, hs.nan_flag, st);
        if (rc != VEILS_OK) break;
        ce = cudaMemcpy2DAsync((char *)out + (size_t)r0 * (size_t)pl.F * (size_t)P * oes, (size_t)P * oes, ds,
                               (size_t)ldp * oes, (size_t)P * oes, (size_t)(nr * pl.F), cudaMemcpyDeviceToHost, st);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "D2H copy");
    }
    for (int sl = 0; sl < 2; ++sl) {
        cudaError_t ce = cudaStreamSynchronize(hs.st[sl]);
        if (ce != cudaSuccess && rc == VEILS_OK) rc = cuda_fail(ce, "stft (host)");
    }
    if (rc == VEILS_OK && *hs.nan_flag)
        return fail(VEILS_ERR_INVALID, "Complex-valued input not allowed for one-sided FFT modes");
    return rc;
}

int32_t veils_stft_host(veils_plan *p, const void *x, int64_t n, int64_t batch, int64_t x_pitch,
                        const int64_t *p0, const int64_t *p1, int64_t k_offset, void *S) {
    return forward_host(p, x, n, batch, x_pitch, p0, p1, k_offset, S, 0);
}
int32_t veils_spectrogram_host(veils_plan *p, const void *x, int64_t n, int64_t batch,
                               int64_t x_pitch, const int64_t *p0, const int64_t *p1,
                               int64_t k_offset, void *Sxx) {
    return forward_host(p, x, n, batch, x_pitch, p0, p1, k_offset, Sxx, 1);
}

int32_t veils_istft_host(veils_plan *p, const void *S, int64_t P, int64_t batch, const int64_t *k0,
                         const int64_t *k1, void *x_out) {
    if (!p || !S || !x_out) return fail(VEILS_ERR_INVALID, "null argument");
    Plan &pl = p->pl;
    IstftRange r;
    std::string e = plan_istft_range(pl, P, k0, k1, &r);
    if (!e.empty()) return fail(VEILS_ERR_INVALID, e);
    if (batch <= 0) return VEILS_OK;
    const size_t es = esz(pl);
    const int64_t len = r.k1 - r.k0;
    const int64_t ldp = host_ldp(P, 2 * es);
    const size_t s_row = (size_t)pl.F * (size_t)ldp * es * 2, o_row = (size_t)len * es;
    const int64_t rows = host_chunk_rows(batch, s_row);
    CU(cudaSetDevice(pl.device));
    HostStage &hs = p->stage;
    std::lock_guard<std::mutex> lock(hs.mu);
    CU(hs.streams());
    const int slots = rows < batch ? 2 : 1;
    for (int sl = 0; sl < slots; ++sl) {
        CU(hs.reserve(2 * sl + 1, (size_t)rows * s_row));      // the big buffer is slot's [1] in both directions
        CU(hs.reserve(2 * sl, (size_t)rows * o_row));
    }
    int32_t rc = VEILS_OK;
    int64_t c = 0;
    for (int64_t r0 = 0; r0 < batch && rc == VEILS_OK; r0 += rows, ++c) {
        const int64_t nr = std::min(rows, batch - r0);
        const int sl = (int)(c & 1) % slots;
        cudaStream_t st = hs.st[sl];
        void *ds = hs.buf[2 * sl + 1], *dx = hs.buf[2 * sl];
        cudaError_t ce = cudaMemcpy2DAsync(ds, (size_t)ldp * es * 2,
                                           (const char *)S + (size_t)r0 * (size_t)pl.F * (size_t)P * es * 2,
                                           (size_t)P * es * 2, (size_t)P * es * 2, (size_t)(nr * pl.F),
                                           cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) { rc = cuda_fail(ce, "H2D copy"); break; }
        rc = veils_istft_dev(p, ds, P, nr, ldp, &r.k0, &r.k1, dx, len, st);
        if (rc != VEILS_OK) break;
        ce = cudaMemcpyAsync((char *)x_out + (size_t)r0 * o_row, dx, (size_t)nr * o_row, cudaMemcpyDeviceToHost, st);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "D2H copy");
    }
    for (int sl = 0; sl < 2; ++sl) {
        cudaError_t ce = cudaStreamSynchronize(hs.st[sl]);
        if (ce != cudaSuccess && rc == VEILS_OK) rc = cuda_fail(ce, "istft (host)");
    }
    return rc;
}

// ---- device buffers / streams ---------------------------------------------------------------
int32_t veils_device_count(int32_t *out) {
    int n = 0;
    CU(cudaGetDeviceCount(&n));
    *out = n;
    return VEILS_OK;
}
int32_t veils_dev_alloc(int32_t device, uint64_t bytes, void **out) {
    CU(cudaSetDevice(device));
    CU(cudaMalloc(out, bytes));
    return VEILS_OK;
}
int32_t veils_dev_free(int32_t device, void *ptr) {
    CU(cudaSetDevice(device));
    CU(cudaFree(ptr));
    return VEILS_OK;
}
int32_t veils_dev_upload(void *dst, const void *src, uint64_t bytes, void *stream) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return VEILS_OK;
}
int32_t veils_dev_download(void *dst, const void *src, uint64_t bytes, void *stream) {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return VEILS_OK;
}
int32_t veils_host_alloc_pinned(uint64_t bytes, void **out) {
    CU(cudaMallocHost(out, bytes));
    return VEILS_OK;
}
int32_t veils_host_free_pinned(void *ptr) {
    CU(cudaFreeHost(ptr));
    return VEILS_OK;
}
int32_t veils_stream_create(int32_t device, void **out) {
    CU(cudaSetDevice(device));
    cudaStream_t s;
    CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (void *)s;
    return VEILS_OK;
}
int32_t veils_stream_destroy(void *stream) {
    CU(cudaStreamDestroy((cudaStream_t)stream));
    return VEILS_OK;
}
int32_t veils_stream_sync(void *stream) {
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return VEILS_OK;
}

uint64_t veils_kernel_launch_count(void) { return (uint64_t)g_launch_count.load(); }
const char *veils_plan_kernel_name(const veils_plan *p, int32_t inverse) {
    if (inverse ? p->pl.wide_inv : p->pl.wide) return "pow2_wide";
    if (inverse ? p->pl.pk_inv : p->pl.pk_fwd) return "pow2_packed";
    if (inverse ? p->pl.pow2_inv : p->pl.pow2_fwd) return "pow2_tile";
    return (inverse ? p->pl.smooth_inv : p->pl.smooth_fwd) ? "mixed_radix" : "dft_generic";
}

}  // extern "C"
