// plan.cpp -- host planner: validation, padding geometry, canonical dual window, scaling.
// Restates the integer / window logic of the reference (src/lib.rs:166-224, 284-333,
// 524-609, 798-847) with 64-bit indices; where the Rust code and its Python twin differ
// the choice is documented in DESIGN.md ("divergences").
#include "plan.hpp"

#include <cmath>
#include <cstring>
#include <limits>

namespace veils {

static std::string i2s(int64_t v) { return std::to_string((long long)v); }

std::string plan_init(Plan &pl, const double *win, int64_t m_num, int64_t hop, double fs,
                      const char *fft_mode, int64_t mfft, const double *dual_win,
                      const char *scale_to, int64_t phase_shift, int precision, int device) {
    // src/lib.rs:176-186
    if (win == nullptr || m_num <= 0) return "Parameter win must be non-empty!";
    for (int64_t i = 0; i < m_num; ++i)
        if (!std::isfinite(win[i])) return "Parameter win must have finite entries!";
    if (hop < 1) return "Parameter hop=" + i2s(hop) + " is not >= 1!";
    // src/lib.rs:66-78, default "onesided" (:192)
    const char *m = fft_mode ? fft_mode : "onesided";
    if (!std::strcmp(m, "twosided")) pl.mode = TWOSIDED;
    else if (!std::strcmp(m, "centered")) pl.mode = CENTERED;
    else if (!std::strcmp(m, "onesided")) pl.mode = ONESIDED;
    else if (!std::strcmp(m, "onesided2X")) pl.mode = ONESIDED2X;
    else return std::string("Unknown fft_mode: ") + m;

    pl.N = m_num;
    pl.H = hop;
    pl.fs = fs;
    pl.M = mfft > 0 ? mfft : m_num;                       // src/lib.rs:193
    pl.mid = m_num / 2;
    pl.phase_shift = phase_shift;
    // scipy _short_time_fft.py:957-965, :1080-1082 -- the reference omits both checks and
    // would silently truncate (mfft < m_num) or panic (negative roll); we report them.
    if (pl.M < pl.N)
        return "Attribute mfft=" + i2s(pl.M) + " needs to be at least the window length m_num=" +
               i2s(pl.N) + "!";
    if (!(-pl.M < phase_shift && phase_shift < pl.M))
        return "-mfft < phase_shift < mfft does not hold for mfft=" + i2s(pl.M) +
               ", phase_shift=" + i2s(phase_shift) + "!";
    if (mfft < 0) return "Parameter mfft=" + i2s(mfft) + " is not >= 0!";
    pl.F = (pl.mode == ONESIDED || pl.mode == ONESIDED2X) ? pl.M / 2 + 1 : pl.M;
    pl.ps = ((phase_shift + pl.mid) % pl.N + pl.N) % pl.N;   // true modulus (twin :236)
    if (precision != F32 && precision != F64) return "precision must be VEILS_F32 or VEILS_F64";
    pl.precision = precision;
    pl.device = device;

    pl.win.assign(win, win + m_num);
    if (dual_win) {                                       // src/lib.rs:197-207
        for (int64_t i = 0; i < m_num; ++i)
            if (!std::isfinite(dual_win[i])) return "Parameter dual_win must have finite entries!";
        pl.dual.assign(dual_win, dual_win + m_num);
        pl.dual_state = 1;
    }

    // scale_to: scipy _short_time_fft.py:1028-1039 with fac_magnitude (:1553) / fac_psd (:1575)
    pl.scaling = SCALE_NONE;
    if (scale_to) {
        double fac;
        if (!std::strcmp(scale_to, "magnitude")) {
            double s = 0;
            for (double w : pl.win) s += w;
            fac = 1.0 / std::fabs(s);
            pl.scaling = SCALE_MAGNITUDE;
        } else if (!std::strcmp(scale_to, "psd")) {
            double s = 0;
            for (double w : pl.win) s += w * w;
            fac = 1.0 / std::sqrt(s * fs);
            pl.scaling = SCALE_PSD;
        } else {
            return std::string("scaling='") + scale_to + "' not in ['magnitude', 'psd']!";
        }
        if (!std::isfinite(fac)) return "Parameter win cannot be scaled (zero sum)!";
        for (double &w : pl.win) w *= fac;
        if (pl.dual_state == 1)
            for (double &d : pl.dual) d /= fac;
    }
    pl.fac2x = (pl.scaling == SCALE_PSD) ? std::sqrt(2.0) : 2.0;

    // zeros at the window borders shorten the padding (src/lib.rs:537-541 / :576, twin :160, :205)
    pl.lead_zeros = 0;
    while (pl.lead_zeros < pl.N && pl.win[pl.lead_zeros] * pl.win[pl.lead_zeros] == 0.0)
        ++pl.lead_zeros;
    pl.trail_zeros = 0;
    while (pl.trail_zeros < pl.N &&
           pl.win[pl.N - 1 - pl.trail_zeros] * pl.win[pl.N - 1 - pl.trail_zeros] == 0.0)
        ++pl.trail_zeros;

    // pre_padding (src/lib.rs:524-558 as corrected by the twin :153-163): slide the window to
    // the left in steps of hop until it no longer overlaps t >= 0.
    {
        const int64_t n0 = -pl.mid;
        bool found = false;
        int64_t p_ = 0;
        for (int64_t n_ = n0; n_ > n0 - pl.N - 1; n_ -= pl.H, ++p_) {
            const int64_t n_next = n_ - pl.H;
            const int64_t tail = std::min<int64_t>(-n_next, pl.N);   // len(w2[n_next:])
            if (n_next + pl.N <= 0 || tail <= pl.trail_zeros) {
                pl.k_min = n_;
                pl.p_min = -p_;
                found = true;
                break;
            }
        }
        if (!found) {
            pl.k_min = n0 - pl.N;
            pl.p_min = -(pl.N / pl.H + 1);
        }
    }
    return "";
}

// post_padding (src/lib.rs:560-585): slide right from q1 = n / hop until no overlap with t < n.
std::string plan_post_padding(const Plan &pl, int64_t n, int64_t *k_max, int64_t *p_max) {
    const int64_t m2p = pl.N - pl.mid;
    if (n < m2p) return "Parameter n must be >= ceil(m_num/2) = " + i2s(m2p) + "!";
    const int64_t q1 = n / pl.H;
    const int64_t k1 = q1 * pl.H - pl.mid;
    int64_t q_ = q1;
    for (int64_t k_ = k1; k_ < n + pl.N; k_ += pl.H, ++q_) {
        const int64_t n_next = k_ + pl.H;
        if (n_next >= n || std::min<int64_t>(n - n_next, pl.N) <= pl.lead_zeros) {
            if (k_max) *k_max = k_ + pl.N;
            if (p_max) *p_max = q_ + 1;
            return "";
        }
    }
    if (k_max) *k_max = n + pl.N;
    if (p_max) *p_max = (n + pl.mid) / pl.H + 1;
    return "";
}

// p_range (src/lib.rs:595-609).  As in the reference there is no upper bound on p1.
std::string plan_p_range(const Plan &pl, int64_t n, const int64_t *p0, const int64_t *p1,
                         int64_t *o0, int64_t *o1) {
    int64_t pmax = 0;
    std::string e = plan_post_padding(pl, n, nullptr, &pmax);
    if (!e.empty()) return e;
    const int64_t a = p0 ? *p0 : pl.p_min;
    const int64_t b = p1 ? *p1 : pmax;
    if (!(pl.p_min <= a && a < b)) return "Invalid slice range: p0=" + i2s(a) + ", p1=" + i2s(b);
    *o0 = a;
    *o1 = b;
    return "";
}

// calc_dual_canonical_window (src/lib.rs:284-322), same accumulation order.
std::string plan_dual(Plan &pl) {
    std::lock_guard<std::mutex> g(pl.mu);
    if (pl.dual_state == 1) return "";
    if (pl.dual_state == 2) return pl.dual_err;
    const int64_t N = pl.N, H = pl.H;
    if (H > N) {
        pl.dual_err = "hop=" + i2s(H) + " is larger than window length of " + i2s(N) +
                      " => STFT not invertible!";
        pl.dual_state = 2;
        return pl.dual_err;
    }
    std::vector<double> w2(N), dd(N);
    for (int64_t i = 0; i < N; ++i) dd[i] = w2[i] = pl.win[i] * pl.win[i];
    for (int64_t k = H; k < N; k += H) {
        for (int64_t i = k; i < N; ++i) dd[i] += w2[i - k];
        for (int64_t i = 0; i < N - k; ++i) dd[i] += w2[i + k];
    }
    double mx = 0;
    for (double d : dd) mx = std::max(mx, d);
    const double res = std::numeric_limits<double>::epsilon() * mx;   // src/lib.rs:313-314
    for (double d : dd)
        if (!(d >= res)) {
            pl.dual_err = "Short-time Fourier Transform not invertible!";
            pl.dual_state = 2;
            return pl.dual_err;
        }
    pl.dual.resize(N);
    for (int64_t i = 0; i < N; ++i) pl.dual[i] = pl.win[i] / dd[i];
    pl.dual_state = 1;
    return "";
}

// istft argument validation (src/lib.rs:798-847; k_min check and floor division from the twin
// python/standalone_stft.py:392, :398).
std::string plan_istft_range(Plan &pl, int64_t P, const int64_t *k0p, const int64_t *k1p,
                             IstftRange *out) {
    if (P <= 0) return "STFT data cannot be empty";
    const int64_t n_min = pl.N - pl.mid;
    int64_t pm = 0;
    std::string e = plan_post_padding(pl, n_min, nullptr, &pm);
    if (!e.empty()) return e;
    const int64_t q_num = pm - pl.p_min;
    if (P < q_num)
        return "STFT time dimension " + i2s(P) + " needs to have at least " + i2s(q_num) + " slices";
    const int64_t q_max = P + pl.p_min;
    const int64_t k_max = (q_max - 1) * pl.H + pl.N - pl.mid;
    const int64_t k0 = k0p ? *k0p : 0;
    const int64_t k1 = k1p ? *k1p : k_max;
    if (k0 >= k1) return "k0=" + i2s(k0) + " must be < k1=" + i2s(k1);
    if (k0 < pl.k_min) return "k0=" + i2s(k0) + " must be >= k_min=" + i2s(pl.k_min);
    if (k1 - k0 < n_min)
        return "Output length " + i2s(k1 - k0) + " has to be at least " + i2s(n_min);
    int64_t pm1 = 0;
    e = plan_post_padding(pl, k1, nullptr, &pm1);
    if (!e.empty()) return e;
    out->k0 = k0;
    out->k1 = k1;
    out->q0 = k0 >= 0 ? k0 / pl.H + pl.p_min : floordiv(k0, pl.H);
    out->q1 = std::min(pm1, q_max);
    return "";
}

}  // namespace veils
