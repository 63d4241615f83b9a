// plan.hpp -- host-side planning state of one StandaloneSTFT instance.
//
// Mirrors the fields of `struct StandaloneSTFT` (reference src/lib.rs:116-127) plus what the
// reference recomputes on every call (pre/post padding, src/lib.rs:524-585) and what the CUDA
// kernels need (device tables, dispatch decision).  Pure host C++: usable without a GPU.
#pragma once
#include <atomic>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

namespace veils {

enum Mode { TWOSIDED = 0, CENTERED = 1, ONESIDED = 2, ONESIDED2X = 3 };
enum Scale { SCALE_NONE = 0, SCALE_MAGNITUDE = 1, SCALE_PSD = 2 };
enum Prec { F32 = 0, F64 = 1 };

struct DeviceTables;   // defined in api.cu (device copies of win / dual / twiddles)

struct Plan {
    // constructor arguments (after scale_to has been applied to win / dual)
    std::vector<double> win;
    int64_t N = 0;            // m_num
    int64_t H = 0;            // hop
    int64_t M = 0;            // mfft
    int64_t F = 0;            // f_pts
    int64_t mid = 0;          // m_num_mid
    int64_t phase_shift = 0;
    int64_t ps = 0;           // (phase_shift + mid) mod N, non-negative (src/lib.rs:343)
    double fs = 1.0;
    int mode = ONESIDED;
    int scaling = SCALE_NONE;
    int precision = F64;
    int device = 0;
    double fac2x = 2.0;       // onesided2X factor: 2, sqrt(2) under psd (scipy :2169)

    // geometry that depends on the window only
    int64_t lead_zeros = 0;   // number of leading zeros of |win|^2
    int64_t trail_zeros = 0;  // number of trailing zeros of |win|^2
    int64_t p_min = 0, k_min = 0;

    // dual window: given, or canonical computed on first use (src/lib.rs:324-329)
    std::mutex mu;
    int dual_state = 0;       // 0 = not yet computed, 1 = available, 2 = not invertible
    std::vector<double> dual;
    std::string dual_err;

    // device side, created lazily by the first compute call
    DeviceTables *dev = nullptr;
    std::atomic<int> dev_ready{0};   // bit 0: forward tables uploaded, bit 1: dual-window tables too (lock-free fast path)
    bool pow2_fwd = false;    // dispatch decisions (see api.cu)
    bool pow2_inv = false;
    bool pk_fwd = false;      // packed f32 family (preferred over pow2 when available)
    bool pk_inv = false;
    bool wide = false;        // wide f32 family (mfft 1024 / 2048 / 4096, default geometry): forward
    bool smooth_fwd = false;  // mixed-radix family (non-power-of-two mfft with prime factors <= 13)
    bool smooth_inv = false;
    int wide_tf = 0;          // slices per tile of the wide kernels, fixed at plan creation
    bool wide_inv = false;    // ... and inverse (one-sided layouts; two-sided ones stay with the packed family)
};

// Each returns "" on success or the reference's error text.
std::string plan_init(Plan &pl, const double *win, int64_t m_num, int64_t hop, double fs,
                      const char *fft_mode, int64_t mfft, const double *dual_win,
                      const char *scale_to, int64_t phase_shift, int precision, int device);
std::string plan_post_padding(const Plan &pl, int64_t n, int64_t *k_max, int64_t *p_max);
std::string plan_p_range(const Plan &pl, int64_t n, const int64_t *p0, const int64_t *p1,
                         int64_t *o0, int64_t *o1);
// ensures pl.dual is available; returns error text ("" ok). Thread-safe.
std::string plan_dual(Plan &pl);

struct IstftRange {
    int64_t k0, k1, q0, q1;
};
std::string plan_istft_range(Plan &pl, int64_t P, const int64_t *k0, const int64_t *k1,
                             IstftRange *out);

inline int64_t floordiv(int64_t a, int64_t b) {
    int64_t q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}

}  // namespace veils
