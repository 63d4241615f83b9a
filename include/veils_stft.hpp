// veils_stft.hpp -- header-only C++17 host class over the C ABI (veils_cuda.h) with the shape of the
// reference's Rust API (src/lib.rs): the same method names, argument order, `Option` defaults and
// `Result<_, String>` error style, so that code (and tests) written against veils::StandaloneSTFT
// reads the same against the B200 engine.  The Rust crate itself cannot be built in this image (no
// cargo); rust/src/lib.rs is the same mapping in Rust, this header is the one that is compiled
// and exercised (tests/cpp/integration_tests.cpp, a restatement of tests/integration_tests.rs).
//
//   Rust                                         here
//   StandaloneSTFT::new(win, hop, fs, Some("onesided"), None, None, None)   (:166-174)
//                                                StandaloneSTFT::create(win, hop, fs, "onesided", {}, {}, {})
//   Result<T, String>: is_ok / is_err / unwrap   veils::Result<T>: is_ok / is_err / unwrap / error
//   stft(&x, None, None, None) -> Vec<Vec<Complex<f64>>> [freq][time]      (:689-749)
//   istft(&S, None, None) -> Vec<f64>                                       (:792-914)
// All arithmetic runs on the GPU (f64 by default, like the reference); there is no CPU path.
#pragma once
#include <complex>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "veils_cuda.h"

namespace veils {

template <typename T>
class Result {
  public:
    static Result ok(T v) { Result r; r.val_ = std::move(v); return r; }
    static Result err(std::string e) { Result r; r.err_ = std::move(e); return r; }
    bool is_ok() const { return val_.has_value(); }
    bool is_err() const { return !val_.has_value(); }
    const std::string &error() const { return err_; }
    T unwrap() {                                   // like Rust: panics (throws) on Err
        if (!val_) throw std::runtime_error("called `Result::unwrap()` on an `Err` value: " + err_);
        return std::move(*val_);
    }
  private:
    std::optional<T> val_;
    std::string err_;
};

using Spectrum = std::vector<std::vector<std::complex<double>>>;   // [freq][time], src/lib.rs:735-746

class StandaloneSTFT {
  public:
    // src/lib.rs:166-174 (+ scale_to, scipy keyword; + precision / device of the engine)
    static Result<StandaloneSTFT> create(std::vector<double> win, size_t hop, double fs,
                                         std::optional<std::string> fft_mode = std::nullopt,
                                         std::optional<size_t> mfft = std::nullopt,
                                         std::optional<std::vector<double>> dual_win = std::nullopt,
                                         std::optional<int> phase_shift = std::nullopt,
                                         std::optional<std::string> scale_to = std::nullopt, int device = 0) {
        if (dual_win && dual_win->size() != win.size())
            return Result<StandaloneSTFT>::err("dual_win.len() must equal win.len()!");       // :197-199
        veils_plan_desc d{};
        d.win = win.empty() ? nullptr : win.data();
        d.m_num = (int64_t)win.size();
        d.hop = (int64_t)hop;
        d.fs = fs;
        d.fft_mode = fft_mode ? fft_mode->c_str() : nullptr;
        d.mfft = mfft ? (int64_t)*mfft : 0;
        d.dual_win = dual_win ? dual_win->data() : nullptr;
        d.scale_to = scale_to ? scale_to->c_str() : nullptr;
        d.phase_shift = phase_shift ? *phase_shift : 0;
        d.precision = VEILS_F64;
        d.device = device;
        veils_plan *p = nullptr;
        if (veils_plan_create(&d, &p) != VEILS_OK) return Result<StandaloneSTFT>::err(veils_last_error());
        return Result<StandaloneSTFT>::ok(StandaloneSTFT(p));
    }
    StandaloneSTFT(StandaloneSTFT &&o) noexcept : p_(o.p_) { o.p_ = nullptr; }
    StandaloneSTFT &operator=(StandaloneSTFT &&o) noexcept {
        if (this != &o) { reset(); p_ = o.p_; o.p_ = nullptr; }
        return *this;
    }
    StandaloneSTFT(const StandaloneSTFT &) = delete;
    StandaloneSTFT &operator=(const StandaloneSTFT &) = delete;
    ~StandaloneSTFT() { reset(); }

    // accessors, src/lib.rs:227-281
    std::vector<double> win() const { std::vector<double> w((size_t)m_num()); veils_win(p_, w.data()); return w; }
    size_t hop() const { return (size_t)veils_hop(p_); }
    double fs() const { return veils_fs(p_); }
    size_t mfft() const { return (size_t)veils_mfft(p_); }
    int phase_shift() const { return (int)veils_phase_shift(p_); }
    double sampling_period() const { return 1.0 / fs(); }
    double delta_t() const { return veils_delta_t(p_); }
    double delta_f() const { return veils_delta_f(p_); }
    size_t m_num() const { return (size_t)veils_m_num(p_); }
    size_t m_num_mid() const { return (size_t)veils_m_num_mid(p_); }
    size_t f_pts() const { return (size_t)veils_f_pts(p_); }
    bool onesided_fft() const { return veils_onesided_fft(p_) != 0; }
    // src/lib.rs:324-333
    Result<std::vector<double>> dual_win() {
        std::vector<double> d(m_num());
        if (veils_dual_win(p_, d.data()) != VEILS_OK) return Result<std::vector<double>>::err(veils_last_error());
        return Result<std::vector<double>>::ok(std::move(d));
    }
    bool invertible() { return veils_invertible(p_) != 0; }
    // src/lib.rs:587-609
    int p_min() const { return (int)veils_p_min(p_); }
    int k_min() const { return (int)veils_k_min(p_); }
    Result<int> p_max(size_t n) const {            // the reference panics for n < ceil(m_num/2) (:562-564)
        int64_t v = 0;
        if (veils_p_max(p_, (int64_t)n, &v) != VEILS_OK) return Result<int>::err(veils_last_error());
        return Result<int>::ok((int)v);
    }
    Result<std::pair<int, int>> p_range(size_t n, std::optional<int> p0, std::optional<int> p1) const {
        int64_t a = 0, b = 0, a_in = p0.value_or(0), b_in = p1.value_or(0);
        if (veils_p_range(p_, (int64_t)n, p0 ? &a_in : nullptr, p1 ? &b_in : nullptr, &a, &b) != VEILS_OK)
            return Result<std::pair<int, int>>::err(veils_last_error());
        return Result<std::pair<int, int>>::ok({(int)a, (int)b});
    }
    // src/lib.rs:916-949
    Result<std::vector<double>> t(size_t n, std::optional<int> p0, std::optional<int> p1, std::optional<int> k_offset) const {
        auto r = p_range(n, p0, p1);
        if (r.is_err()) return Result<std::vector<double>>::err(r.error());
        const auto ab = r.unwrap();
        std::vector<double> out((size_t)(ab.second - ab.first));
        int64_t a = ab.first, b = ab.second;
        if (veils_t(p_, (int64_t)n, &a, &b, k_offset.value_or(0), out.data()) != VEILS_OK)
            return Result<std::vector<double>>::err(veils_last_error());
        return Result<std::vector<double>>::ok(std::move(out));
    }
    std::vector<double> f() const { std::vector<double> out(f_pts()); veils_f(p_, out.data()); return out; }

    // src/lib.rs:689-749
    Result<Spectrum> stft(const std::vector<double> &x, std::optional<int> p0 = std::nullopt,
                          std::optional<int> p1 = std::nullopt, std::optional<int> k_offset = std::nullopt) const {
        auto r = p_range(x.size(), p0, p1);
        if (x.size() < m_num() - m_num_mid())
            return Result<Spectrum>::err("Input length " + std::to_string(x.size()) + " must be >= ceil(m_num/2) = " +
                                         std::to_string(m_num() - m_num_mid()));                      // :700-707
        if (r.is_err()) return Result<Spectrum>::err(r.error());
        const auto ab = r.unwrap();
        const size_t F = f_pts(), P = (size_t)(ab.second - ab.first);
        std::vector<std::complex<double>> flat(F * P);
        int64_t a = ab.first, b = ab.second;
        if (veils_stft_host(p_, x.data(), (int64_t)x.size(), 1, (int64_t)x.size(), &a, &b, k_offset.value_or(0),
                            flat.data()) != VEILS_OK)
            return Result<Spectrum>::err(veils_last_error());
        Spectrum S(F);
        for (size_t fi = 0; fi < F; ++fi) S[fi].assign(flat.begin() + fi * P, flat.begin() + (fi + 1) * P);
        return Result<Spectrum>::ok(std::move(S));
    }
    // scipy ShortTimeFFT.spectrogram (not in the reference): |stft|^2, [freq][time]
    Result<std::vector<std::vector<double>>> spectrogram(const std::vector<double> &x) const {
        using R = Result<std::vector<std::vector<double>>>;
        auto r = p_range(x.size(), std::nullopt, std::nullopt);
        if (r.is_err()) return R::err(r.error());
        const auto ab = r.unwrap();
        const size_t F = f_pts(), P = (size_t)(ab.second - ab.first);
        std::vector<double> flat(F * P);
        if (veils_spectrogram_host(p_, x.data(), (int64_t)x.size(), 1, (int64_t)x.size(), nullptr, nullptr, 0,
                                   flat.data()) != VEILS_OK)
            return R::err(veils_last_error());
        std::vector<std::vector<double>> S(F);
        for (size_t fi = 0; fi < F; ++fi) S[fi].assign(flat.begin() + fi * P, flat.begin() + (fi + 1) * P);
        return R::ok(std::move(S));
    }
    // src/lib.rs:792-914
    Result<std::vector<double>> istft(const Spectrum &S, std::optional<int> k0 = std::nullopt,
                                      std::optional<int> k1 = std::nullopt) {
        using R = Result<std::vector<double>>;
        if (S.empty() || S[0].empty()) return R::err("STFT data cannot be empty");                    // :798-800
        if (S.size() != f_pts())
            return R::err("STFT frequency dimension " + std::to_string(S.size()) + " must equal " + std::to_string(f_pts()));
        const size_t F = S.size(), P = S[0].size();
        std::vector<std::complex<double>> flat(F * P);
        for (size_t fi = 0; fi < F; ++fi) {
            if (S[fi].size() != P) return R::err("STFT rows must have equal length");
            std::copy(S[fi].begin(), S[fi].end(), flat.begin() + fi * P);
        }
        int64_t a_in = k0.value_or(0), b_in = k1.value_or(0), a = 0, b = 0;
        if (veils_istft_range(p_, (int64_t)P, k0 ? &a_in : nullptr, k1 ? &b_in : nullptr, &a, &b) != VEILS_OK)
            return R::err(veils_last_error());
        std::vector<double> y((size_t)(b - a));
        if (veils_istft_host(p_, flat.data(), (int64_t)P, 1, &a, &b, y.data()) != VEILS_OK) return R::err(veils_last_error());
        return R::ok(std::move(y));
    }

  private:
    explicit StandaloneSTFT(veils_plan *p) : p_(p) {}
    void reset() { if (p_) veils_plan_destroy(p_); p_ = nullptr; }
    veils_plan *p_ = nullptr;
};

}  // namespace veils
