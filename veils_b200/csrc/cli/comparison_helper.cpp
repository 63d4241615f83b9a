// comparison_helper.cpp -- the reference's src/bin/python_rust_comparison_helper.rs protocol on
// top of the C ABI (include/veils_cuda.h), so python/final_comparison.py's round trip (:250-260,
// 1e-12) can be replayed with only the subprocess command changed:
//
//     veils_comparison_helper <input_json_file>
//
// input  {"signal":[f64], "window":[f64], "hop_length":int, "fs":f64, "fft_mode":str}   (:14-21)
// output {"stft":[[{"real","imag"}]]  /* [f_pts][time] */, "istft":[f64],
//         "properties":{"m_num","f_pts","p_min","p_max","mfft","hop","fs"}}              (:38-54)
// Errors go to stderr as "Error: <message>" with exit status 1, like the Rust binary's
// `Result<(), Box<dyn Error>>` main.  All arithmetic runs on the GPU in f64 (there is no CPU path).
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/veils_cuda.h"

namespace {

struct Input {
    std::vector<double> signal, window;
    long long hop = 0;
    double fs = 0;
    std::string mode;
    bool has_signal = false, has_window = false, has_hop = false, has_fs = false, has_mode = false;
};

// Minimal JSON reader for the flat object above (numbers, strings, arrays of numbers).
struct Parser {
    const std::string &s;
    size_t i = 0;
    std::string err;
    explicit Parser(const std::string &src) : s(src) {}
    void ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
    bool eat(char c) { ws(); if (i < s.size() && s[i] == c) { ++i; return true; } return false; }
    bool fail(const std::string &m) { if (err.empty()) err = m + " at byte " + std::to_string(i); return false; }
    bool str(std::string &out) {
        ws();
        if (i >= s.size() || s[i] != '"') return fail("expected string");
        ++i;
        out.clear();
        while (i < s.size() && s[i] != '"') {
            if (s[i] == '\\' && i + 1 < s.size()) { ++i; }
            out.push_back(s[i++]);
        }
        if (i >= s.size()) return fail("unterminated string");
        ++i;
        return true;
    }
    bool num(double &v) {
        ws();
        const char *b = s.c_str() + i;
        char *e = nullptr;
        v = std::strtod(b, &e);
        if (e == b) return fail("expected number");
        i += (size_t)(e - b);
        return true;
    }
    bool arr(std::vector<double> &v) {
        if (!eat('[')) return fail("expected '['");
        v.clear();
        if (eat(']')) return true;
        do {
            double x;
            if (!num(x)) return false;
            v.push_back(x);
        } while (eat(','));
        return eat(']') ? true : fail("expected ']'");
    }
    bool skip() {                                   // any value
        ws();
        if (i >= s.size()) return fail("unexpected end");
        if (s[i] == '"') { std::string t; return str(t); }
        if (s[i] == '[' || s[i] == '{') {
            const char open = s[i], close = open == '[' ? ']' : '}';
            int depth = 0;
            bool in_str = false;
            for (; i < s.size(); ++i) {
                if (in_str) { if (s[i] == '\\') ++i; else if (s[i] == '"') in_str = false; continue; }
                if (s[i] == '"') in_str = true;
                else if (s[i] == open) ++depth;
                else if (s[i] == close && --depth == 0) { ++i; return true; }
            }
            return fail("unterminated value");
        }
        while (i < s.size() && s[i] != ',' && s[i] != '}' && s[i] != ']') ++i;
        return true;
    }
    bool object(Input &in) {
        if (!eat('{')) return fail("expected '{'");
        if (eat('}')) return true;
        do {
            std::string key;
            if (!str(key)) return false;
            if (!eat(':')) return fail("expected ':'");
            if (key == "signal") { if (!arr(in.signal)) return false; in.has_signal = true; }
            else if (key == "window") { if (!arr(in.window)) return false; in.has_window = true; }
            else if (key == "hop_length") { double v; if (!num(v)) return false; in.hop = (long long)v; in.has_hop = v >= 0 && v == (double)(long long)v; if (!in.has_hop) return fail("invalid value for `hop_length`"); }
            else if (key == "fs") { if (!num(in.fs)) return false; in.has_fs = true; }
            else if (key == "fft_mode") { if (!str(in.mode)) return false; in.has_mode = true; }
            else if (!skip()) return false;
        } while (eat(','));
        return eat('}') ? true : fail("expected '}'");
    }
};

int die(const std::string &m) {
    std::fprintf(stderr, "Error: %s\n", m.c_str());
    return 1;
}

void put_f64(std::string &o, double v) {
    char b[40];
    std::snprintf(b, sizeof b, "%.17g", v);
    // serde_json prints finite floats only; keep JSON valid for non-finite values too
    if (std::strstr(b, "inf") || std::strstr(b, "nan")) { o += "null"; return; }
    o += b;
    if (!std::strpbrk(b, ".eE")) o += ".0";
}

}  // namespace

int main(int argc, char **argv) {
    if (argc != 2) {
        std::fprintf(stderr, "Usage: %s <input_json_file>\n", argv[0]);
        return 1;
    }
    std::ifstream f(argv[1]);
    if (!f) return die(std::string("No such file or directory: ") + argv[1]);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string text = ss.str();
    Input in;
    Parser ps(text);
    if (!ps.object(in)) return die(ps.err);
    for (const char *k : {"signal", "window", "hop_length", "fs", "fft_mode"}) {
        const bool ok = !std::strcmp(k, "signal") ? in.has_signal : !std::strcmp(k, "window") ? in.has_window
                      : !std::strcmp(k, "hop_length") ? in.has_hop : !std::strcmp(k, "fs") ? in.has_fs : in.has_mode;
        if (!ok) return die(std::string("missing field `") + k + "`");
    }

    veils_plan_desc d;
    std::memset(&d, 0, sizeof d);
    d.win = in.window.data();
    d.m_num = (int64_t)in.window.size();
    d.hop = in.hop;
    d.fs = in.fs;
    d.fft_mode = in.mode.c_str();
    d.precision = VEILS_F64;
    veils_plan *pl = nullptr;
    if (veils_plan_create(&d, &pl) != VEILS_OK) return die(veils_last_error());
    const int64_t n = (int64_t)in.signal.size();
    int64_t P = 0, p_max = 0, k0 = 0, k1 = 0;
    if (veils_p_num(pl, n, &P) != VEILS_OK) return die(veils_last_error());
    if (veils_p_max(pl, n, &p_max) != VEILS_OK) return die(veils_last_error());
    const int64_t F = veils_f_pts(pl);
    std::vector<double> S((size_t)(2 * F * P));
    if (veils_stft_host(pl, in.signal.data(), n, 1, n, nullptr, nullptr, 0, S.data()) != VEILS_OK)
        return die(veils_last_error());
    if (veils_istft_range(pl, P, nullptr, nullptr, &k0, &k1) != VEILS_OK) return die(veils_last_error());
    std::vector<double> y((size_t)(k1 - k0));
    if (veils_istft_host(pl, S.data(), P, 1, nullptr, nullptr, y.data()) != VEILS_OK) return die(veils_last_error());

    std::string o;
    o.reserve((size_t)(F * P) * 64 + y.size() * 24 + 256);
    o += "{\"stft\":[";
    for (int64_t fi = 0; fi < F; ++fi) {
        o += fi ? ",[" : "[";
        for (int64_t p = 0; p < P; ++p) {
            o += p ? ",{\"real\":" : "{\"real\":";
            put_f64(o, S[(size_t)(2 * (fi * P + p))]);
            o += ",\"imag\":";
            put_f64(o, S[(size_t)(2 * (fi * P + p) + 1)]);
            o += "}";
        }
        o += "]";
    }
    o += "],\"istft\":[";
    for (size_t i = 0; i < y.size(); ++i) {
        if (i) o += ",";
        put_f64(o, y[i]);
    }
    o += "],\"properties\":{\"m_num\":" + std::to_string(veils_m_num(pl)) + ",\"f_pts\":" + std::to_string(F) +
         ",\"p_min\":" + std::to_string(veils_p_min(pl)) + ",\"p_max\":" + std::to_string(p_max) +
         ",\"mfft\":" + std::to_string(veils_mfft(pl)) + ",\"hop\":" + std::to_string(veils_hop(pl)) + ",\"fs\":";
    put_f64(o, veils_fs(pl));
    o += "}}\n";
    std::fwrite(o.data(), 1, o.size(), stdout);
    veils_plan_destroy(pl);
    return 0;
}
