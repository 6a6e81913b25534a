// abfit_io.cu — result writers of the reference, byte for byte (host only).
//
//   f64 Display of Rust (`{}`)           shortest round-trip digits, never scientific, "1" not "1.0", NaN / inf
//   Pedigree::to_file                    src/pedigree.rs:81-90
//   Analysis::to_file / Display          src/analysis.rs:102-187
//   write_npy(raw.npy)                   src/cli/alphabeta.rs:34-35, src/cli/metaprofile.rs:110-111 (NPY v1.0, '<f8', C order)
//   metaprofile results.txt              src/cli/metaprofile.rs:74-99
//   steady_state                         src/alphabeta.rs:71-79
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "abfit_internal.h"

namespace {

// Rust's `impl Display for f64` without a precision: flt2dec shortest digits, positional notation.
std::string rust_f64(double v)
{
    if (v != v) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    std::string out;
    if (std::signbit(v)) {
        out += '-';
        v = -v;
    }
    if (v == 0.0) return out + "0";
    char buf[64];
    // shortest round-trip digits in scientific form: d[.ddd]e[+-]xx
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);
    std::string s(buf, r.ptr);
    const size_t epos = s.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; ++i)
        if (s[i] != '.') digits += s[i];
    const int exp10 = std::atoi(s.c_str() + epos + 1);  // value = d.ddd * 10^exp10
    const int point = exp10 + 1;                        // digits before the decimal point
    const int n = (int)digits.size();
    if (point <= 0) {
        out += "0.";
        out.append((size_t)(-point), '0');
        out += digits;
    } else if (point < n) {
        out += digits.substr(0, (size_t)point);
        out += '.';
        out += digits.substr((size_t)point);
    } else {
        out += digits;
        out.append((size_t)(point - n), '0');
    }
    return out;
}

int write_file(const char *path, const std::string &content)
{
    FILE *f = std::fopen(path, "wb");
    if (!f) {
        abfit::set_error(std::string("cannot create ") + path);
        return ABFIT_ERR_ARG;
    }
    const size_t w = std::fwrite(content.data(), 1, content.size(), f);
    std::fclose(f);
    if (w != content.size()) {
        abfit::set_error(std::string("short write to ") + path);
        return ABFIT_ERR_ARG;
    }
    return 0;
}

// src/alphabeta.rs:62-79
double p_mm_est_h(double a, double b)
{
    return (a * ((1.0 - a) * (1.0 - a) - (1.0 - b) * (1.0 - b) - 1.0)) / ((a + b) * ((a + b - 1.0) * (a + b - 1.0) - 2.0));
}

std::string analysis_text(const double a[32])
{
    // abfit_analyze layout: 8 means, 8 sds, 8 (lo, hi) pairs; field order alpha, beta, beta/alpha, weight, intercept, mm, um, uu
    static const char *names[8] = {"Alpha", "Beta", "AlphaBeta", "Weight", "Intercept", "PrMM", "PrUM", "PrUU"};
    std::string s;
    for (int i = 0; i < 8; ++i) s += std::string(names[i]) + "\t" + rust_f64(a[i]) + "\n";
    for (int i = 0; i < 8; ++i) s += std::string("SD") + names[i] + "\t" + rust_f64(a[8 + i]) + "\n";
    for (int i = 0; i < 8; ++i)
        s += std::string("CI") + names[i] + "\t" + rust_f64(a[16 + 2 * i]) + "-" + rust_f64(a[17 + 2 * i]) + "\n";
    return s;
}

}  // namespace

extern "C" {

int abfit_format_f64(double v, char *buf, int32_t cap)
{
    const std::string s = rust_f64(v);
    if (!buf || cap <= (int32_t)s.size()) return ABFIT_ERR_ARG;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

double abfit_steady_state(double alpha, double beta)
{
    const double pi_2 = p_mm_est_h(alpha, beta);  // src/alphabeta.rs:71-79
    const double pi_1 = (4.0 * alpha * beta * (alpha + beta - 2.0)) / ((alpha + beta) * ((alpha + beta - 1.0) * (alpha + beta - 1.0) - 2.0));
    return pi_2 + 0.5 * pi_1;
}

int abfit_write_pedigree(const char *path, const double *pedigree, int32_t n_pairs)
{
    if (!path || (!pedigree && n_pairs > 0) || n_pairs < 0) return ABFIT_ERR_ARG;
    std::string c = "time0\ttime1\ttime2\tD.value\n";
    for (int32_t i = 0; i < n_pairs; ++i) {
        const double *r = pedigree + 4 * (size_t)i;
        c += rust_f64(r[0]) + "\t" + rust_f64(r[1]) + "\t" + rust_f64(r[2]) + "\t" + rust_f64(r[3]) + "\n";
    }
    return write_file(path, c);
}

int abfit_write_analysis(const char *path, const double analysis[32])
{
    if (!path || !analysis) return ABFIT_ERR_ARG;
    return write_file(path, analysis_text(analysis));
}

int abfit_format_analysis(const double analysis[32], char *buf, int32_t cap)
{
    if (!analysis) return ABFIT_ERR_ARG;
    const std::string s = analysis_text(analysis);
    if (!buf || cap <= (int32_t)s.size()) return ABFIT_ERR_ARG;
    std::memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

int abfit_write_npy_f64(const char *path, const double *data, int32_t ndim, const int64_t *shape)
{
    if (!path || ndim < 0 || ndim > 8 || (ndim && !shape)) return ABFIT_ERR_ARG;
    size_t n = 1;
    std::string sh = "(";
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 0) return ABFIT_ERR_ARG;
        n *= (size_t)shape[i];
        sh += std::to_string(shape[i]) + (ndim == 1 || i + 1 < ndim ? "," : "") + (i + 1 < ndim ? " " : "");
    }
    sh += ")";
    std::string dict = "{'descr': '<f8', 'fortran_order': False, 'shape': " + sh + ", }";
    // NPY 1.0: magic(6) + version(2) + header length u16 LE + header, the whole preamble padded with spaces to a
    // multiple of 64 bytes and terminated by '\n'
    size_t total = 10 + dict.size() + 1;
    const size_t pad = (64 - total % 64) % 64;
    dict.append(pad, ' ');
    dict += '\n';
    std::string c("\x93NUMPY\x01\x00", 8);
    const uint16_t hl = (uint16_t)dict.size();
    c += (char)(hl & 0xff);
    c += (char)(hl >> 8);
    c += dict;
    if (n && !data) return ABFIT_ERR_ARG;
    c.append(reinterpret_cast<const char *>(data), n * 8);
    return write_file(path, c);
}

int abfit_write_metaprofile_results(const char *path, const char *run_name, int32_t n_windows, const int32_t *cg_count,
                                    const int32_t *region, const abfit_fit *best, const double *analysis,
                                    const double *obs_steady_state)
{
    if (!path || !run_name || n_windows < 0 || (n_windows && (!cg_count || !region || !best || !analysis || !obs_steady_state)))
        return ABFIT_ERR_ARG;
    static const char *reg[3] = {"upstream", "gene", "downstream"};
    std::string c =
        "run;window;cg_count;region;alpha;beta;1/2*(alpha+beta);pred_steady_state;obs_steady_state;sd_alpha;sd_beta;"
        "ci_alpha_0.025;ci_alpha_0.975;ci_beta_0.025;ci_beta_0.975\n";
    for (int32_t i = 0; i < n_windows; ++i) {
        if (region[i] < 0 || region[i] > 2) return ABFIT_ERR_ARG;
        const double a = best[i].theta[0], b = best[i].theta[1];
        const double *an = analysis + 32 * (size_t)i;
        c += std::string(run_name) + ";" + std::to_string(i) + ";" + std::to_string(cg_count[i]) + ";" + reg[region[i]] + ";" +
             rust_f64(a) + ";" + rust_f64(b) + ";" + rust_f64(0.5 * (a + b)) + ";" + rust_f64(abfit_steady_state(a, b)) + ";" +
             rust_f64(obs_steady_state[i]) + ";" + rust_f64(an[8]) + ";" + rust_f64(an[9]) + ";" + rust_f64(an[16]) + ";" +
             rust_f64(an[17]) + ";" + rust_f64(an[18]) + ";" + rust_f64(an[19]) + "\n";
    }
    return write_file(path, c);
}

}  // extern "C"
