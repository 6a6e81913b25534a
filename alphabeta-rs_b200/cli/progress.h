// progress.h — the reference's progress bars (src/progress.rs:5-43, indicatif) for the two command-line tools.
//
//   multi(total)      "Progress  {bar:40} [{elapsed}] {pos:>7}/{len:7} ETA: {eta}"        src/progress.rs:24-43
//   specific(bars, n) "ABNeutral {bar:40} [{elapsed}] {pos:>7}/{len:7}" + "BootModel ..." src/progress.rs:5-23
//
// Like indicatif the bars go to stderr and are drawn only when stderr is a terminal (ABFIT_PROGRESS=1 forces them,
// =0 hides them).  The GPU fits a whole batch in one call, so a bar advances per finished stage or batch, not per
// start: it tells which stage runs and how long it has been running, which is what the reference's bars are for.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <unistd.h>

namespace progress {

inline bool enabled()
{
    static const int on = [] {
        if (const char *e = std::getenv("ABFIT_PROGRESS")) return std::atoi(e) != 0 ? 1 : 0;
        return isatty(2) ? 1 : 0;
    }();
    return on != 0;
}

inline std::string hms(double seconds)  // indicatif's {elapsed}: "3s", "2m", "1h" style is HumanDuration; {elapsed} is "MM:SS"
{
    const long s = (long)(seconds + 0.5);
    char buf[32];
    if (s >= 3600) std::snprintf(buf, sizeof buf, "%02ld:%02ld:%02ld", s / 3600, s / 60 % 60, s % 60);
    else std::snprintf(buf, sizeof buf, "%02ld:%02ld", s / 60, s % 60);
    return buf;
}

inline std::string human(double seconds)  // HumanDuration of the ETA (src/progress.rs:33-40)
{
    const long s = (long)(seconds + 0.5);
    char buf[48];
    if (s < 60) std::snprintf(buf, sizeof buf, "%ld second%s", s, s == 1 ? "" : "s");
    else if (s < 3600) std::snprintf(buf, sizeof buf, "%ld minute%s", s / 60, s / 60 == 1 ? "" : "s");
    else std::snprintf(buf, sizeof buf, "%ld hour%s", s / 3600, s / 3600 == 1 ? "" : "s");
    return buf;
}

class Bar {
public:
    Bar(std::string msg, unsigned long long len, bool with_eta) : msg_(std::move(msg)), len_(len), eta_(with_eta), t0_(clock::now()) {}
    void inc(unsigned long long d = 1) { set(pos_ + d); }
    void set(unsigned long long pos)
    {
        pos_ = pos > len_ ? len_ : pos;
        draw(false);
    }
    void tick() { draw(false); }
    void finish()
    {
        pos_ = len_;
        draw(true);
    }
    // the line as it would be drawn (tests)
    std::string render() const
    {
        const double el = std::chrono::duration<double>(clock::now() - t0_).count();
        const int width = 40;
        const int full = len_ ? (int)((double)pos_ / (double)len_ * width) : width;
        std::string bar;
        for (int i = 0; i < width; ++i) bar += i < full ? "\xe2\x96\x88" : (i == full ? "\xe2\x96\x91" : " ");
        char tail[160];
        std::snprintf(tail, sizeof tail, " [%s] %7llu/%-7llu", hms(el).c_str(), pos_, len_);
        std::string line = msg_ + " " + bar + tail;
        if (eta_) {
            const double eta = pos_ ? el / (double)pos_ * (double)(len_ - pos_) : 0.0;
            line += " ETA: " + human(eta);
        }
        return line;
    }

private:
    using clock = std::chrono::steady_clock;
    void draw(bool last)
    {
        if (!enabled()) return;
        std::fprintf(stderr, "\r\033[2K%s%s", render().c_str(), last ? "\n" : "");
        std::fflush(stderr);
    }
    std::string msg_;
    unsigned long long len_, pos_ = 0;
    bool eta_;
    clock::time_point t0_;
};

}  // namespace progress
