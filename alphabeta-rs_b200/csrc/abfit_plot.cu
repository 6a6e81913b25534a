// abfit_plot.cu — the two pictures of the reference (src/plot.rs), drawn on the host.
//
//   abfit_plot_metaplot   src/plot.rs:6-82    "Metaplot": alpha (red) and beta (blue) per window over x = 0..300,
//                                             y = 0..0.01, their 95 % confidence bands as translucent polygons, legend
//   abfit_plot_bootstrap  src/plot.rs:84-137  "Bootstrap Boxplot": box-and-whisker of the bootstrap alphas and betas,
//                                             y = 0..1.3 max, axis label "Epimutation rate"
//
// Same canvas (1280 x 960), same series, colours, ranges and captions as the reference; not the same pixels — the
// reference rasterises with the `plotters` crate and a system font, this file with its own few primitives and a
// bitmap font (csrc/abfit_font.inc, tools/make_font.py).  PNG encoder: 8-bit RGB, filter 0, deflate with fixed Huffman
// codes and run-length matches at distance 3 (one pixel) — no zlib dependency, flat areas shrink by two orders of
// magnitude.
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/abfit.h"
#include "abfit_font.inc"
#include "abfit_internal.h"

namespace abfit {
namespace {

struct RGB {
    unsigned char r, g, b;
};
const RGB WHITE{255, 255, 255}, BLACK{0, 0, 0}, RED{255, 0, 0}, BLUE{0, 0, 255}, GRID{215, 215, 215}, GRID_LIGHT{238, 238, 238};

struct Canvas {
    int w, h;
    std::vector<unsigned char> px;
    Canvas(int w_, int h_) : w(w_), h(h_), px((size_t)w_ * h_ * 3, 255) {}
    void blend(int x, int y, RGB c, double a = 1.0)
    {
        if (x < 0 || y < 0 || x >= w || y >= h) return;
        unsigned char *p = &px[((size_t)y * w + x) * 3];
        p[0] = (unsigned char)lround(p[0] * (1.0 - a) + c.r * a);
        p[1] = (unsigned char)lround(p[1] * (1.0 - a) + c.g * a);
        p[2] = (unsigned char)lround(p[2] * (1.0 - a) + c.b * a);
    }
    void rect(int x0, int y0, int x1, int y1, RGB c, double a = 1.0)  // filled, inclusive
    {
        for (int y = std::max(y0, 0); y <= std::min(y1, h - 1); ++y)
            for (int x = std::max(x0, 0); x <= std::min(x1, w - 1); ++x) blend(x, y, c, a);
    }
    void frame(int x0, int y0, int x1, int y1, RGB c)
    {
        rect(x0, y0, x1, y0, c);
        rect(x0, y1, x1, y1, c);
        rect(x0, y0, x0, y1, c);
        rect(x1, y0, x1, y1, c);
    }
    void line(double xa, double ya, double xb, double yb, RGB c, int width = 1)
    {
        const double dx = xb - xa, dy = yb - ya;
        const int n = (int)std::max(fabs(dx), fabs(dy)) + 1;
        for (int i = 0; i <= n; ++i) {
            const double t = n ? (double)i / n : 0.0;
            const int x = (int)lround(xa + t * dx), y = (int)lround(ya + t * dy);
            for (int oy = -(width / 2); oy <= (width - 1) / 2; ++oy)
                for (int ox = -(width / 2); ox <= (width - 1) / 2; ++ox) blend(x + ox, y + oy, c);
        }
    }
    // even-odd scanline fill of a closed polygon
    void polygon(const std::vector<std::pair<double, double>> &pts, RGB c, double a)
    {
        if (pts.size() < 3) return;
        double ymin = pts[0].second, ymax = pts[0].second;
        for (auto &p : pts) {
            ymin = std::min(ymin, p.second);
            ymax = std::max(ymax, p.second);
        }
        std::vector<double> xs;
        for (int y = std::max(0, (int)floor(ymin)); y <= std::min(h - 1, (int)ceil(ymax)); ++y) {
            const double yc = y + 0.5;
            xs.clear();
            for (size_t i = 0; i < pts.size(); ++i) {
                const auto &p = pts[i], &q = pts[(i + 1) % pts.size()];
                if ((p.second <= yc) != (q.second <= yc)) xs.push_back(p.first + (yc - p.second) / (q.second - p.second) * (q.first - p.first));
            }
            std::sort(xs.begin(), xs.end());
            for (size_t i = 0; i + 1 < xs.size(); i += 2)
                for (int x = std::max(0, (int)ceil(xs[i] - 0.5)); x <= std::min(w - 1, (int)floor(xs[i + 1] - 0.5)); ++x) blend(x, y, c, a);
        }
    }
    static int text_width(const std::string &s, int scale)
    {
        int wpx = 0;
        for (unsigned char ch : s) wpx += (ch >= 32 && ch < 127 ? FONT_ADV[ch - 32] : FONT_ADV[0]) * scale;
        return wpx;
    }
    // vertical = true: rotated by 90 degrees, reading upwards (y-axis description)
    void text(int x, int y, const std::string &s, RGB c, int scale = 1, bool vertical = false)
    {
        int pen = 0;
        for (unsigned char ch : s) {
            const int g = ch >= 32 && ch < 127 ? ch - 32 : 0;
            for (int gy = 0; gy < FONT_H; ++gy)
                for (int gx = 0; gx < 16; ++gx)
                    if (FONT_ROWS[g][gy] >> gx & 1)
                        for (int sy = 0; sy < scale; ++sy)
                            for (int sx = 0; sx < scale; ++sx) {
                                const int u = pen + gx * scale + sx, v = gy * scale + sy;
                                if (vertical) blend(x + v, y - u, c);
                                else blend(x + u, y + v, c);
                            }
            pen += FONT_ADV[g] * scale;
        }
    }
};

// ---- PNG ------------------------------------------------------------------------------------------
struct BitWriter {
    std::vector<unsigned char> out;
    uint32_t acc = 0;
    int n = 0;
    void bits(uint32_t v, int k)  // LSB first
    {
        acc |= v << n;
        n += k;
        while (n >= 8) {
            out.push_back((unsigned char)(acc & 0xff));
            acc >>= 8;
            n -= 8;
        }
    }
    void huff(uint32_t code, int k)  // Huffman codes go MSB first
    {
        uint32_t r = 0;
        for (int i = 0; i < k; ++i) r |= ((code >> i) & 1u) << (k - 1 - i);
        bits(r, k);
    }
    void flush()
    {
        if (n > 0) bits(0, 8 - n);
    }
};

void put_symbol(BitWriter &bw, int sym)  // fixed literal/length code (RFC 1951, 3.2.6)
{
    if (sym <= 143) bw.huff(0x30 + sym, 8);
    else if (sym <= 255) bw.huff(0x190 + (sym - 144), 9);
    else if (sym <= 279) bw.huff(sym - 256, 7);
    else bw.huff(0xc0 + (sym - 280), 8);
}

void put_length(BitWriter &bw, int len)  // 3 .. 258
{
    static const int base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const int extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    int k = 28;
    while (base[k] > len) --k;
    put_symbol(bw, 257 + k);
    if (extra[k]) bw.bits((uint32_t)(len - base[k]), extra[k]);
}

std::vector<unsigned char> deflate_rle3(const std::vector<unsigned char> &d)
{
    BitWriter bw;
    bw.bits(1, 1);  // final block
    bw.bits(1, 2);  // fixed Huffman codes
    size_t i = 0;
    while (i < d.size()) {
        size_t len = 0;
        if (i >= 3)
            while (len < 258 && i + len < d.size() && d[i + len] == d[i + len - 3]) ++len;
        if (len >= 3) {
            put_length(bw, (int)len);
            bw.huff(2, 5);  // distance code 2 = distance 3, no extra bits
            i += len;
        } else {
            put_symbol(bw, d[i]);
            ++i;
        }
    }
    put_symbol(bw, 256);
    bw.flush();
    return bw.out;
}

uint32_t crc32_of(const unsigned char *p, size_t n, uint32_t crc = 0)
{
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = c & 1 ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

void be32(std::vector<unsigned char> &v, uint32_t x)
{
    for (int s = 24; s >= 0; s -= 8) v.push_back((unsigned char)(x >> s));
}

void chunk(std::vector<unsigned char> &png, const char type[4], const std::vector<unsigned char> &data)
{
    be32(png, (uint32_t)data.size());
    const size_t at = png.size();
    png.insert(png.end(), type, type + 4);
    png.insert(png.end(), data.begin(), data.end());
    be32(png, crc32_of(&png[at], png.size() - at));
}

int write_png(const char *path, const Canvas &cv)
{
    std::vector<unsigned char> raw;
    raw.reserve((size_t)cv.h * (cv.w * 3 + 1));
    for (int y = 0; y < cv.h; ++y) {
        raw.push_back(0);  // filter: none
        raw.insert(raw.end(), cv.px.begin() + (size_t)y * cv.w * 3, cv.px.begin() + (size_t)(y + 1) * cv.w * 3);
    }
    uint32_t s1 = 1, s2 = 0;  // Adler-32 of the uncompressed stream
    for (unsigned char c : raw) {
        s1 = (s1 + c) % 65521u;
        s2 = (s2 + s1) % 65521u;
    }
    std::vector<unsigned char> z{0x78, 0x01};
    const std::vector<unsigned char> body = deflate_rle3(raw);
    z.insert(z.end(), body.begin(), body.end());
    be32(z, (s2 << 16) | s1);
    std::vector<unsigned char> png{0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
    std::vector<unsigned char> ihdr;
    be32(ihdr, (uint32_t)cv.w);
    be32(ihdr, (uint32_t)cv.h);
    const unsigned char tail[5] = {8, 2, 0, 0, 0};  // 8 bits, RGB, deflate, adaptive filtering, no interlace
    ihdr.insert(ihdr.end(), tail, tail + 5);
    chunk(png, "IHDR", ihdr);
    chunk(png, "IDAT", z);
    chunk(png, "IEND", {});
    FILE *f = fopen(path, "wb");
    if (!f) {
        set_error(std::string("cannot write ") + path);
        return ABFIT_ERR_ARG;
    }
    const bool ok = fwrite(png.data(), 1, png.size(), f) == png.size();
    if (fclose(f) != 0 || !ok) {
        set_error(std::string("short write to ") + path);
        return ABFIT_ERR_ARG;
    }
    return 0;
}

// ---- chart frame shared by the two pictures ----------------------------------------------------------
struct Chart {
    Canvas &cv;
    int x0, y0, x1, y1;  // plotting area in pixels (y0 = top)
    double xa, xb, ya, yb;
    double px(double x) const { return x0 + (x - xa) / (xb - xa) * (x1 - x0); }
    double py(double y) const { return y1 - (y - ya) / (yb - ya) * (y1 - y0); }
};

std::string tick_label(double v, double span)
{
    char buf[64];
    if (v == 0.0) return "0";
    if (span >= 10.0) snprintf(buf, sizeof buf, "%.0f", v);
    else if (span >= 0.1) snprintf(buf, sizeof buf, "%.2f", v);
    else if (span >= 1e-4) snprintf(buf, sizeof buf, "%.*f", (int)ceil(-log10(span)) + 2, v);
    else snprintf(buf, sizeof buf, "%.2e", v);
    return buf;
}

void y_mesh(Chart &c, int n_ticks, bool light)
{
    for (int i = 0; i <= n_ticks; ++i) {
        const double v = c.ya + (c.yb - c.ya) * i / n_ticks;
        const int y = (int)lround(c.py(v));
        if (i > 0 && i < n_ticks) c.cv.rect(c.x0, y, c.x1, y, light ? GRID_LIGHT : GRID);
        c.cv.rect(c.x0 - 5, y, c.x0, y, BLACK);
        const std::string t = tick_label(v, c.yb - c.ya);
        c.cv.text(c.x0 - 8 - Canvas::text_width(t, 1), y - FONT_H / 2, t, BLACK);
    }
}

bool finite_all(const double *a, int n)
{
    for (int i = 0; i < n; ++i)
        if (!std::isfinite(a[i])) return false;
    return true;
}

// plotters' Quartiles: linear-interpolated percentiles 25 / 50 / 75, whiskers at the fences q1 - 1.5 iqr, q3 + 1.5 iqr
void quartiles(std::vector<double> v, double out[5])
{
    std::sort(v.begin(), v.end());
    auto pct = [&](double p) {
        if (v.size() == 1) return v[0];
        const double pos = p / 100.0 * (double)(v.size() - 1);
        const size_t lo = (size_t)floor(pos);
        const size_t hi = std::min(lo + 1, v.size() - 1);
        return v[lo] + (v[hi] - v[lo]) * (pos - (double)lo);
    };
    const double q1 = pct(25), q2 = pct(50), q3 = pct(75), iqr = q3 - q1;
    out[0] = q1 - 1.5 * iqr;
    out[1] = q1;
    out[2] = q2;
    out[3] = q3;
    out[4] = q3 + 1.5 * iqr;
}

}  // namespace
}  // namespace abfit

using namespace abfit;

extern "C" int abfit_plot_metaplot(const char *path, int32_t n_windows, const double *alpha, const double *beta,
                                   const double *ci_alpha_lo, const double *ci_alpha_hi, const double *ci_beta_lo,
                                   const double *ci_beta_hi)
{
    if (!path || n_windows < 0 || (n_windows > 0 && (!alpha || !beta))) {
        set_error("abfit_plot_metaplot: bad arguments");
        return ABFIT_ERR_ARG;
    }
    Canvas cv(1280, 960);  // BitMapBackend::new(.., (640 * 2, 480 * 2))
    const std::string title = "Metaplot";
    cv.text((cv.w - Canvas::text_width(title, 3)) / 2, 8, title, BLACK, 3);
    Chart c{cv, 5 + 30 + 40, 5 + 3 * FONT_H + 10, cv.w - 5 - 10, cv.h - 5 - 30, 0.0, 300.0, 0.0, 0.01};  // build_cartesian_2d(0..300, 0..0.01)
    for (int i = 0; i <= 10; ++i) {  // mesh
        const double v = 30.0 * i;
        const int x = (int)lround(c.px(v));
        if (i > 0 && i < 10) cv.rect(x, c.y0, x, c.y1, GRID);
        cv.rect(x, c.y1, x, c.y1 + 5, BLACK);
        const std::string t = tick_label(v, 300.0);
        cv.text(x - Canvas::text_width(t, 1) / 2, c.y1 + 8, t, BLACK);
    }
    y_mesh(c, 10, false);
    // confidence bands first, lines on top.  The reference lists (i, hi), (i, lo) for every window in turn and fills
    // that zig-zag; the band between the two envelopes is what it shows, and what is drawn here.
    auto band = [&](const double *lo, const double *hi, RGB col) {
        if (!lo || !hi || n_windows < 2) return;
        std::vector<std::pair<double, double>> pts;
        auto clampy = [&](double v) { return std::min(std::max(v, c.ya), c.yb); };
        for (int i = 0; i < n_windows; ++i)
            if (std::isfinite(hi[i])) pts.push_back({c.px(i), c.py(clampy(hi[i]))});
        for (int i = n_windows - 1; i >= 0; --i)
            if (std::isfinite(lo[i])) pts.push_back({c.px(i), c.py(clampy(lo[i]))});
        cv.polygon(pts, col, 0.2);
    };
    band(ci_alpha_lo, ci_alpha_hi, RED);
    band(ci_beta_lo, ci_beta_hi, BLUE);
    auto series = [&](const double *v, RGB col) {
        for (int i = 0; i + 1 < n_windows; ++i) {
            if (!std::isfinite(v[i]) || !std::isfinite(v[i + 1])) continue;
            // clip to the plotting area by sampling: segments that leave the y range are cut at the frame
            const double xa = c.px(i), xb = c.px(i + 1), ya = c.py(v[i]), yb = c.py(v[i + 1]);
            const int n = (int)std::max(fabs(xb - xa), fabs(yb - ya)) + 1;
            for (int k = 0; k <= n; ++k) {
                const double t = (double)k / n, x = xa + t * (xb - xa), y = ya + t * (yb - ya);
                if (x >= c.x0 && x <= c.x1 && y >= c.y0 && y <= c.y1) cv.blend((int)lround(x), (int)lround(y), col);
            }
        }
    };
    series(alpha, RED);
    series(beta, BLUE);
    cv.frame(c.x0, c.y0, c.x1, c.y1, BLACK);
    {  // series labels: white box (80 %), black border, upper right (plotters' default position is MiddleRight)
        const int bw = 110, bh = 2 * FONT_H + 14, bx = c.x1 - bw - 12, by = (c.y0 + c.y1) / 2 - bh / 2;
        cv.rect(bx, by, bx + bw, by + bh, WHITE, 0.8);
        cv.frame(bx, by, bx + bw, by + bh, BLACK);
        cv.line(bx + 8, by + 7 + FONT_H / 2, bx + 28, by + 7 + FONT_H / 2, RED);
        cv.text(bx + 36, by + 5, "Alpha", BLACK);
        cv.line(bx + 8, by + 7 + FONT_H + FONT_H / 2, bx + 28, by + 7 + FONT_H + FONT_H / 2, BLUE);
        cv.text(bx + 36, by + 5 + FONT_H, "Beta", BLACK);
    }
    return write_png(path, cv);
}

extern "C" int abfit_plot_bootstrap(const char *path, const double *alphas, const double *betas, int32_t n)
{
    if (!path || n <= 0 || !alphas || !betas) {
        set_error("abfit_plot_bootstrap: bad arguments");
        return ABFIT_ERR_ARG;
    }
    if (!finite_all(alphas, n) || !finite_all(betas, n)) {
        set_error("abfit_plot_bootstrap: non-finite bootstrap estimates");
        return ABFIT_ERR_ARG;
    }
    double vmax = 0.0;  // fold(0.0, max) over both columns (src/plot.rs:87-91)
    for (int i = 0; i < n; ++i) vmax = std::max(vmax, std::max(alphas[i], betas[i]));
    if (!(vmax > 0.0)) vmax = 1.0;
    Canvas cv(1280, 960);
    const std::string title = "Bootstrap Boxplot";
    cv.text((cv.w - Canvas::text_width(title, 2)) / 2, 24, title, BLACK, 2);
    Chart c{cv, 20 + 90 + 30, 20 + 2 * FONT_H + 20, cv.w - 20 - 10, cv.h - 20 - 30, 0.0, 2.0, 0.0, vmax * 1.3};
    y_mesh(c, 10, true);
    cv.text(24, (c.y0 + c.y1) / 2 + Canvas::text_width("Epimutation rate", 2) / 2, "Epimutation rate", BLACK, 2, true);
    const char *names[2] = {"Alpha", "Beta"};
    const double *cols[2] = {alphas, betas};
    const RGB colour[2] = {RED, BLUE};
    for (int k = 0; k < 2; ++k) {
        const int xc = (int)lround(c.px(0.5 + k));
        cv.rect(xc, c.y1, xc, c.y1 + 5, BLACK);
        cv.text(xc - Canvas::text_width(names[k], 1) / 2, c.y1 + 8, names[k], BLACK);
        double q[5];
        quartiles(std::vector<double>(cols[k], cols[k] + n), q);
        auto yy = [&](double v) { return (int)lround(c.py(std::min(std::max(v, c.ya), c.yb))); };
        const int half = 50;  // .width(100)
        cv.line(xc, yy(q[0]), xc, yy(q[1]), colour[k], 2);  // whiskers
        cv.line(xc, yy(q[3]), xc, yy(q[4]), colour[k], 2);
        cv.line(xc - half, yy(q[0]), xc + half, yy(q[0]), colour[k], 2);  // .whisker_width(1.0)
        cv.line(xc - half, yy(q[4]), xc + half, yy(q[4]), colour[k], 2);
        for (int t = 0; t < 2; ++t) cv.frame(xc - half + t, yy(q[3]) + t, xc + half - t, yy(q[1]) - t, colour[k]);  // box
        cv.line(xc - half, yy(q[2]), xc + half, yy(q[2]), colour[k], 2);  // median
    }
    cv.frame(c.x0, c.y0, c.x1, c.y1, BLACK);
    return write_png(path, cv);
}
