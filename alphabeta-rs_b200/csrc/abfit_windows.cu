// abfit_windows.cu — CG site -> metaprofile window assignment (host only).
//
// Replaces, bit for bit:
//   MethylationSite::is_in_gene        src/methylation_site.rs:368-378
//   MethylationSite::find_gene         src/methylation_site.rs:385-418
//   MethylationSite::place_in_windows  src/methylation_site.rs:423-490
//   Windows::new                       src/windows.rs:28-44
//   the gene-caching loop of Windows::extract   src/windows.rs:331-337
//   Windows::distribution              src/windows.rs:158-165
//   Genome / GenesByStrand (per chromosome, three lists stably sorted by start)  src/genes.rs:24-57,127-163
//
// The loop over sites carries state (the previously matched gene is reused while the site is still
// inside it, which decides the gene when genes overlap), so this runs on the host in site order; it is
// integer and f64-compare logic, O(L log G).  Its output (window of every kept site) is what turns a
// methylome into the `seg_offsets` of abfit_divergence.
//
// u32 arithmetic wraps like the reference's release build (`site.start + cutoff`, `gene.end - gene.start`).
#include <algorithm>
#include <cmath>
#include <map>
#include <vector>

#include "abfit_internal.h"

namespace {

struct Lists {
    std::vector<abfit_gene> sense, antisense, combined;
};

// Strand equality of the reference: "if one of them is unknown, they are equal" (src/genes.rs:88-96)
inline bool strand_eq(int a, int b) { return !((a > 0 && b < 0) || (a < 0 && b > 0)); }

inline uint32_t gene_cutoff(const abfit_gene &g, const abfit_window_args &a)
{
    return a.cutoff_gene_length ? (uint32_t)(g.end - g.start) : a.cutoff;
}

inline bool is_in_gene(const abfit_cg_site &s, const abfit_gene &g, const abfit_window_args &a)
{
    const uint32_t cutoff = gene_cutoff(g, a);
    return s.chromosome == g.chromosome && g.start <= (uint32_t)(s.start + cutoff) && s.end <= (uint32_t)(g.end + cutoff) &&
           strand_eq(s.strand, g.strand);
}

// slice::binary_search_by_key of Rust 1.52 .. 1.81 (the toolchains contemporary with the reference's
// Cargo.lock).  The keys `gene.end + cutoff` are NOT sorted in general (the lists are sorted by start),
// so the probe sequence itself is part of the behaviour and is restated literally.
inline size_t rust_binary_search(const std::vector<abfit_gene> &v, uint32_t target, const abfit_window_args &a)
{
    size_t size = v.size(), left = 0, right = size;
    while (left < right) {
        const size_t mid = left + size / 2;
        const uint32_t key = (uint32_t)(v[mid].end + gene_cutoff(v[mid], a));
        if (key < target)
            left = mid + 1;
        else if (key > target)
            right = mid;
        else
            return mid;  // Ok(mid)
        size = right - left;
    }
    return left;  // Err(left), collapsed by unwrap_or_else(|x| x)
}

const abfit_gene *find_gene(const abfit_cg_site &s, const std::map<int32_t, Lists> &genome, const abfit_window_args &a)
{
    auto it = genome.find(s.chromosome);
    if (it == genome.end()) return nullptr;
    const std::vector<abfit_gene> &strand = s.strand > 0 ? it->second.sense : s.strand < 0 ? it->second.antisense : it->second.combined;
    const size_t idx = rust_binary_search(strand, s.start, a);
    if (strand.size() < idx + 1) return nullptr;
    const abfit_gene *g = &strand[idx];
    return is_in_gene(s, *g, a) ? g : nullptr;
}

struct Counts {
    uint32_t up_down, gene;
};

int window_counts(abfit_window_args &a, Counts &c)
{
    if (a.window_step == 0) a.window_step = a.window_size;  // src/extract.rs:26-28
    if (a.window_step == 0) {
        abfit::set_error("window_size and window_step are both 0");
        return ABFIT_ERR_ARG;
    }
    c.gene = a.absolute ? a.max_gene_length / a.window_step : 100u / a.window_step;  // src/windows.rs:29-33
    c.up_down = a.absolute ? a.cutoff / a.window_step : 100u / a.window_step;        // :34-38
    return 0;
}

}  // namespace

extern "C" {

int abfit_window_counts(const abfit_window_args *args, int32_t n_windows_out[3])
{
    if (!args || !n_windows_out) return ABFIT_ERR_ARG;
    abfit_window_args a = *args;
    Counts c;
    if (int rc = window_counts(a, c)) return rc;
    n_windows_out[0] = (int32_t)c.up_down;
    n_windows_out[1] = (int32_t)c.gene;
    n_windows_out[2] = (int32_t)c.up_down;
    return 0;
}

int abfit_place_sites(const abfit_gene *genes, int32_t n_genes, const abfit_cg_site *sites, int64_t n_sites,
                      const abfit_window_args *args, int32_t *distribution_out, int64_t *n_assign_out,
                      int64_t assign_cap, int64_t *assign_site, int32_t *assign_window)
{
    if (!args || n_genes < 0 || n_sites < 0 || (n_genes && !genes) || (n_sites && !sites)) return ABFIT_ERR_ARG;
    abfit_window_args a = *args;
    Counts cnt;
    if (int rc = window_counts(a, cnt)) return rc;
    const int64_t n_total = (int64_t)cnt.up_down * 2 + cnt.gene;

    std::map<int32_t, Lists> genome;  // Genome::new + insert_gene + sort (src/extract.rs:62-66)
    for (int32_t i = 0; i < n_genes; ++i) {
        Lists &l = genome[genes[i].chromosome];
        l.combined.push_back(genes[i]);
        if (genes[i].strand > 0) l.sense.push_back(genes[i]);
        if (genes[i].strand < 0) l.antisense.push_back(genes[i]);
    }
    auto by_start = [](const abfit_gene &x, const abfit_gene &y) { return x.start < y.start; };
    for (auto &kv : genome) {  // Vec::sort_by is a stable sort
        std::stable_sort(kv.second.sense.begin(), kv.second.sense.end(), by_start);
        std::stable_sort(kv.second.antisense.begin(), kv.second.antisense.end(), by_start);
        std::stable_sort(kv.second.combined.begin(), kv.second.combined.end(), by_start);
    }

    std::vector<int32_t> dist((size_t)n_total, 0);
    int64_t n_assign = 0;
    const double E = 0.1;  // src/methylation_site.rs:430
    const double cutoff = (double)a.cutoff, step = (double)a.window_step, size = (double)a.window_size;
    const abfit_gene *last = nullptr;

    for (int64_t si = 0; si < n_sites; ++si) {
        const abfit_cg_site &s = sites[si];
        if (!last || !is_in_gene(s, *last, a)) last = find_gene(s, genome, a);  // src/windows.rs:331-333
        if (!last) continue;
        const abfit_gene &g = *last;

        // place_in_windows (src/methylation_site.rs:423-490); Unknown-strand sites are treated as Sense (:439-443)
        const bool antisense = s.strand < 0;
        const double location = (double)s.start, start = (double)g.start, end = (double)g.end;
        const double length = end - start;
        const double offset = antisense ? end - location : location - start;
        int region;  // 0 upstream, 1 gene, 2 downstream
        if (offset < 0.0)
            region = 0;
        else if (offset > length)
            region = 2;  // a site exactly on the end of the gene is still in the gene
        else
            region = 1;
        double position;
        if (!antisense)
            position = region == 0 ? location - start + cutoff : region == 1 ? location - start : location - end;
        else
            position = region == 0 ? end - location + cutoff : region == 1 ? end - location : start - location;
        if (!a.absolute) {
            position = region == 1 ? position / length : position / cutoff;
            position *= 100.0;
        }
        const int64_t n_reg = region == 1 ? cnt.gene : cnt.up_down;
        const int64_t base = region == 0 ? 0 : region == 1 ? (int64_t)cnt.up_down : (int64_t)cnt.up_down + cnt.gene;
        if (!(position == position) || n_reg == 0) continue;  // NaN compares false with every bound
        // The reference tests every window i: lower = i*step - E, upper = lower + size + E (:480-488).  The bounds
        // grow with i, so the hits are contiguous; only a neighbourhood of position/step needs the exact test.
        const double guess = position / step;
        int64_t hi = guess >= (double)n_reg ? n_reg - 1 : guess < 0.0 ? 0 : (int64_t)guess + 2;
        if (hi > n_reg - 1) hi = n_reg - 1;
        int64_t lo = hi - (int64_t)(size / step) - 4;
        if (lo < 0) lo = 0;
        for (int64_t i = lo; i <= hi; ++i) {
            const double lower_bound = (double)i * step - E;
            const double upper_bound = lower_bound + size + E;
            if (position >= lower_bound && position <= upper_bound) {
                ++dist[(size_t)(base + i)];
                if (assign_site && assign_window && n_assign < assign_cap) {
                    assign_site[n_assign] = si;
                    assign_window[n_assign] = (int32_t)(base + i);
                }
                ++n_assign;
            }
        }
    }
    if (distribution_out) std::copy(dist.begin(), dist.end(), distribution_out);
    if (n_assign_out) *n_assign_out = n_assign;
    return 0;
}

}  // extern "C"
