// abfit_plan.cu — pedigree -> micro-op program (host).  O(pairs) bookkeeping; no numerics.
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <thread>
#include <map>
#include <set>
#include <tuple>

#include "abfit_plan.h"
#include "abfit_wide.cuh"

namespace abfit {

// `x as i8` in Rust (saturating, NaN -> 0), src/divergence.rs:52
static inline int as_i8(double x)
{
    if (x != x) return 0;
    if (x >= 127.0) return 127;
    if (x <= -128.0) return -128;
    return (int)x;
}

static inline OpWord mk_op(uint32_t op, uint32_t a = 0, uint32_t b = OP_NONE, uint32_t c = OP_NONE)
{
    OpWord w;
    w.x = op | (a << 16);  // n_step (bits 8..15) is filled in when the program is flattened
    w.y = b | (c << 16);
    return w;
}

// ---- shared-memory footprints (the carve-up itself is carve_and_stage in abfit_kernels.cu) ----------
size_t smem_need(const DevProblem &pb, int simplex_doubles, bool d_shared, int n_warps)
{
    size_t b = (size_t)n_warps * ((size_t)pb.n_lane + (size_t)simplex_doubles) * 32 * 8;
    if (d_shared) b += (((size_t)pb.n_pairs + 1) & ~(size_t)1) * 8;
    b += (size_t)pb.n_offs * 4;
    b += (size_t)pb.n_ops * 8;
    b += 16;
    return (b + 15) & ~(size_t)15;
}

size_t smem_need_wide(const DevProblem &pb)
{
    return ((size_t)wide_warp_doubles(pb.tmax, pb.n_trip) * 8 + (size_t)pb.n_trip * 4 + 15) & ~(size_t)15;
}

size_t smem_need_boot_gather(const DevProblem &pb, int simplex_doubles)
{
    return smem_need(pb, simplex_doubles, false, 1) + (size_t)boot_gather_lead(pb.n_pairs) * 8;
}

// run f(p) for p in [0, n) on the host's cores (row scans of large batches; small ones stay on the caller's thread)
template <class F>
static void scan_parallel(int n, F f)
{
    const unsigned nt = std::min<unsigned>(std::max(1u, std::thread::hardware_concurrency()), (unsigned)(n / 512));
    if (nt <= 1) {
        for (int p = 0; p < n; ++p) f(p);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
        th.emplace_back([=]() {
            const int lo = (int)((int64_t)n * t / nt), hi = (int)((int64_t)n * (t + 1) / nt);
            for (int p = lo; p < hi; ++p) f(p);
        });
    for (auto &x : th) x.join();
}

int compile_problems(const abfit_problem *probs, int n_probs, HostPlan &hp)
{
    hp.clear();
    if (!probs || n_probs <= 0) {
        set_error("no problems");
        return ABFIT_ERR_ARG;
    }
    hp.probs.resize(n_probs);
    hp.flops.resize(n_probs);
    hp.fp64_instr.resize(n_probs);
    hp.tmax.resize(n_probs);
    hp.n_triples.resize(n_probs);
    hp.d_has_nan.assign(n_probs, 0);
    typedef std::tuple<int, int, int> Tri;  // (t0, a = t1 - t0, b = t2 - t0)
    typedef std::pair<int, int> AB;

    // ---- pass 1: pool layout, then one scan of every pedigree row (parallel over problems) ------------
    std::vector<int64_t> d_off(n_probs), pair_off(n_probs);
    int64_t d_total = 0;
    for (int p = 0; p < n_probs; ++p) {
        if (!probs[p].pedigree || probs[p].n_pairs <= 0) {
            set_error("problem " + std::to_string(p) + ": empty pedigree");
            return ABFIT_ERR_ARG;
        }
        d_off[p] = d_total;  // even: 16-byte aligned columns
        pair_off[p] = hp.total_pairs;
        d_total += ((int64_t)probs[p].n_pairs + 1) & ~(int64_t)1;
        hp.total_pairs += probs[p].n_pairs;
        hp.max_pairs = std::max(hp.max_pairs, probs[p].n_pairs);
    }
    hp.D.resize((size_t)d_total);  // every element is written by the scan below (the pad of an odd column too)
    std::vector<uint32_t> &keys = hp.keys;
    keys.resize((size_t)hp.total_pairs);
    std::vector<int32_t> bad_row(n_probs, -1), max_exp_of(n_probs, 0);
    scan_parallel(n_probs, [&](int p) {
        const abfit_problem &ap = probs[p];
        uint32_t *k = keys.data() + pair_off[p];
        double *Dp = hp.D.data() + d_off[p];
        int mx = 0;
        uint8_t nan = 0;
        for (int i = 0; i < ap.n_pairs; ++i) {
            const double *row = ap.pedigree + 4 * (size_t)i;
            const int t0 = as_i8(row[0]), t1 = as_i8(row[1]), t2 = as_i8(row[2]);
            if (t0 < 0 || t1 < t0 || t2 < t0) {
                bad_row[p] = i;
                return;
            }
            k[i] = (uint32_t)t0 | ((uint32_t)(t1 - t0) << 8) | ((uint32_t)(t2 - t0) << 16);
            mx = std::max(mx, std::max(t0, std::max(t1 - t0, t2 - t0)));
            nan |= (uint8_t)(row[3] != row[3]);
            Dp[i] = row[3];
        }
        if (ap.n_pairs & 1) Dp[ap.n_pairs] = 0.0;
        max_exp_of[p] = mx;
        hp.d_has_nan[p] = nan;
    });
    // does problem p repeat the time structure of problem p - 1?  (parallel; equality is transitive, so "same as the
    // previous problem" is "same as the last distinct one" in the loop below)
    std::vector<uint8_t> same_as_prev(n_probs, 0);
    scan_parallel(n_probs, [&](int p) {
        if (p > 0 && probs[p - 1].n_pairs == probs[p].n_pairs && bad_row[p] < 0 && bad_row[p - 1] < 0)
            same_as_prev[p] = std::memcmp(keys.data() + pair_off[p], keys.data() + pair_off[p - 1], (size_t)probs[p].n_pairs * 4) == 0;
    });

    // ---- pass 2: programs ------------------------------------------------------------------------------
    // Windows of one metaprofile share the pedigree's time structure (same nodes and edges, only D differs):
    // a problem whose (t0,t1,t2) sequence equals the previous one's shares its program and offsets in the pools.
    int prev_p = -1;
    for (int p = 0; p < n_probs; ++p) {
        const abfit_problem &ap = probs[p];
        DevProblem dp;
        dp.d_off = d_off[p];
        dp.pair_off = pair_off[p];
        dp.offs_off = (int64_t)hp.offs.size();
        dp.ops_off = (int64_t)hp.ops.size();
        dp.wtrip_off = (int64_t)hp.wtrip.size();
        dp.wtid_off = (int64_t)hp.wtid.size();
        dp.n_pairs = ap.n_pairs;
        dp.p_uu0 = ap.p0uu;
        dp.p_mm0 = 1.0 - ap.p0uu;               // src/ab_neutral.rs:23
        if (dp.p_mm0 + dp.p_uu0 + 0.0 != 1.0) {  // src/ab_neutral.rs:31 assert_eq!
            set_error("problem " + std::to_string(p) + ": p0mm + p0uu + p0um != 1");
            return ABFIT_ERR_NAN;
        }
        if (bad_row[p] >= 0) {
            set_error("problem " + std::to_string(p) + " row " + std::to_string(bad_row[p]) +
                      ": needs 0 <= t0 <= t1,t2 <= 127 (the reference would invert the matrix)");
            return ABFIT_ERR_TIME;
        }
        dp.eqp = ap.eqp;
        dp.penw = ap.eqp_weight * (double)ap.n_pairs;  // src/structs.rs:210-211
        const uint32_t *key32 = keys.data() + pair_off[p];
        const int max_exp = max_exp_of[p];
        if (prev_p >= 0 && same_as_prev[p]) {
            const DevProblem &q = hp.probs[prev_p];
            dp.offs_off = q.offs_off;
            dp.ops_off = q.ops_off;
            dp.n_offs = q.n_offs;
            dp.n_ops = q.n_ops;
            dp.n_lane = q.n_lane;
            dp.tmax = q.tmax;
            dp.n_trip = q.n_trip;
            dp.wtrip_off = q.wtrip_off;
            dp.wtid_off = q.wtid_off;
            hp.probs[p] = dp;
            hp.tmax[p] = hp.tmax[prev_p];
            hp.n_triples[p] = hp.n_triples[prev_p];
            hp.flops[p] = hp.flops[prev_p];
            hp.fp64_instr[p] = hp.fp64_instr[prev_p];
            continue;
        }
        prev_p = p;
        // ---- distinct triples --------------------------------------------------------------
        std::vector<Tri> key(ap.n_pairs);
        std::map<Tri, int> tri_id;
        for (int i = 0; i < ap.n_pairs; ++i)
            key[i] = Tri((int)(key32[i] & 0xff), (int)((key32[i] >> 8) & 0xff), (int)((key32[i] >> 16) & 0xff));
        for (int i = 0; i < ap.n_pairs; ++i) tri_id.emplace(key[i], 0);
        int U = 0;
        for (auto &kv : tri_id) kv.second = U++;

        // ---- storage layout: [G^m matrices][s vectors][deferred d vectors][dt per triple] ----
        std::map<AB, std::vector<Tri>> by_ab;
        std::set<int> Mset, T0set;
        for (auto &kv : tri_id) {
            const int t0 = std::get<0>(kv.first), a = std::get<1>(kv.first), b = std::get<2>(kv.first);
            by_ab[AB(a, b)].push_back(kv.first);
            if (a != b && std::min(a, b) >= 2) Mset.insert(std::min(a, b));
            if (t0 >= 1) T0set.insert(t0);
        }
        std::map<int, uint32_t> idx_M, idx_S;
        std::map<AB, uint32_t> idx_D;
        uint32_t n_lane = 0;
        for (int m : Mset) {
            idx_M[m] = n_lane;
            n_lane += 9;
        }
        for (int t : T0set) {
            idx_S[t] = n_lane;
            n_lane += 3;
        }
        for (auto &kv : by_ab) {
            const int k = std::max(kv.first.first, kv.first.second);
            bool deferred = false;
            for (auto &t : kv.second) deferred |= std::get<0>(t) > k;
            if (deferred) {
                idx_D[kv.first] = n_lane;
                n_lane += 3;
            }
        }
        const uint32_t dt_base = n_lane;
        n_lane += (uint32_t)U;
        if (n_lane >= SRC_SPECIAL) {
            set_error("problem " + std::to_string(p) + ": too many distinct (t0,t1,t2) triples");
            return ABFIT_ERR_TOO_LARGE;
        }

        // ---- micro-ops per chain exponent ------------------------------------------------------
        struct Ev {
            std::vector<OpWord> head, body, tail;  // [STORE_M, CALC_S] [per exponent pair] [deferred triples]
        };
        std::map<int, Ev> evs;
        for (int m : Mset) evs[m].head.push_back(mk_op(OP_STORE_M, idx_M[m]));
        for (int t : T0set) evs[t].head.push_back(mk_op(OP_CALC_S, idx_S[t]));
        auto src_of = [&](int e, int k) -> uint32_t {
            if (e == k) return e == 0 ? SRC_IDENT : SRC_CUR;
            if (e == 0) return SRC_IDENT;
            if (e == 1) return SRC_G;
            return idx_M[e];
        };
        // the D op of exponent pair (a, b) with its fused tail: dt0 = dt slot of the t0 == 0 triple,
        // keep = slot of the d-vector when a later t0 needs it
        auto op_d = [&](int a, int b, int k, uint32_t dt0, uint32_t keep, std::vector<OpWord> &out) {
            const uint32_t sa = src_of(a, k), sb = src_of(b, k);
            if (sa == SRC_CUR && sb == SRC_CUR) return out.push_back(mk_op(OP_D_CC, 0, dt0, keep));
            if (sb == SRC_CUR) {
                if (sa == SRC_G) return out.push_back(mk_op(OP_D_GC, 0, dt0, keep));
                if (sa < SRC_SPECIAL) return out.push_back(mk_op(OP_D_MC, sa, dt0, keep));
            } else if (sa == SRC_CUR) {
                if (sb == SRC_G) return out.push_back(mk_op(OP_D_CG, 0, dt0, keep));
                if (sb < SRC_SPECIAL) return out.push_back(mk_op(OP_D_CM, sb, dt0, keep));
            }
            out.push_back(mk_op(OP_D_GEN, sa, sb));
            if (keep != OP_NONE) out.push_back(mk_op(OP_STORE_D, keep));
            if (dt0 != OP_NONE) out.push_back(mk_op(OP_DT0, dt0));
        };
        for (auto &kv : by_ab) {
            const int a = kv.first.first, b = kv.first.second, k = std::max(a, b);
            Ev &ev = evs[k];
            uint32_t dt0 = OP_NONE;
            for (auto &t : kv.second)
                if (std::get<0>(t) == 0) dt0 = dt_base + (uint32_t)tri_id[t];
            op_d(a, b, k, dt0, idx_D.count(kv.first) ? idx_D[kv.first] : OP_NONE, ev.body);
            for (auto &t : kv.second) {
                const int t0 = std::get<0>(t);
                if (t0 > 0 && t0 <= k) ev.body.push_back(mk_op(OP_DT, dt_base + (uint32_t)tri_id[t], idx_S[t0]));
            }
        }
        for (auto &kv : by_ab) {  // triples whose t0 comes after their exponent pair
            const int k = std::max(kv.first.first, kv.first.second);
            for (auto &t : kv.second)
                if (std::get<0>(t) > k)
                    evs[std::get<0>(t)].tail.push_back(
                        mk_op(OP_LDT, dt_base + (uint32_t)tri_id[t], idx_S[std::get<0>(t)], idx_D[kv.first]));
        }
        // flatten; the chain starts at R = G^1 and the steps up to an exponent ride on its first op
        int tmax = 0, cur_k = 1;
        for (auto &kv : evs) {
            const int k = kv.first;
            const size_t first = hp.ops.size();
            for (auto &o : kv.second.head) hp.ops.push_back(o);
            for (auto &o : kv.second.body) hp.ops.push_back(o);
            for (auto &o : kv.second.tail) hp.ops.push_back(o);
            if (hp.ops.size() == first) continue;
            if (k > cur_k) {
                hp.ops[first].x |= (uint32_t)(k - cur_k) << 8;  // <= 127 by the time validation above
                cur_k = k;
            }
            tmax = std::max(tmax, k);
        }
        hp.ops.push_back(mk_op(OP_NOP));  // end sentinel: the interpreter prefetches one word ahead
        dp.n_ops = (int32_t)(hp.ops.size() - (size_t)dp.ops_off);
        dp.n_lane = (int32_t)n_lane;
        dp.tmax = tmax;
        dp.n_trip = U;
        for (auto &kv : tri_id)  // map order == id order
            hp.wtrip.push_back((uint32_t)std::get<0>(kv.first) | ((uint32_t)std::get<1>(kv.first) << 8) |
                               ((uint32_t)std::get<2>(kv.first) << 16));
        for (int i = 0; i < ap.n_pairs; ++i) hp.wtid.push_back((uint32_t)tri_id[key[i]]);

        for (int i = 0; i < ap.n_pairs; ++i) hp.offs.push_back(256u * (dt_base + (uint32_t)tri_id[key[i]]));
        while (hp.offs.size() & 3) hp.offs.push_back(256u * dt_base);
        dp.n_offs = (int32_t)(hp.offs.size() - (size_t)dp.offs_off);

        hp.probs[p] = dp;
        hp.tmax[p] = max_exp;
        hp.n_triples[p] = U;
        hp.flops[p] = 45.0 * (max_exp > 1 ? max_exp - 1 : 0) + 56.0 * U + 5.0 * ap.n_pairs + 40.0;
        // FP64 instructions per evaluation and lane as the kernels execute them (FMA = one instruction): 27 per chain
        // step, 36 per conditional-divergence op, 9 per sv0.G^t0, 5 per dt1t2, 5 per pair, 48 for genmatrix / sv0 /
        // sv0.I / p_uu_est / penalty.  The DFMA peak is one such instruction per lane and slot, so
        // fp64_instr / (flops / 2) is the factor by which the bit-exact formulation (FMA only in the 3x3 dots)
        // stays below the FMA roofline even with a perfectly busy pipe.
        {
            double n = 48.0 + 5.0 * ap.n_pairs;
            for (size_t i = (size_t)dp.ops_off; i < hp.ops.size(); ++i) {
                const uint32_t op = hp.ops[i].x & 0xff, steps = (hp.ops[i].x >> 8) & 0xff, bb = hp.ops[i].y & 0xffff;
                n += 27.0 * steps;
                if (op >= OP_D_CC && op <= OP_D_GEN) n += 36.0 + ((op != OP_D_GEN && bb != OP_NONE) ? 5.0 : 0.0);
                else if (op == OP_CALC_S) n += 9.0;
                else if (op == OP_DT0 || op == OP_DT || op == OP_LDT) n += 5.0;
            }
            hp.fp64_instr[p] = n;
        }
    }
    return 0;
}

double nm_var_threshold(double sd_tol)
{
    if (!(sd_tol > 0.0)) return -1.0;  // sqrt(y) >= 0 is never below a tolerance <= 0 (or NaN)
    if (std::isinf(sd_tol)) return DBL_MAX;
    double y = sd_tol * sd_tol;
    if (std::isinf(y)) y = DBL_MAX;
    while (y > 0.0 && !(std::sqrt(y) < sd_tol)) y = std::nextafter(y, 0.0);
    for (;;) {
        const double up = std::nextafter(y, INFINITY);
        if (std::isinf(up) || !(std::sqrt(up) < sd_tol)) break;
        y = up;
    }
    return std::sqrt(y) < sd_tol ? y : -1.0;
}

double nm_range_threshold(double var_thr)
{
    if (var_thr < 0.0) return -1.0;  // the sd test can never fire: every spread (>= 0) rejects, NaN costs take the exact path
    const double t = 8.0 * std::sqrt(var_thr);  // 4 sqrt(var_thr) suffices in exact arithmetic; the factor 2 covers every rounding
    // spreads whose square is not a normal double are left to the exact test
    if (!(t >= 1e-140) || std::isinf(t)) return INFINITY;
    return t;
}

NMParams nm_params(int max_iters, double sd_tol, uint32_t flags)
{
    NMParams nm;
    nm.max_iters = max_iters;
    nm.sd_tol = sd_tol;
    nm.flags = flags;
    nm.var_thr = nm_var_threshold(sd_tol);
    nm.range_thr = nm_range_threshold(nm.var_thr);
    return nm;
}

int choose_launch_shape(const HostPlan &hp, size_t smem_cap, size_t smem_per_sm, int fits_per_prob, LaunchShape &out)
{
    auto worst = [&](int simplex_doubles, bool d_shared, int nw) {
        size_t m = 0;
        for (auto &pb : hp.probs) m = std::max(m, smem_need(pb, simplex_doubles, d_shared, nw));
        return m;
    };
    // Multi-start kernel: as many resident warps as shared memory and registers allow.  With the
    // vertices in shared memory the kernel is built for 3 blocks x 4 warps (168 registers); with the
    // vertices in global scratch for 4 x 4 (128 registers).
    int best_w = -1;
    const int cand_nw[4] = {4, 3, 2, 1};
    const int64_t total_fits = (int64_t)hp.probs.size() * fits_per_prob;
    const char *force = getenv("ABFIT_DEV_NWARPS");  // tuning experiments only
    const char *force_x = getenv("ABFIT_DEV_XGLOBAL");
    for (int pass = 0; pass < 2; ++pass) {  // pass 0: D in shared memory; pass 1: D broadcast from L1/L2
        for (int xg = 0; xg < 2; ++xg) {
            if (force_x && xg != atoi(force_x)) continue;
            for (int nw : cand_nw) {
                if (force && nw != atoi(force)) continue;
                if (!force && nw > 1 && fits_per_prob < 64 * nw) continue;  // too few fits for a multi-warp block
                // a small batch (one pedigree x 1000 starts) is spread over as many SMs as it has warps
                if (!force && nw > 1 && total_fits < (int64_t)148 * 64 * nw) continue;
                const size_t s = worst(xg ? 5 : 25, pass == 0, nw);
                if (s > smem_cap) continue;
                const int reg_warps = xg ? 16 : 12;
                int w = (int)std::min<size_t>((size_t)reg_warps, (smem_per_sm / (s + 1024)) * (size_t)nw);
                // measured on the C4 shape: 16 warps with the vertices in L2 are no faster than 12 with them
                // in shared memory (the kernel is not latency-bound there), so the vertices only move out
                // when shared memory would otherwise leave fewer than 8 resident warps
                // equal residency: fewer warps per block = more fits per lane of a window's queue (shorter tails)
                if (xg ? (best_w < 8 && w > best_w) : (w > best_w || (w == best_w && !out.x_global && nw < out.n_warps))) {
                    best_w = w;
                    out.n_warps = nw;
                    out.d_shared = pass == 0;
                    out.x_global = xg != 0;
                    out.smem_fit = s;
                }
            }
        }
        if (best_w >= 4) break;  // only give up the shared D column when occupancy would collapse
    }
    out.n_lane_max = 0;
    size_t max_ops = 0;
    for (auto &pb : hp.probs) {
        out.n_lane_max = std::max(out.n_lane_max, pb.n_lane);
        max_ops = std::max(max_ops, (size_t)pb.n_ops);
    }
    out.smem_boot = worst(25, false, 1);
    out.d_shared_aux = worst(0, true, 1) <= smem_cap / 2;
    out.smem_aux = worst(0, out.d_shared_aux, 1);
    out.big = best_w <= 0 || out.smem_boot > smem_cap || out.smem_aux > smem_cap || getenv("ABFIT_DEV_BIG");
    out.wide = false;
    out.smem_wide = 0;
    {
        size_t m = 0;
        for (auto &pb : hp.probs) m = std::max(m, smem_need_wide(pb));
        const char *fw = getenv("ABFIT_DEV_WIDE");  // tests: force the warp-per-fit kernels on (1) or off (0)
        if (m <= smem_cap && (fw ? atoi(fw) != 0 : out.big)) {
            out.wide = true;
            out.smem_wide = m;
        }
    }
    if (out.big) {
        // lane state, vertices, D and offsets in global scratch (carve_big): shared memory only holds the simplex
        // costs, the program and the queue word
        out.n_warps = (force ? atoi(force) : (fits_per_prob >= 256 && total_fits >= (int64_t)148 * 64 * 4 ? 4 : 1));
        out.d_shared = false;
        out.x_global = true;
        out.smem_fit = ((size_t)out.n_warps * 5 * 32 * 8 + max_ops * 8 + 16 + 15) & ~(size_t)15;
        out.smem_boot = ((size_t)5 * 32 * 8 + max_ops * 8 + 16 + 15) & ~(size_t)15;
        out.smem_aux = (max_ops * 8 + 16 + 15) & ~(size_t)15;
        out.d_shared_aux = false;
        out.smem_boot_gather = 0;
        if (out.smem_fit > smem_cap) {
            set_error("pedigree program too large for shared memory");
            return ABFIT_ERR_TOO_LARGE;
        }
        return 0;
    }
    out.smem_boot_gather = 0;
    out.boot_x_global = false;
    if (hp.max_pairs <= 8191) {  // u16 byte offsets into resid
        size_t m25 = 0, m5 = 0;
        for (auto &pb : hp.probs) {
            m25 = std::max(m25, smem_need_boot_gather(pb, 25));
            m5 = std::max(m5, smem_need_boot_gather(pb, 5));
        }
        if (m25 <= smem_cap / 4 && !getenv("ABFIT_DEV_BOOT_TILE")) {  // else: stored-D* kernel
            // one-warp blocks, 158 registers: at most 12 fit on an SM.  The simplex vertices move to a global scratch
            // area when that buys resident warps (measured on the C4 shape: 8 -> 11 warps per SM, bootstrap -4 %)
            const size_t b25 = std::min<size_t>(12, smem_per_sm / (m25 + 1024)), b5 = std::min<size_t>(12, smem_per_sm / (m5 + 1024));
            const char *fx = getenv("ABFIT_DEV_BOOT_XGLOBAL");
            // ... and only when there are more blocks than resident slots (one block per window and ~128 replicates):
            // a batch that is resident all at once gains nothing and pays the global-memory vertices
            out.boot_x_global = fx ? atoi(fx) != 0 : (b5 > b25 && hp.probs.size() > (size_t)148 * b25);
            out.smem_boot_gather = out.boot_x_global ? m5 : m25;
        }
    }
    return 0;
}

std::vector<WorkItem> make_items_wide(const HostPlan &hp, int count_per_prob, int n_sm, bool skip_nan)
{
    // one warp per block, one fit at a time per warp: the kernel is bound by the latency of the sequential pair
    // sum, so the batch is cut into about three waves of 14 warps per SM (more do not add throughput: 14 chains
    // saturate the FP64 pipe) and a batch smaller than that gets one fit per warp
    const int n_probs = (int)hp.probs.size();
    const int64_t total = (int64_t)n_probs * count_per_prob;
    const int64_t target_blocks = (int64_t)n_sm * 14 * 3;
    int64_t chunk = std::max<int64_t>(1, (total + target_blocks - 1) / target_blocks);
    chunk = std::min<int64_t>(chunk, count_per_prob);
    std::vector<WorkItem> items;
    for (int p = 0; p < n_probs; ++p) {
        if (skip_nan && hp.d_has_nan[p]) continue;
        for (int f = 0; f < count_per_prob; f += (int)chunk) {
            WorkItem it;
            it.prob = p;
            it.first = f;
            it.count = std::min<int>((int)chunk, count_per_prob - f);
            it.pad = 0;
            items.push_back(it);
        }
    }
    return items;
}

std::vector<WorkItem> make_items_guided(const HostPlan &hp, int count_per_prob, int resident_warps, int max_chunk,
                                        bool skip_nan, int p_first, int p_end)
{
    const int n_probs = p_end < 0 ? (int)hp.probs.size() : p_end;
    int64_t remaining = 0;
    for (int p = p_first; p < n_probs; ++p)
        if (!(skip_nan && hp.d_has_nan[p])) remaining += count_per_prob;
    const int64_t P2 = 2 * (int64_t)std::max(resident_warps, 1);
    if (const char *fc = getenv("ABFIT_DEV_CHUNK")) max_chunk = std::max(1, atoi(fc));
    std::vector<WorkItem> items;
    for (int p = p_first; p < n_probs; ++p) {
        if (skip_nan && hp.d_has_nan[p]) continue;
        for (int f = 0; f < count_per_prob;) {
            int64_t chunk = remaining / P2;
            chunk = std::min<int64_t>(std::max<int64_t>(chunk, 32), max_chunk);
            chunk = (chunk / 32) * 32 > 0 ? (chunk / 32) * 32 : chunk;  // whole rounds of a warp
            int c = (int)std::min<int64_t>(chunk, count_per_prob - f);
            if (count_per_prob - f - c < 16 && count_per_prob - f - c > 0) c = count_per_prob - f;  // no crumbs
            WorkItem it;
            it.prob = p;
            it.first = f;
            it.count = c;
            it.pad = 0;
            items.push_back(it);
            f += c;
            remaining -= c;
        }
    }
    return items;
}

std::vector<WorkItem> make_items(const HostPlan &hp, int count_per_prob, int n_sm, int n_warps, bool skip_nan)
{
    // One block per item.  Enough blocks to fill the machine several times over, but chunks as
    // long as possible so that idle lanes can pull further fits of the same window.
    const int lanes = 32 * n_warps;
    const int64_t target_blocks = (int64_t)n_sm * (16 / n_warps) * 3;
    const int n_probs = (int)hp.probs.size();
    const int64_t total = (int64_t)n_probs * count_per_prob;
    int64_t chunk = (total + target_blocks - 1) / target_blocks;
    chunk = ((chunk + lanes - 1) / lanes) * lanes;
    // at least two fits per lane so that the queue can even out the run lengths — unless the whole batch is
    // smaller than the machine, where one fit per lane on more SMs finishes sooner
    if (chunk < 2 * lanes && total >= (int64_t)n_sm * 2 * lanes) chunk = 2 * lanes;
    if (chunk > count_per_prob) chunk = count_per_prob;
    if (const char *fc = getenv("ABFIT_DEV_CHUNK")) chunk = std::max(1, std::min(atoi(fc), count_per_prob));
    // nearest, not ceiling: 1000 starts against a chunk of 960 are one block, not two halves (longer queues per
    // block even out the run lengths better than a few more blocks do)
    // (only when a block's queue is deep: with one or two fits per lane an extra fit doubles a lane's work)
    const int n_chunks = chunk >= 4 * lanes ? (int)std::max<int64_t>(1, (count_per_prob + chunk / 2) / chunk)
                                            : (int)((count_per_prob + chunk - 1) / chunk);
    chunk = (count_per_prob + n_chunks - 1) / n_chunks;  // even split
    std::vector<WorkItem> items;
    for (int p = 0; p < n_probs; ++p) {
        if (skip_nan && hp.d_has_nan[p]) continue;
        for (int f = 0; f < count_per_prob; f += (int)chunk) {
            WorkItem it;
            it.prob = p;
            it.first = f;
            it.count = std::min<int>((int)chunk, count_per_prob - f);
            it.pad = 0;
            items.push_back(it);
        }
    }
    return items;
}

}  // namespace abfit
