// emul.cpp — HOST build of the device-side model / objective / Nelder-Mead state machine.
//
// TEST INFRASTRUCTURE ONLY (never loaded by the product package).  The per-lane device code of
// alphabeta-rs_b200/csrc/abfit_model.cuh + abfit_nm.cuh has no cross-lane communication, so it
// can be compiled for the CPU with a few shims and run one lane at a time.  That lets the
// `-m "not gpu"` suite execute the SAME source the kernels are built from (micro-op program,
// software-pipelined pair loop, NM phases) against the oracle, without a GPU.
//
//   see Makefile (g++, -ffp-contract=off, shims.h force-included into abfit_plan.cu built as C++)
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "shims.h"

#include "../../alphabeta-rs_b200/csrc/abfit_plan.h"
#include "../../alphabeta-rs_b200/csrc/abfit_wide.cuh"

namespace abfit {
static std::string g_err;
void set_error(const std::string &m) { g_err = m; }
int cuda_fail(cudaError_t, const char *) { return ABFIT_ERR_CUDA; }
}  // namespace abfit

using namespace abfit;

struct Staged {
    HostPlan hp;
    std::vector<double> lane_mem;  // (n_lane + 25) * 32
    WarpCtx c;
    LaneSimplex S;
};

static int stage(const abfit_problem *pb, Staged &s, int lane)
{
    if (int rc = compile_problems(pb, 1, s.hp)) return rc;
    const DevProblem &d = s.hp.probs[0];
    s.lane_mem.assign((size_t)(d.n_lane + 25) * 32, 0.0);
    s.c.D = s.hp.D.data() + d.d_off;
    s.c.offs = s.hp.offs.data() + d.offs_off;
    s.c.ops = s.hp.ops.data() + d.ops_off;
    s.c.lm = s.lane_mem.data();
    s.c.n_pairs = d.n_pairs;
    s.c.n_ops = d.n_ops;
    s.c.p_uu0 = d.p_uu0;
    s.c.p_mm0 = d.p_mm0;
    s.c.eqp = d.eqp;
    s.c.penw = d.penw;
    s.S.X = s.lane_mem.data() + (size_t)d.n_lane * 32 + lane;
    s.S.C = s.S.X + 20 * 32;
    return 0;
}

extern "C" {

const char *emul_last_error(void) { return g_err.c_str(); }

// objective of theta[B][4] on one problem; every theta runs on lane (i % 32)
int emul_cost(const abfit_problem *pb, const double *theta, int B, double *cost, double *lse)
{
    Staged s;
    if (int rc = stage(pb, s, 0)) return rc;
    const DBroadcast Dat{s.c.D};
    for (int i = 0; i < B; ++i) {
        const double *t = theta + 4 * (size_t)i;
        const int lane = i & 31;
        cost[i] = objective(s.c, Dat, lane, t[0], t[1], t[2], t[3], true);
        if (lse) lse[i] = objective(s.c, Dat, lane, t[0], t[1], t[2], t[3], false);
    }
    return 0;
}

// one Nelder-Mead run per simplex ([n][5][4]), the state machine driven exactly as the kernels drive it.
// dstar: optional per-fit D* columns [n][n_pairs] (bootstrap replicates), else the problem's D.
int emul_fit(const abfit_problem *pb, const double *simplices, int n, const double *dstar, int max_iters,
             double sd_tol, uint32_t flags, abfit_fit *out)
{
    const NMParams nm = nm_params(max_iters, sd_tol, flags);
    for (int f = 0; f < n; ++f) {
        const int lane = f & 31;
        Staged s;
        if (int rc = stage(pb, s, lane)) return rc;
        std::vector<double> tile;
        if (dstar) {
            tile.assign((size_t)s.c.n_pairs * 32, 0.0);
            for (int i = 0; i < s.c.n_pairs; ++i) tile[(size_t)i * 32 + lane] = dstar[(size_t)f * s.c.n_pairs + i];
        }
        LaneNM L;
        std::memset(&L, 0, sizeof L);
        for (int q = 0; q < 20; ++q) s.S.X[q * 32] = simplices[(size_t)f * 20 + q];
        nm_begin(L, s.S, f);
        abfit_fit res;
        std::memset(&res, 0, sizeof res);
        for (;;) {
            double v;
            if (dstar) {
                const DLaneColumn Dat{tile.data() + lane};
                v = objective(s.c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            } else {
                const DBroadcast Dat{s.c.D};
                v = objective(s.c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            }
            if (nm_advance(L, s.S, nm, v, res, 1u << lane)) break;
        }
        out[f] = res;
    }
    return 0;
}

// ---- warp-per-fit formulation (abfit_wide.cuh), run with a "warp" of ONE lane (ABFIT_WIDE_WIDTH=1): same
// tables, same chunked pair sum, same Nelder-Mead driver as k_fit_wide
struct StagedWide {
    HostPlan hp;
    std::vector<double> mem;
    WideCtx c;
    LaneSimplex S;
};

static int stage_wide(const abfit_problem *pb, StagedWide &s)
{
    if (int rc = compile_problems(pb, 1, s.hp)) return rc;
    const DevProblem &d = s.hp.probs[0];
    s.mem.assign(wide_warp_doubles(d.tmax, d.n_trip), 0.0);
    double *after = wide_carve(s.c, s.mem.data(), d.tmax, d.n_trip);
    s.S.X = after;
    s.S.C = after + 20 * 32;
    s.c.trip = s.hp.wtrip.data() + d.wtrip_off;
    s.c.tid = s.hp.wtid.data() + d.wtid_off;
    s.c.D = s.hp.D.data() + d.d_off;
    s.c.n_pairs = d.n_pairs;
    s.c.n_trip = d.n_trip;
    s.c.tmax = d.tmax;
    s.c.ss = nullptr;  // the exact sum (the sufficient-statistics experiment is a GPU-side flag)
    s.c.p_uu0 = d.p_uu0;
    s.c.p_mm0 = d.p_mm0;
    s.c.eqp = d.eqp;
    s.c.penw = d.penw;
    return 0;
}

int emul_cost_wide(const abfit_problem *pb, const double *theta, int B, double *cost, double *lse)
{
    StagedWide s;
    if (int rc = stage_wide(pb, s)) return rc;
    for (int i = 0; i < B; ++i) {
        const double *t = theta + 4 * (size_t)i;
        cost[i] = objective_wide(s.c, 0, t[0], t[1], t[2], t[3], true);
        if (lse) lse[i] = objective_wide(s.c, 0, t[0], t[1], t[2], t[3], false);
    }
    return 0;
}

int emul_fit_wide(const abfit_problem *pb, const double *simplices, int n, const double *dstar, int max_iters,
                  double sd_tol, uint32_t flags, abfit_fit *out)
{
    const NMParams nm = nm_params(max_iters, sd_tol, flags);
    StagedWide s;
    if (int rc = stage_wide(pb, s)) return rc;
    for (int f = 0; f < n; ++f) {
        if (dstar) s.c.D = dstar + (size_t)f * s.c.n_pairs;
        LaneNM L;
        std::memset(&L, 0, sizeof L);
        for (int q = 0; q < 20; ++q) s.S.X[q * 32] = simplices[(size_t)f * 20 + q];
        nm_begin(L, s.S, f);
        abfit_fit res;
        std::memset(&res, 0, sizeof res);
        for (;;) {
            const double v = objective_wide(s.c, 0, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            if (nm_advance(L, s.S, nm, v, res, 1u)) break;
        }
        out[f] = res;
    }
    return 0;
}

// launch shape the library would choose on a B200 (227 KB opt-in shared memory per block, 228 KB per SM)
int emul_launch_shape(const abfit_problem *pb, int n_probs, int fits_per_prob, int64_t out[10])
{
    HostPlan hp;
    if (int rc = compile_problems(pb, n_probs, hp)) return rc;
    LaunchShape sh;
    if (int rc = choose_launch_shape(hp, 227 * 1024, 228 * 1024, fits_per_prob, sh)) return rc;
    out[0] = sh.n_warps;
    out[1] = sh.d_shared;
    out[2] = sh.x_global;
    out[3] = sh.big;
    out[4] = sh.wide;
    out[5] = sh.boot_x_global;
    out[6] = (int64_t)sh.smem_fit;
    out[7] = (int64_t)sh.smem_wide;
    out[8] = (int64_t)sh.smem_boot_gather;
    out[9] = (int64_t)(sh.wide ? make_items_wide(hp, fits_per_prob, 148, true).size()
                               : make_items(hp, fits_per_prob, 148, sh.n_warps, true).size());
    return 0;
}

double emul_var_threshold(double sd_tol) { return nm_var_threshold(sd_tol); }
double emul_range_threshold(double var_thr) { return nm_range_threshold(var_thr); }

// plan statistics: per-lane doubles, micro-ops, events, chain length
int emul_plan_stats(const abfit_problem *pb, int32_t out[6])
{
    HostPlan hp;
    if (int rc = compile_problems(pb, 1, hp)) return rc;
    out[0] = hp.probs[0].n_lane;
    out[1] = hp.probs[0].n_ops;
    out[2] = 0;
    out[3] = hp.probs[0].tmax;
    out[4] = hp.n_triples[0];
    out[5] = (int32_t)smem_need(hp.probs[0], 25, true, 1);
    return 0;
}
}
