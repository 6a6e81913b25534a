// abfit_device.cuh — device-side model, objective and Nelder-Mead state machine.
//
// Execution model (B200, sm_100a): one LANE per fit.  A warp works on 32 fits of the
// same window (different starts / bootstrap replicates), so the heavy part — the
// objective — is perfectly uniform across the warp: every lane walks the same power
// chain, the same (t0,t1,t2) triples and the same pedigree rows, on its own theta.
// The pedigree D column is staged in shared memory and read as a warp broadcast.
// Nothing in the objective needs a cross-lane exchange, so the FP64 pipe sees 32
// independent streams per warp and the pair sum can be accumulated SEQUENTIALLY in
// pedigree order, exactly as the reference does (src/structs.rs:208-213) — results
// are bit-identical to the CPU restatement instead of "close".
//
// Compile with -fmad=false: the only fused operations are the explicit __fma_rn
// chains of the 3x3 products (the pattern that reproduces the reference's cost KAT).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abfit.h"

namespace abfit {

constexpr unsigned FULL = 0xffffffffu;

// One window, compiled on the host (compile_problems in abfit_api.cu): the distinct
// (t0,t1,t2) triples, the id of every pair's triple in pedigree order, and the exponents
// whose powers of G the triples need.
struct DevProblem {
    int64_t d_off;     // into the D pool (even: 16-byte aligned)
    int64_t pair_off;  // into pred / resid (problems concatenated without padding)
    int64_t ids_off;   // into the pair-offset pool (u32 = 256 * triple id per pair; padded to a multiple of 4 per problem)
    int64_t tri_off;   // into the triple pool (u32: slot_t0 | slot_a << 8 | slot_b << 16)
    int64_t exp_off;   // into the exponent pool (u8, ascending, all > 0); slot s>0 = exps[s-1], slot 0 = G^0
    int32_t n_pairs, n_ids, n_triples, n_exps;  // n_ids = n_pairs rounded up to 4
    double p_uu0, p_mm0;  // state at G0 (p0um = 0), src/ab_neutral.rs:23-24
    double eqp, penw;     // penw = eqp_weight * (double)n_pairs, src/structs.rs:210-211
};

// per-warp view of the staged problem
struct WarpCtx {
    const double *D;       // [n_pairs] shared (or global in the large-problem variant)
    const uint32_t *offs;  // shared, 16-byte aligned: 256 * triple id of every pair (byte offset into a dt column)
    const uint32_t *tris;  // shared
    const uint8_t *exps;   // shared
    double *pw;            // per-lane power slots: element e of slot s at pw[((s-1)*9+e)*32 + lane]
    double *dt;            // per-lane theoretical divergence per triple: dt[u*32 + lane]
    int32_t n_pairs, n_triples, n_exps;
    double p_uu0, p_mm0, eqp, penw;
};

// src/divergence.rs:96-114 (powi(2) == x*x)
__device__ __forceinline__ void genmatrix(double a, double b, double G[9])
{
    double oma = 1.0 - a, omb = 1.0 - b;
    double b1a = b + 1.0 - a;
    double a1b = a + 1.0 - b;
    G[0] = oma * oma;
    G[1] = 2.0 * oma * a;
    G[2] = a * a;
    G[3] = 0.25 * (b1a * b1a);
    G[4] = 0.5 * b1a * a1b;
    G[5] = 0.25 * (a1b * a1b);
    G[6] = b * b;
    G[7] = 2.0 * omb * b;
    G[8] = omb * omb;
}

// one step of matrix_power's chain R <- R.G (src/divergence.rs:27-29); FMA chain over k ascending
__device__ __forceinline__ void mat3_step(double R[9], const double G[9])
{
    double n[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = R[3 * i] * G[j];
            acc = __fma_rn(R[3 * i + 1], G[3 + j], acc);
            acc = __fma_rn(R[3 * i + 2], G[6 + j], acc);
            n[3 * i + j] = acc;
        }
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = n[i];
}

// src/alphabeta.rs:62-65
__device__ __forceinline__ double p_uu_est(double a, double b)
{
    double omb = 1.0 - b, oma = 1.0 - a, s = a + b, sm1 = s - 1.0;
    return (b * (omb * omb - oma * oma - 1.0)) / (s * (sm1 * sm1 - 2.0));
}
// src/structs.rs:146-149
__device__ __forceinline__ double p_mm_est(double a, double b)
{
    double omb = 1.0 - b, oma = 1.0 - a, s = a + b, sm1 = s - 1.0;
    return (a * (oma * oma - omb * omb - 1.0)) / (s * (sm1 * sm1 - 2.0));
}
// src/structs.rs:151-154
__device__ __forceinline__ double p_um_est(double a, double b)
{
    double s = a + b, sm1 = s - 1.0;
    return (4.0 * a * b * (s - 2.0)) / (s * (sm1 * sm1 - 2.0));
}

// src/divergence.rs:68-87
__device__ __forceinline__ double cond_div(const double *a, const double *b)
{
    return 0.5 * (a[0] * b[1] + a[1] * b[0] + a[1] * b[2] + a[2] * b[1]) + (a[0] * b[2] + a[2] * b[0]);
}

__device__ __forceinline__ void load_slot(const WarpCtx &c, int slot, int lane, double M[9])
{
    if (slot == 0) {  // matrix_power(.., 0) = identity (src/divergence.rs:21-24)
        M[0] = 1.0; M[1] = 0.0; M[2] = 0.0;
        M[3] = 0.0; M[4] = 1.0; M[5] = 0.0;
        M[6] = 0.0; M[7] = 0.0; M[8] = 1.0;
    } else {
        const double *p = c.pw + ((slot - 1) * 9) * 32 + lane;
#pragma unroll
        for (int e = 0; e < 9; ++e) M[e] = p[e * 32];
    }
}

// Power chain + per-triple theoretical divergence (src/divergence.rs:44-90), de-duplicated:
// G^k is built once per evaluation by the same left-associated chain matrix_power uses,
// so each dt1t2 value is bit-identical to the reference's per-pair recomputation.
__device__ __forceinline__ void model_divergence(const WarpCtx &c, int lane, double alpha, double beta,
                                                 double weight)
{
    double G[9], R[9];
    genmatrix(alpha, beta, G);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = G[i];

    int ei = 0;
    int next_e = c.n_exps > 0 ? (int)c.exps[0] : -1;
    const int tmax = c.n_exps > 0 ? (int)c.exps[c.n_exps - 1] : 0;
    for (int k = 1; k <= tmax; ++k) {
        if (k > 1) mat3_step(R, G);
        if (k == next_e) {  // warp-uniform
            double *p = c.pw + (ei * 9) * 32 + lane;
#pragma unroll
            for (int e = 0; e < 9; ++e) p[e * 32] = R[e];
            ++ei;
            next_e = ei < c.n_exps ? (int)c.exps[ei] : -1;
        }
    }

    // sv_gzero (src/divergence.rs:44)
    const double sv0 = c.p_uu0, sv1 = weight * c.p_mm0, sv2 = (1.0 - weight) * c.p_mm0;
    int cur_t0 = -1, cur_a = -1, cur_b = -1;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    double A[9], B[9];
    for (int u = 0; u < c.n_triples; ++u) {
        const uint32_t t = c.tris[u];
        const int st0 = t & 255, sa = (t >> 8) & 255, sb = (t >> 16) & 255;
        if (st0 != cur_t0) {  // all branches here are warp-uniform
            double P[9];
            load_slot(c, st0, lane, P);
            // svt0 = sv_gzero^T . G^t0 (src/divergence.rs:55)
            s0 = __fma_rn(sv2, P[6], __fma_rn(sv1, P[3], sv0 * P[0]));
            s1 = __fma_rn(sv2, P[7], __fma_rn(sv1, P[4], sv0 * P[1]));
            s2 = __fma_rn(sv2, P[8], __fma_rn(sv1, P[5], sv0 * P[2]));
            cur_t0 = st0;
        }
        if (sa != cur_a) {
            load_slot(c, sa, lane, A);
            cur_a = sa;
        }
        if (sb != cur_b) {
            load_slot(c, sb, lane, B);
            cur_b = sb;
        }
        const double d_mm = cond_div(A + 6, B + 6);
        const double d_um = cond_div(A + 3, B + 3);
        const double d_uu = cond_div(A, B);
        c.dt[u * 32 + lane] = s0 * d_uu + s1 * d_um + s2 * d_mm;  // src/divergence.rs:89
    }
}

// D access: broadcast (every lane fits the same observed column, staged in shared memory and
// read four pairs at a time) or a per-lane column (bootstrap replicates: D*[i] at col[i*32],
// lane already folded into the pointer; coalesced across the warp).
struct DBroadcast {
    const double *D;  // 16-byte aligned
    __device__ __forceinline__ void load4(int i, double d[4]) const
    {
        const double2 a = *reinterpret_cast<const double2 *>(D + i);
        const double2 b = *reinterpret_cast<const double2 *>(D + i + 2);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
    }
    __device__ __forceinline__ double operator()(int i) const { return D[i]; }
};
struct DLaneColumn {
    const double *col;
    __device__ __forceinline__ void load4(int i, double d[4]) const
    {
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = __ldcg(col + (size_t)(i + q) * 32);
    }
    __device__ __forceinline__ double operator()(int i) const { return __ldcg(col + (size_t)i * 32); }
};

// Objective (src/structs.rs:191-217).  penalty=false gives the penalty-free LSE of
// src/ab_neutral.rs:88-93 (r*r + 0.0 == r*r, so one loop serves both).
//
// The pair sum is the reference's: sequential, in pedigree order.  Only the accumulate
// (one DADD per pair) is a dependent chain; the loop is software-pipelined over groups of
// four pairs so the next group's loads / residuals / squares issue under the chain latency.
template <class DAcc>
__device__ __forceinline__ double objective(const WarpCtx &c, const DAcc &Dat, int lane, double alpha,
                                            double beta, double weight, double icpt, bool penalty)
{
    model_divergence(c, lane, alpha, beta, weight);
    double pen = 0.0;
    if (penalty) {
        const double dq = p_uu_est(alpha, beta) - c.eqp;
        pen = c.penw * (dq * dq);
    }
    // byte base of this lane's dt column; c.offs[i] = 256 * (triple id of pair i)
    const char *dtb = reinterpret_cast<const char *>(c.dt + lane);
    const int ng = c.n_pairs >> 2;
    double sum = 0.0;
    double t[4] = {0.0, 0.0, 0.0, 0.0};  // terms of the previous group (adding +0.0 first is exact)

    struct Grp {
        double d[4], raw[4];
        uint32_t off[4];
    };
    // stage 1: loads of one group of four pairs.  The first dt of a group is always fetched, the
    // others only where the (warp-uniform) triple id changes inside the group.
    auto load = [&](int g, Grp &G) {
        Dat.load4(4 * g, G.d);
        const uint4 o = *reinterpret_cast<const uint4 *>(c.offs + 4 * g);
        G.off[0] = o.x; G.off[1] = o.y; G.off[2] = o.z; G.off[3] = o.w;
        G.raw[0] = *reinterpret_cast<const double *>(dtb + o.x);
#pragma unroll
        for (int q = 1; q < 4; ++q)
            G.raw[q] = (G.off[q] != G.off[q - 1]) ? *reinterpret_cast<const double *>(dtb + G.off[q]) : 0.0;
    };
    // stage 2 + 3: residuals / squares of group G, then the sequential accumulate of the previous one
    auto step = [&](const Grp &G) {
        double v[4], n[4];
        v[0] = G.raw[0];
#pragma unroll
        for (int q = 1; q < 4; ++q) v[q] = (G.off[q] != G.off[q - 1]) ? G.raw[q] : v[q - 1];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double res = G.d[q] - icpt - v[q];
            n[q] = res * res + pen;
        }
        sum += t[0];
        sum += t[1];
        sum += t[2];
        sum += t[3];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = n[q];
    };

    if (ng > 0) {
        Grp A, B;
        load(0, A);
        int g = 0;
        for (; g + 1 < ng; g += 2) {
            load(g + 1, B);
            step(A);
            load(min(g + 2, ng - 1), A);
            step(B);
        }
        if (g < ng) step(A);
        sum += t[0];
        sum += t[1];
        sum += t[2];
        sum += t[3];
    }
    for (int i = 4 * ng; i < c.n_pairs; ++i) {
        const double res = Dat(i) - icpt - *reinterpret_cast<const double *>(dtb + c.offs[i]);
        sum += res * res + pen;
    }
    return sum;
}

// ---------------------------------------------------------------------------------
// Nelder-Mead, argmin 0.8.1 semantics (call sites src/ab_neutral.rs:49-64 and
// src/boot_model.rs:69-84), as a per-lane state machine: every trip of the warp loop
// evaluates exactly one trial point per active lane, so lanes in different NM phases
// (reflect / expand / contract / ...) still share the objective's instruction stream.
// ---------------------------------------------------------------------------------
enum Phase : int {
    PH_IDLE = 0,  // needs a new fit
    PH_INIT,      // evaluating initial vertex k
    PH_REFLECT,
    PH_EXPAND,
    PH_CONTRACT,
    PH_SHRINK,  // evaluating shrunk vertex k (sorted position)
    PH_LSE      // final penalty-free evaluation of the best vertex
};

struct LaneNM {
    double xt[4];  // trial point of the current evaluation
    double fr;     // reflection cost kept across the expansion evaluation
    int phase, k;
    uint32_t ord;  // 5 x 3 bits: physical slot of the sorted vertex at position p
    int iters, evals, status;
    int fit_id;
};

// per-lane simplex storage in shared memory: X[(slot*4+j)*32 + lane], C[slot*32 + lane]
struct LaneSimplex {
    double *X;
    double *C;
    __device__ __forceinline__ double &x(int slot, int j) const { return X[(slot * 4 + j) * 32]; }
    __device__ __forceinline__ double &c(int slot) const { return C[slot * 32]; }
};

__device__ __forceinline__ int ord_at(uint32_t ord, int p) { return (ord >> (3 * p)) & 7; }

// stable insertion sort of the five vertices by cost (sort_by(partial_cmp().unwrap_or(Equal)))
// written as a fixed compare-exchange sequence; with strict '<' it performs exactly the
// swaps the insertion sort would, NaNs included.
__device__ __forceinline__ uint32_t sort5(const LaneSimplex &S, uint32_t ord)
{
    double c[5];
    int o[5];
#pragma unroll
    for (int p = 0; p < 5; ++p) {
        o[p] = ord_at(ord, p);
        c[p] = S.c(o[p]);
    }
#pragma unroll
    for (int i = 1; i < 5; ++i)
#pragma unroll
        for (int j = i; j >= 1; --j) {
            const bool sw = c[j] < c[j - 1];
            const double tc = sw ? c[j - 1] : c[j];
            c[j - 1] = sw ? c[j] : c[j - 1];
            c[j] = tc;
            const int to = sw ? o[j - 1] : o[j];
            o[j - 1] = sw ? o[j] : o[j - 1];
            o[j] = to;
        }
    uint32_t r = 0;
#pragma unroll
    for (int p = 0; p < 5; ++p) r |= (uint32_t)o[p] << (3 * p);
    return r;
}

// x0 = (x[0]+x[1]+x[2]+x[3]) * (1/4)   (NelderMead::calculate_centroid)
__device__ __forceinline__ void centroid(const LaneSimplex &S, uint32_t ord, double x0[4])
{
    const int a = ord_at(ord, 0), b = ord_at(ord, 1), c = ord_at(ord, 2), d = ord_at(ord, 3);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double v = S.x(a, j);
        v = v + S.x(b, j);
        v = v + S.x(c, j);
        v = v + S.x(d, j);
        x0[j] = v * (1.0 / 4.0);
    }
}

// xr = x0 + (x0 - x[4]) * alpha, alpha = 1   (NelderMead::reflect)
__device__ __forceinline__ void reflect_point(const LaneSimplex &S, uint32_t ord, double xr[4])
{
    double x0[4];
    centroid(S, ord, x0);
    const int w = ord_at(ord, 4);
#pragma unroll
    for (int j = 0; j < 4; ++j) xr[j] = x0[j] + (x0[j] - S.x(w, j)) * 1.0;
}

// shrink point for sorted position k: x[0] + (x[k] - x[0]) * sigma, sigma = 0.5
__device__ __forceinline__ void shrink_point(const LaneSimplex &S, uint32_t ord, int k, double xs[4])
{
    const int b = ord_at(ord, 0), s = ord_at(ord, k);
#pragma unroll
    for (int j = 0; j < 4; ++j) xs[j] = S.x(b, j) + (S.x(s, j) - S.x(b, j)) * 0.5;
}

struct NMParams {
    int max_iters;
    double sd_tol;
    uint32_t flags;
};

// Consume the objective value f of the lane's current trial point and advance the
// state machine to the next trial point.  Returns true when the fit has finished
// (lane.phase == PH_IDLE, result in `res`).
__device__ __forceinline__ bool nm_advance(LaneNM &L, const LaneSimplex &S, const NMParams &P, double f,
                                           abfit_fit &res)
{
    bool iter_done = false;   // an NM iteration (or init) completed: sort + termination test follow
    bool finish = false;      // go to the final LSE evaluation
    switch (L.phase) {
        case PH_INIT: {
            S.c(L.k) = f;
            ++L.evals;
            if (++L.k < 5) {
#pragma unroll
                for (int j = 0; j < 4; ++j) L.xt[j] = S.x(L.k, j);
            } else {
                iter_done = true;  // sort + termination test of the Executor's first loop pass
                L.iters = -1;      // the shared tail below counts an iteration; init is not one
            }
            break;
        }
        case PH_REFLECT: {
            ++L.evals;
            const double c0 = S.c(ord_at(L.ord, 0)), c3 = S.c(ord_at(L.ord, 3));
            if (f < c3 && f >= c0) {  // Action::Reflection
                const int w = ord_at(L.ord, 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) S.x(w, j) = L.xt[j];
                S.c(w) = f;
                iter_done = true;
            } else if (f < c0) {  // Action::Expansion: xe = x0 + (xr - x0) * 2
                double x0[4];
                centroid(S, L.ord, x0);
                L.fr = f;
#pragma unroll
                for (int j = 0; j < 4; ++j) L.xt[j] = x0[j] + (L.xt[j] - x0[j]) * 2.0;
                L.phase = PH_EXPAND;
            } else if (f >= c3) {  // Action::ContractionInside: xc = x0 + (x[4] - x0) * 0.5
                double x0[4];
                centroid(S, L.ord, x0);
                const int w = ord_at(L.ord, 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) L.xt[j] = x0[j] + (S.x(w, j) - x0[j]) * 0.5;
                L.phase = PH_CONTRACT;
            } else {  // NaN reflection cost: Action::Shrink
                L.k = 1;
                shrink_point(S, L.ord, 1, L.xt);
                L.phase = PH_SHRINK;
            }
            break;
        }
        case PH_EXPAND: {
            ++L.evals;
            const int w = ord_at(L.ord, 4);
            if (f < L.fr) {
#pragma unroll
                for (int j = 0; j < 4; ++j) S.x(w, j) = L.xt[j];
                S.c(w) = f;
            } else {  // keep the reflection point (recomputed from the untouched simplex: same bits)
                double xr[4];
                reflect_point(S, L.ord, xr);
#pragma unroll
                for (int j = 0; j < 4; ++j) S.x(w, j) = xr[j];
                S.c(w) = L.fr;
            }
            iter_done = true;
            break;
        }
        case PH_CONTRACT: {
            ++L.evals;
            const int w = ord_at(L.ord, 4);
            if (f < S.c(w)) {
#pragma unroll
                for (int j = 0; j < 4; ++j) S.x(w, j) = L.xt[j];
                S.c(w) = f;
                iter_done = true;
            } else if (P.flags & ABFIT_SHRINK_ON_FAILED_CONTRACTION) {
                L.k = 1;
                shrink_point(S, L.ord, 1, L.xt);
                L.phase = PH_SHRINK;
            } else if (P.flags & ABFIT_NO_EARLY_EXIT_ON_STALL) {
                iter_done = true;  // argmin 0.8.1: nothing replaced, iteration counted
            } else {
                // Simplex unchanged and next_iter is a pure function of it: all remaining
                // iterations repeat this one.  Same best vertex as after max_iters.
                L.iters = P.max_iters;
                L.status = ABFIT_TERM_STALLED;
                finish = true;
            }
            break;
        }
        case PH_SHRINK: {
            ++L.evals;
            const int s = ord_at(L.ord, L.k);
#pragma unroll
            for (int j = 0; j < 4; ++j) S.x(s, j) = L.xt[j];
            S.c(s) = f;
            if (++L.k < 5) {
                shrink_point(S, L.ord, L.k, L.xt);
            } else {
                iter_done = true;
            }
            break;
        }
        case PH_LSE: {
            const int b = ord_at(L.ord, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) res.theta[j] = S.x(b, j);
            res.cost = S.c(b);
            res.lse = f;
            res.iters = L.iters;
            res.evals = L.evals;
            // IterState::update never accepts a NaN cost: best_param stays None and the
            // reference panics on unwrap (src/ab_neutral.rs:66)
            res.status = (res.cost != res.cost) ? ABFIT_FIT_NAN : L.status;
            res.start_id = L.fit_id;
            L.phase = PH_IDLE;
            return true;
        }
        default: break;
    }

    if (iter_done) {
        L.ord = sort5(S, L.ord);  // sort_param_vecs (stable)
        ++L.iters;
        // Executor: terminate_internal at the top of the next iteration
        double c[5];
#pragma unroll
        for (int p = 0; p < 5; ++p) c[p] = S.c(ord_at(L.ord, p));
        double sum = 0.0;
#pragma unroll
        for (int p = 0; p < 5; ++p) sum += c[p];
        const double c0 = sum / 5.0;
        double ss = 0.0;
#pragma unroll
        for (int p = 0; p < 5; ++p) {
            const double d = c[p] - c0;
            ss += d * d;
        }
        const double sd = sqrt(1.0 / (5.0 - 1.0) * ss);
        if (sd < P.sd_tol) {
            L.status = ABFIT_TERM_SD;
            finish = true;
        } else if (L.iters >= P.max_iters) {
            L.status = ABFIT_TERM_MAX_ITERS;
            finish = true;
        } else {
            reflect_point(S, L.ord, L.xt);
            L.phase = PH_REFLECT;
        }
    }
    if (finish) {
        const int b = ord_at(L.ord, 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) L.xt[j] = S.x(b, j);
        L.phase = PH_LSE;
    }
    return false;
}

// start a new fit on this lane from a 5x4 simplex in global memory
__device__ __forceinline__ void nm_begin(LaneNM &L, const LaneSimplex &S, int fit_id)
{
    L.phase = PH_INIT;
    L.k = 0;
    L.ord = 0u | (1u << 3) | (2u << 6) | (3u << 9) | (4u << 12);
    L.iters = 0;
    L.evals = 0;
    L.status = 0;
    L.fit_id = fit_id;
    L.fr = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) L.xt[j] = S.x(0, j);
}

}  // namespace abfit
