// abfit_wide.cuh — the objective with ONE WARP per fit, for pedigrees with thousands of pairs.
//
// The lane-per-fit kernels (abfit_model.cuh) keep every lane's model state on chip.  A pedigree like
// C5's (19 900 pairs, ~400 distinct triples) needs 836 doubles per lane, which only fits in global
// scratch, and a batch of 1000 starts fills 32 warps: the evaluation is then bound by the latency
// of L2 gathers.  Here the 32 lanes of a warp share one evaluation instead:
//
//   1. the power chain G, G.G, (G.G).G, ... (src/divergence.rs:27-29) is computed once (every lane
//      holds the same theta and runs the same chain redundantly) and lane 0 publishes every power
//      and every sv0^T.G^t0 (:55) in a small per-warp table in shared memory;
//   2. the distinct (t0,t1,t2) triples are spread over the lanes: conditional divergences (:68-87)
//      and dt1t2 (:89) per triple, written to a per-warp dt table;
//   3. the pair terms (D_i - c - dt_i)^2 + pen are computed 32 at a time, one pair per lane, with
//      coalesced loads of D and of the pair's triple id, and exchanged through shared memory;
//      every lane then adds the 32 terms in pedigree order (src/structs.rs:208-213).
//
// Each value is produced by the same operations in the same order as in abfit_model.cuh, so the
// two formulations agree bit for bit; only the sequential sum (one dependent DADD per pair) is a
// serial chain, and it is the floor of the evaluation time: 8 cycles x n_pairs.
#pragma once
#include "abfit_nm.cuh"

namespace abfit {

#ifdef ABFIT_WIDE_WIDTH
constexpr int WW = ABFIT_WIDE_WIDTH;  // host emulation runs the same code with a "warp" of one lane
#else
constexpr int WW = 32;
#endif
constexpr int WIDE_CH = 32;   // pairs per exchange buffer
static_assert(WIDE_CH % WW == 0, "chunk must be a multiple of the warp width");

struct WideCtx {
    const double *D;       // global: observed column (or the replicate's D* row), [n_pairs]
    const uint32_t *tid;   // global: triple id of every pair
    const uint32_t *trip;  // shared: t0 | (t1-t0) << 8 | (t2-t0) << 16 per distinct triple
    double *pw;            // per warp, shared: G^0 .. G^tmax, 9 doubles each
    double *sv;            //   sv0^T . G^t for t = 0 .. tmax, 3 doubles each
    double *dt;            //   dt1t2 per triple
    double *term;          //   three exchange buffers of WIDE_CH pair terms
    int32_t n_pairs, n_trip, tmax;
    double p_uu0, p_mm0, eqp, penw;
    // EXPERIMENT (ABFIT_EXPERIMENT_SUFFSTATS=1, never the default): per-triple sufficient statistics
    // of the observed column — ss[u] = mean of the D of triple u's pairs, ss[n_trip + u] = their number, ss[2 n_trip] =
    // the centred sum of squares over all pairs — turn step 3 from O(pairs) into O(triples).  null: the exact sum.
    const double *ss;
};

// doubles of per-warp shared memory: tables + exchange buffers + the simplex (25 x 32, LaneSimplex layout)
__host__ __device__ inline size_t wide_table_doubles(int tmax, int n_trip)
{
    return (size_t)(tmax + 1) * 12 + (((size_t)n_trip + 1) & ~(size_t)1) + 3 * WIDE_CH;
}
__host__ __device__ inline size_t wide_warp_doubles(int tmax, int n_trip)
{
    return wide_table_doubles(tmax, n_trip) + 25 * 32;
}

// Carve the per-warp tables out of `base` (16-byte aligned); returns the first double after them.
__host__ __device__ inline double *wide_carve(WideCtx &c, double *base, int tmax, int n_trip)
{
    c.pw = base;
    c.sv = c.pw + (size_t)(tmax + 1) * 9;
    c.term = c.sv + (size_t)(tmax + 1) * 3;  // (tmax + 1) * 12 doubles so far: 16-byte aligned
    c.dt = c.term + 3 * WIDE_CH;
    return c.dt + (((size_t)n_trip + 1) & ~(size_t)1);
}

__device__ __forceinline__ double objective_wide(const WideCtx &c, int lane, double alpha, double beta,
                                                 double weight, double icpt, bool penalty)
{
    // ---- 1. power chain and initial-state products ------------------------------------------
    double G[9], R[9];
    genmatrix(alpha, beta, G);
    const double sv0 = c.p_uu0, sv1 = weight * c.p_mm0, sv2 = (1.0 - weight) * c.p_mm0;  // src/divergence.rs:44
    {
        const double I[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};  // matrix_power(.., 0), :21-24
        double s[3];
        vec_mat(sv0, sv1, sv2, I, s);
        if (lane == 0) {
#pragma unroll
            for (int e = 0; e < 9; ++e) c.pw[e] = I[e];
            c.sv[0] = s[0]; c.sv[1] = s[1]; c.sv[2] = s[2];
        }
    }
#pragma unroll
    for (int e = 0; e < 9; ++e) R[e] = G[e];
    for (int k = 1; k <= c.tmax; ++k) {
        if (k > 1) mat3_step(R, G);
        double s[3];
        vec_mat(sv0, sv1, sv2, R, s);
        if (lane == 0) {
            double *p = c.pw + 9 * k;
#pragma unroll
            for (int e = 0; e < 9; ++e) p[e] = R[e];
            c.sv[3 * k] = s[0]; c.sv[3 * k + 1] = s[1]; c.sv[3 * k + 2] = s[2];
        }
    }
    __syncwarp(FULL);

    // ---- 2. dt1t2 of every distinct triple, triples spread over the lanes ------------------------
    for (int u = lane; u < c.n_trip; u += WW) {
        const uint32_t w = c.trip[u];
        const double *pa = c.pw + 9 * ((w >> 8) & 0xffu), *pb = c.pw + 9 * ((w >> 16) & 0xffu);
        double A[9], B[9], d[3];
#pragma unroll
        for (int e = 0; e < 9; ++e) {
            A[e] = pa[e];
            B[e] = pb[e];
        }
        cond_div3(A, B, d);
        const double *s = c.sv + 3 * (w & 0xffu);
        c.dt[u] = s[0] * d[0] + s[1] * d[1] + s[2] * d[2];  // src/divergence.rs:89
    }
    __syncwarp(FULL);

    // ---- 3. pair terms 32 at a time, sequential sum in pedigree order --------------------------
    double pen = 0.0;
    if (penalty) {
        const double dq = p_uu_est(alpha, beta) - c.eqp;
        pen = c.penw * (dq * dq);
    }
    if (c.ss) {
        // sum_i (D_i - c - dt_u(i))^2 = Q + sum_u n_u (m_u - c - dt_u)^2: triples over the lanes, fixed shuffle tree
        double part = 0.0;
        for (int u = lane; u < c.n_trip; u += WW) {
            const double r = c.ss[u] - icpt - c.dt[u];
            part += c.ss[c.n_trip + u] * (r * r);
        }
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(FULL, part, o);
#endif
        return c.ss[2 * c.n_trip] + part + (double)c.n_pairs * pen;
    }
    constexpr int PER = WIDE_CH / WW;  // 1 on the device
    const int n = c.n_pairs;
    // Global loads of one chunk (D, triple id), held in registers four chunks ahead of their use: D and tid
    // stream from L2 once per evaluation, and one chunk of the dependent sum lasts about 260 cycles.
    struct Slot {
        double d[PER];
        uint32_t t[PER];
    };
    auto fetch = [&](int base, Slot &s) {  // beyond the end of the pedigree: dummies, their terms are never added
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int i = base + lane + p * WW;
            const bool in = i < n;
            s.d[p] = in ? c.D[i] : 0.0;
            s.t[p] = in ? c.tid[i] : 0u;
        }
    };
    auto emit = [&](const Slot &s, double *buf) {
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const double res = s.d[p] - icpt - c.dt[s.t[p]];
            buf[lane + p * WW] = res * res + pen;
        }
    };
    auto read_terms = [&](const double *buf, double t[WIDE_CH]) {
#pragma unroll
        for (int q = 0; q < WIDE_CH; q += 2) {
            const double2 v = *reinterpret_cast<const double2 *>(buf + q);
            t[q] = v.x;
            t[q + 1] = v.y;
        }
    };
    // Main loop: groups of four full chunks, no branches inside, so that everything between two barriers is ONE
    // basic block the scheduler can interleave: the 32 dependent DADDs of chunk j (terms already in registers),
    // the shared-memory reads of chunk j + 1's terms, the residuals of chunk j + 2 and the global loads of chunk
    // j + 6.  The chain — 8 cycles per pair, the floor of the evaluation time — then runs without gaps.
    // Three exchange buffers: iteration j writes buffer (j + 2) % 3, whose previous content (chunk j - 1) was
    // read in iteration j - 2; it reads buffer (j + 1) % 3, written in iteration j - 1 and published by its barrier.
    double sum = 0.0;
    const int n_main = ((n / WIDE_CH) / 4) * 4;  // chunks that go through the pipelined loop
    if (n_main > 0) {
        double *b0 = c.term, *b1 = c.term + WIDE_CH, *b2 = c.term + 2 * WIDE_CH;  // chunks j, j + 1, j + 2
        Slot s0, s1, s2, s3;  // slot (k % 4) holds the loads of chunk k
        fetch(0, s0);
        fetch(1 * WIDE_CH, s1);
        fetch(2 * WIDE_CH, s2);
        fetch(3 * WIDE_CH, s3);
        emit(s0, b0);
        emit(s1, b1);
        fetch(4 * WIDE_CH, s0);
        fetch(5 * WIDE_CH, s1);
        __syncwarp(FULL);
        double ta[WIDE_CH], tb[WIDE_CH];
        read_terms(b0, ta);
        // adds chunk j (terms in `t`), reads chunk j + 1 into `tn`, emits chunk j + 2 from `s` and refills `s`
        auto step = [&](int j, const double t[WIDE_CH], double tn[WIDE_CH], Slot &s) {
            read_terms(b1, tn);
            emit(s, b2);
            fetch((j + 6) * WIDE_CH, s);
#pragma unroll
            for (int q = 0; q < WIDE_CH; ++q) sum += t[q];
            __syncwarp(FULL);
            double *r = b0;
            b0 = b1;
            b1 = b2;
            b2 = r;
        };
        for (int j = 0; j < n_main; j += 4) {
            step(j, ta, tb, s2);
            step(j + 1, tb, ta, s3);
            step(j + 2, ta, tb, s0);
            step(j + 3, tb, ta, s1);
        }
    }
    // Tail (fewer than four full chunks plus the partial one, < 160 pairs): every lane on its own, same operations
    for (int i = n_main * WIDE_CH; i < n; ++i) {
        const double res = c.D[i] - icpt - c.dt[c.tid[i]];
        sum += res * res + pen;
    }
    return sum;
}

}  // namespace abfit
