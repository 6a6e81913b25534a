// abfit_model.cuh — device-side ABneutral model and objective.
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
#ifdef __CUDACC_RTC__
// run-time build (abfit_jit.cu): NVRTC has the CUDA built-ins; the fixed-width integer types and the C ABI structs
// come from in-memory headers
#include "abfit_rtc_prelude.h"
#include "abfit.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/abfit.h"
#endif

namespace abfit {

constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------
// The per-window "program".  divergence() (src/divergence.rs:33-94) recomputes three
// matrix powers per pair; all of them are members of ONE left-associated chain
// G, G.G, (G.G).G, ...  The host (compile_problems, abfit_api.cu) turns a pedigree into
// a list of micro-ops that fire while the kernel walks that chain once per evaluation,
// so that only the few values that are needed again later occupy per-lane storage:
//   - G^m            where m is the smaller exponent of some (t1-t0, t2-t0) pair  (9 doubles)
//   - sv0 . G^t0     for t0 > 0                                                   (3 doubles)
//   - the three conditional divergences of an exponent pair whose t0 comes later  (3 doubles)
//   - dt1t2 of every distinct (t0,t1,t2) triple                                   (1 double)
// Every value is produced by the same operations in the same order as the reference
// produces it, so the result is bit-identical; only the redundancy is gone.
// Per-lane storage lives in shared memory as lane_mem[index * 32 + lane].
// ---------------------------------------------------------------------------------
enum MicroOp : uint32_t {
    OP_NOP = 0,      //                             nothing (carries chain steps; also the end sentinel)
    OP_STORE_M = 1,  // a = dst                     lane_mem[dst..dst+8]  <- R  (= G^k)
    OP_CALC_S = 2,   // a = dst                     lane_mem[dst..dst+2]  <- sv0^T . R          (src/divergence.rs:55)
    // dreg <- conditional divergences of rows UU,UM,MM for A = G^(t1-t0), B = G^(t2-t0) (:57-87).  The
    // op fires at k = max(t1-t0, t2-t0), so one operand is always the chain's current power R; the
    // variants name where the other one lives (no register copies, no source decoding).  Fused tail:
    //   c != NONE: lane_mem[c..c+2] <- dreg (kept for a triple whose t0 comes later)
    //   b != NONE: lane_mem[b] <- sv0 . dreg  (the t0 == 0 triple of this exponent pair, :89)
    OP_D_CC = 3,     //                             A = R,            B = R
    OP_D_MC = 4,     // a = src                     A = lane_mem[src], B = R
    OP_D_CM = 5,     // a = src                     A = R,            B = lane_mem[src]
    OP_D_GC = 6,     //                             A = G,            B = R
    OP_D_CG = 7,     //                             A = R,            B = G
    OP_D_GEN = 8,    // a = A source, b = B source  identity operands (t1 == t0 or t2 == t0); no fused tail
    OP_STORE_D = 9,  // a = dst                     lane_mem[dst..dst+2]  <- dreg          (after OP_D_GEN)
    OP_DT0 = 10,     // a = dst                     lane_mem[dst] <- sv0 . dreg            (after OP_D_GEN)
    OP_DT = 11,      // a = dst, b = s source       lane_mem[dst] <- s0*d_uu + s1*d_um + s2*d_mm (:89)
    OP_LDT = 12      // a = dst, b = s source, c = d source   dreg <- lane_mem[c..c+2], then OP_DT
};
// operand sources (16 bit); anything below SRC_SPECIAL is a lane_mem index
constexpr uint32_t SRC_SPECIAL = 0xfff0;
constexpr uint32_t SRC_CUR = 0xfff0;    // the chain's current power R = G^k
constexpr uint32_t SRC_G = 0xfff1;      // G itself (exponent 1), always in registers
constexpr uint32_t SRC_IDENT = 0xfff2;  // G^0 (src/divergence.rs:21-24)
constexpr uint32_t OP_NONE = 0xffff;    // absent b / c field
// op word: x = opcode | n_step << 8 | a << 16,  y = b | c << 16.
// n_step: chain steps R <- R.G (matrix_power's loop, :27-29) executed BEFORE the op.
struct OpWord {
    uint32_t x, y;
};

struct DevProblem {
    int64_t d_off;     // into the D pool (even: 16-byte aligned)
    int64_t pair_off;  // into pred / resid (problems concatenated without padding)
    int64_t offs_off;  // into the pair-offset pool: u32 = 256 * lane_mem index of the pair's dt1t2
    int64_t ops_off;   // into the micro-op pool
    int32_t n_pairs, n_offs;  // n_offs = n_pairs rounded up to 4
    int32_t n_ops;
    int32_t n_lane;   // doubles of per-lane model storage
    int32_t tmax;     // last exponent the chain has to reach
    int32_t n_trip;   // distinct (t0,t1,t2) triples
    int64_t wtrip_off;  // into the triple pool (u32 per triple: t0 | (t1-t0) << 8 | (t2-t0) << 16), see abfit_wide.cuh
    int64_t wtid_off;   // into the per-pair triple-id pool
    double p_uu0, p_mm0;  // state at G0 (p0um = 0), src/ab_neutral.rs:23-24
    double eqp, penw;     // penw = eqp_weight * (double)n_pairs, src/structs.rs:210-211
};

// chunk of consecutive starts / replicates / thetas of one problem, processed by one block
struct WorkItem {
    int32_t prob;
    int32_t first;
    int32_t count;
    int32_t pad;
};

struct DevicePools {  // device pointers of a compiled batch
    const DevProblem *probs;
    const double *D;
    const uint32_t *offs;
    const OpWord *ops;
    const uint32_t *wtrip;
    const uint32_t *wtid;
};

// per-warp view of the staged problem
struct WarpCtx {
    const double *D;       // [n_pairs] shared (or global for very long pedigrees), 16-byte aligned
    const uint32_t *offs;  // shared, 16-byte aligned
    const OpWord *ops;     // shared
    double *lm;            // this warp's lane_mem (lane NOT folded in)
    int32_t n_pairs, n_ops;
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

// the same step into a second matrix, N = R.G: lets a rolled loop ping-pong between two register sets instead of
// copying nine values per step (specialised objectives, abfit_jit.cu)
__device__ __forceinline__ void mat3_to(const double R[9], const double G[9], double N[9])
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double acc = R[3 * i] * G[j];
            acc = __fma_rn(R[3 * i + 1], G[3 + j], acc);
            acc = __fma_rn(R[3 * i + 2], G[6 + j], acc);
            N[3 * i + j] = acc;
        }
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

// sv^T . M with the same FMA chain as the 3x3 products
__device__ __forceinline__ void vec_mat(double v0, double v1, double v2, const double M[9], double s[3])
{
    s[0] = __fma_rn(v2, M[6], __fma_rn(v1, M[3], v0 * M[0]));
    s[1] = __fma_rn(v2, M[7], __fma_rn(v1, M[4], v0 * M[1]));
    s[2] = __fma_rn(v2, M[8], __fma_rn(v1, M[5], v0 * M[2]));
}

__device__ __forceinline__ void fetch_matrix(uint32_t src, const double *lml, const double R[9], const double G[9],
                                             double M[9])
{
    if (src == SRC_CUR) {
#pragma unroll
        for (int e = 0; e < 9; ++e) M[e] = R[e];
    } else if (src == SRC_G) {
#pragma unroll
        for (int e = 0; e < 9; ++e) M[e] = G[e];
    } else if (src == SRC_IDENT) {
        M[0] = 1.0; M[1] = 0.0; M[2] = 0.0;
        M[3] = 0.0; M[4] = 1.0; M[5] = 0.0;
        M[6] = 0.0; M[7] = 0.0; M[8] = 1.0;
    } else {
        const double *p = lml + src * 32;
#pragma unroll
        for (int e = 0; e < 9; ++e) M[e] = p[e * 32];
    }
}

// conditional divergences of the three rows, in the reference's order MM, UM, UU (:68-87)
__device__ __forceinline__ void cond_div3(const double A[9], const double B[9], double dreg[3])
{
    dreg[2] = cond_div(A + 6, B + 6);
    dreg[1] = cond_div(A + 3, B + 3);
    dreg[0] = cond_div(A, B);
}

__device__ __forceinline__ void load_matrix(const double *lml, uint32_t src, double M[9])
{
    const double *p = lml + src * 32;
#pragma unroll
    for (int e = 0; e < 9; ++e) M[e] = p[e * 32];
}

// Power chain + per-triple theoretical divergence (src/divergence.rs:44-90): ONE flat loop over the
// window's micro-ops.  The chain steps (OP_STEP) live in the same loop as the ops that read R on
// purpose: with R loop-variant the compiler cannot hoist the pure cond_div3 variants out of the loop
// and execute all of them speculatively (it did, for ~1000 FP64 instructions per evaluation, when
// the ops of one exponent formed an inner loop).  All control flow here is warp-uniform.
__device__ __forceinline__ void model_divergence(const WarpCtx &c, int lane, double alpha, double beta,
                                                 double weight)
{
    double *lml = c.lm + lane;
    double G[9], R[9], dreg[3] = {0.0, 0.0, 0.0};
    genmatrix(alpha, beta, G);
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = G[i];
    // sv_gzero (src/divergence.rs:44) and its product with G^0
    const double sv0 = c.p_uu0, sv1 = weight * c.p_mm0, sv2 = (1.0 - weight) * c.p_mm0;
    const double I[9] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0};
    double sid[3];
    vec_mat(sv0, sv1, sv2, I, sid);

    // The program ends with an OP_NOP sentinel, so the next word can always be fetched while the
    // current op executes (the decode no longer waits on its own shared-memory load).
    OpWord nxt = c.ops[0];
    for (int i = 1; i < c.n_ops; ++i) {
        const OpWord w = nxt;
        nxt = c.ops[i];
        const uint32_t op = w.x & 0xff, a = w.x >> 16, b = w.y & 0xffff, cc = w.y >> 16;
        for (uint32_t s = (w.x >> 8) & 0xff; s > 0; --s) mat3_step(R, G);
        bool d_tail = false;
        switch (op) {
            case OP_STORE_M: {
                double *p = lml + a * 32;
#pragma unroll
                for (int e = 0; e < 9; ++e) p[e * 32] = R[e];
                break;
            }
            case OP_CALC_S: {
                double s[3];
                vec_mat(sv0, sv1, sv2, R, s);
                double *p = lml + a * 32;
                p[0] = s[0]; p[32] = s[1]; p[64] = s[2];
                break;
            }
            case OP_D_CC: cond_div3(R, R, dreg); d_tail = true; break;
            case OP_D_MC: {
                double M[9];
                load_matrix(lml, a, M);
                cond_div3(M, R, dreg);
                d_tail = true;
                break;
            }
            case OP_D_CM: {
                double M[9];
                load_matrix(lml, a, M);
                cond_div3(R, M, dreg);
                d_tail = true;
                break;
            }
            case OP_D_GC: cond_div3(G, R, dreg); d_tail = true; break;
            case OP_D_CG: cond_div3(R, G, dreg); d_tail = true; break;
            case OP_D_GEN: {  // identity operands (t1 == t0 or t2 == t0) and anything unusual
                double A[9], B[9];
                fetch_matrix(a, lml, R, G, A);
                fetch_matrix(b, lml, R, G, B);
                cond_div3(A, B, dreg);
                break;
            }
            case OP_STORE_D: {
                double *p = lml + a * 32;
                p[0] = dreg[0]; p[32] = dreg[1]; p[64] = dreg[2];
                break;
            }
            case OP_DT0:
                lml[a * 32] = sid[0] * dreg[0] + sid[1] * dreg[1] + sid[2] * dreg[2];  // src/divergence.rs:89
                break;
            case OP_LDT: {
                const double *p = lml + cc * 32;
                dreg[0] = p[0]; dreg[1] = p[32]; dreg[2] = p[64];
            }
            // fall through
            case OP_DT: {
                const double *p = lml + b * 32;
                const double s0 = p[0], s1 = p[32], s2 = p[64];
                lml[a * 32] = s0 * dreg[0] + s1 * dreg[1] + s2 * dreg[2];
                break;
            }
            default: break;
        }
        if (d_tail) {
            if (cc != OP_NONE) {
                double *p = lml + cc * 32;
                p[0] = dreg[0]; p[32] = dreg[1]; p[64] = dreg[2];
            }
            if (b != OP_NONE) lml[b * 32] = sid[0] * dreg[0] + sid[1] * dreg[1] + sid[2] * dreg[2];  // :89
        }
    }
}

// D access: broadcast (every lane fits the same observed column, staged in shared memory and
// read four pairs at a time) or a per-lane column (bootstrap replicates: D*[i] at col[i*32],
// lane already folded into the pointer; coalesced across the warp).
// (ROLL: a generated objective may roll its pair loop per accessor type, see abfit_jit.cu)
#ifndef ABFIT_ROLL_FIT
#define ABFIT_ROLL_FIT 0
#endif
#ifndef ABFIT_ROLL_BOOT
#define ABFIT_ROLL_BOOT 0
#endif
#ifndef ABFIT_SUFF_FIT
#define ABFIT_SUFF_FIT 0
#endif
#ifndef ABFIT_GATHER_PREFETCH
#define ABFIT_GATHER_PREFETCH 4  // groups of four pairs the bootstrap's index tile is prefetched ahead (L2 -> L1)
#endif
struct DBroadcast {
    static constexpr bool ROLL = ABFIT_ROLL_FIT != 0;
    static constexpr bool SUFF = ABFIT_SUFF_FIT != 0;  // EXPERIMENT: the slot holds per-triple statistics instead of D
    const double *D;  // 16-byte aligned
    __device__ __forceinline__ void load4(int i, double d[4]) const
    {
        const double2 a = *reinterpret_cast<const double2 *>(D + i);
        const double2 b = *reinterpret_cast<const double2 *>(D + i + 2);
        d[0] = a.x; d[1] = a.y; d[2] = b.x; d[3] = b.y;
    }
    __device__ __forceinline__ double operator()(int i) const { return D[i]; }
    __device__ __forceinline__ double stat(int u) const { return D[u]; }  // SUFF: the slot holds [mean_u][Q]
};
struct DLaneColumn {
    static constexpr bool ROLL = false;
    static constexpr bool SUFF = false;
    const double *col;
    __device__ __forceinline__ void load4(int i, double d[4]) const
    {
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = __ldcg(col + (size_t)(i + q) * 32);
    }
    __device__ __forceinline__ double operator()(int i) const { return __ldcg(col + (size_t)i * 32); }
    __device__ __forceinline__ double stat(int) const { return 0.0; }
};

// Bootstrap replicate (src/boot_model.rs:43-57): D*_i = pred_i + resid[idx_i], rebuilt on the fly from
// the replicate's resample indices.  Only the indices are per-lane data: four u16 per lane and group of
// four pairs in one coalesced 8-byte load from an L2-resident tile (a quarter of the bytes of a stored D*
// column); pred is a shared-memory broadcast and resid a shared-memory gather.
#ifndef ABFIT_SUFF_BOOT
#define ABFIT_SUFF_BOOT 0
#endif
struct DGather {
    static constexpr bool ROLL = ABFIT_ROLL_BOOT != 0;
    static constexpr bool SUFF = ABFIT_SUFF_BOOT != 0;  // EXPERIMENT: `st` holds the replicate's per-triple statistics
    const uint2 *tile;    // [group][32] (lane folded in): four u16 byte offsets into resid
    const double *pred;   // shared, 16-byte aligned
    const char *resid;    // shared, at the start of dynamic shared memory
    const double *st;     // EXPERIMENT (ABFIT_EXPERIMENT_SUFFSTATS): this lane's [mean_u][Q] of the replicate, else null
    __device__ __forceinline__ void load4(int i, double d[4]) const
    {
        const uint2 *p = tile + (size_t)(i >> 2) * 32;
#ifdef __CUDA_ARCH__
        // the tile streams from L2 once per evaluation: pull the line 4 groups ahead into L1
        asm volatile("prefetch.global.L1 [%0];" ::"l"(p + ABFIT_GATHER_PREFETCH * 32));
#endif
        const uint2 w = *p;
        const double2 a = *reinterpret_cast<const double2 *>(pred + i);
        const double2 b = *reinterpret_cast<const double2 *>(pred + i + 2);
        d[0] = a.x + *reinterpret_cast<const double *>(resid + (w.x & 0xffffu));
        d[1] = a.y + *reinterpret_cast<const double *>(resid + (w.x >> 16));
        d[2] = b.x + *reinterpret_cast<const double *>(resid + (w.y & 0xffffu));
        d[3] = b.y + *reinterpret_cast<const double *>(resid + (w.y >> 16));
    }
    __device__ __forceinline__ double operator()(int i) const
    {
        const uint2 w = tile[(size_t)(i >> 2) * 32];
        const uint32_t h = (i & 2) ? w.y : w.x;
        return pred[i] + *reinterpret_cast<const double *>(resid + ((i & 1) ? (h >> 16) : (h & 0xffffu)));
    }
    __device__ __forceinline__ double stat(int u) const { return st[u]; }
};

// Objective (src/structs.rs:191-217).  penalty=false gives the penalty-free LSE of
// src/ab_neutral.rs:88-93 (r*r + 0.0 == r*r, so one loop serves both).
//
// The pair sum is the reference's: sequential, in pedigree order.  Only the accumulate
// (one DADD per pair) is a dependent chain; the loop is software-pipelined over groups of
// four pairs (load -> residual/square -> accumulate) so the next groups' loads and
// arithmetic issue under the chain latency.
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
    // byte base of this lane's storage; c.offs[i] = 256 * (lane_mem index of the pair's dt1t2)
    const char *dtb = reinterpret_cast<const char *>(c.lm + lane);
    const int ng = c.n_pairs >> 2;
    double sum = 0.0;
    double t[4] = {0.0, 0.0, 0.0, 0.0};  // terms of the previous group (adding +0.0 first is exact)

    struct Grp {
        double d[4], v[4];
    };
    // stage 1: loads of one group of four pairs (D, the four dt offsets, the four dt values).  All
    // loads are unconditional: comparing offsets to skip repeated triples costs more issue slots
    // than the shared-memory wavefronts it saves (ncu: the kernel is issue-bound, not LSU-bound).
    auto load = [&](int g, Grp &G) {
        Dat.load4(4 * g, G.d);
        const uint4 o = *reinterpret_cast<const uint4 *>(c.offs + 4 * g);
        G.v[0] = *reinterpret_cast<const double *>(dtb + o.x);
        G.v[1] = *reinterpret_cast<const double *>(dtb + o.y);
        G.v[2] = *reinterpret_cast<const double *>(dtb + o.z);
        G.v[3] = *reinterpret_cast<const double *>(dtb + o.w);
    };
    // stage 2 + 3: residuals / squares of group G, then the sequential accumulate of the previous one
    auto step = [&](const Grp &G) {
        double n[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double res = G.d[q] - icpt - G.v[q];
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

}  // namespace abfit
