// abfit_kernels.cu — sm_100a kernels of the ABneutral fit path.
//
//   k_fit_starts   multi-start Nelder-Mead      (src/ab_neutral.rs:37-78)
//   k_select       best-of-starts + pred/resid  (src/ab_neutral.rs:83-135)
//   k_fit_boot     bootstrap refits             (src/boot_model.rs:41-100)
//   k_cost_batch   objective only               (src/structs.rs:191-217)
//   k_model_div    dt1t2 per pair               (src/divergence.rs:33-94)
//   k_fp64_peak    DFMA roofline micro-benchmark
//
// Grid shape: one block per work item; an item is a chunk of consecutive starts /
// replicates of ONE window, so all warps of the block share the staged pedigree (D column,
// per-pair offsets, micro-op program) and every warp runs a perfectly uniform objective on 32
// different thetas.  Lanes that finish a fit pull the next start of the chunk from a
// block-wide counter (one shared-memory atomic per warp refill), which takes the 2-4x
// spread in Nelder-Mead iteration counts off the warp's critical path.  The kernel is bound by
// FP64 issue latency, so the per-fit on-chip footprint (shared bytes per lane) is what
// buys throughput: see DESIGN.md for the accounting.
#include <cstdio>
#include <cstdlib>

#include "abfit_internal.h"
#include "abfit_wide.cuh"

namespace abfit {

// ---------------------------------------------------------------------------------
// shared-memory carve-up
//   [ per warp: (n_lane + simplex_doubles) x 32 doubles ] [ D ] [ offs ] [ ops ] [ queue ]
// simplex_doubles: 0 = no Nelder-Mead state, 25 = vertices X (20) + costs C (5) in shared memory,
// 5 = only the costs in shared memory, the vertices in a global (L2-resident) scratch area — they are
// touched once per evaluation, and giving up their 5 KB per warp is what lets 16 warps share an SM.
// ---------------------------------------------------------------------------------
struct Carved {
    WarpCtx ctx;
    LaneSimplex simplex;
    int *queue;
};

template <bool D_SHARED>
__device__ __forceinline__ Carved carve_and_stage(const DevProblem &pb, const DevicePools &P, int simplex_doubles,
                                                  double *x_scratch = nullptr, int lead_doubles = 0)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, n_warps = nthr >> 5;
    const int per_warp = (pb.n_lane + simplex_doubles) * 32;
    Carved cv;
    double *mine = smem + lead_doubles + (size_t)warp * per_warp;  // lead_doubles: multiple of 32
    cv.ctx.lm = mine;
    if (simplex_doubles == 25) {
        cv.simplex.X = mine + pb.n_lane * 32 + lane;
        cv.simplex.C = cv.simplex.X + 20 * 32;
    } else if (simplex_doubles == 5) {
        cv.simplex.X = x_scratch + ((size_t)blockIdx.x * n_warps + warp) * (20 * 32) + lane;
        cv.simplex.C = mine + pb.n_lane * 32 + lane;
    } else {
        cv.simplex.X = nullptr;
        cv.simplex.C = nullptr;
    }
    double *p = smem + lead_doubles + (size_t)n_warps * per_warp;  // multiple of 256 bytes
    const double *Dg = P.D + pb.d_off;
    if (D_SHARED) {
        double *Ds = p;
        p += (pb.n_pairs + 1) & ~1;
        for (int i = tid; i < pb.n_pairs; i += nthr) Ds[i] = Dg[i];
        cv.ctx.D = Ds;
    } else {
        cv.ctx.D = Dg;  // pool offsets are even: 16-byte aligned
    }
    uint32_t *offs = reinterpret_cast<uint32_t *>(p);  // 16-byte aligned; n_offs is a multiple of 4
    OpWord *ops = reinterpret_cast<OpWord *>(offs + pb.n_offs);
    cv.queue = reinterpret_cast<int *>(ops + pb.n_ops);
    for (int i = tid; i < pb.n_offs; i += nthr) offs[i] = P.offs[pb.offs_off + i];
    for (int i = tid; i < pb.n_ops; i += nthr) ops[i] = P.ops[pb.ops_off + i];
    cv.ctx.offs = offs;
    cv.ctx.ops = ops;
    cv.ctx.n_pairs = pb.n_pairs;
    cv.ctx.n_ops = pb.n_ops;
    cv.ctx.p_uu0 = pb.p_uu0;
    cv.ctx.p_mm0 = pb.p_mm0;
    cv.ctx.eqp = pb.eqp;
    cv.ctx.penw = pb.penw;
    return cv;
}

// Pedigrees whose per-lane model state does not fit in shared memory (hundreds of distinct triples: C5's
// 19 900-pair pedigree needs 836 doubles per lane): lane state, simplex vertices, D and the pair offsets all
// live in global memory (an L2-resident scratch, [index][32 lanes] so every access is one coalesced 256-byte
// request); only the simplex costs, the program and the queue stay in shared memory.  Same code, same bits —
// just slower per evaluation.
__device__ __forceinline__ Carved carve_big(const DevProblem &pb, const DevicePools &P, bool with_nm,
                                            double *x_scratch, double *lm_scratch, size_t lm_stride)
{
    extern __shared__ double smem[];
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int warp = tid >> 5, lane = tid & 31, n_warps = nthr >> 5;
    const size_t slot = (size_t)blockIdx.x * n_warps + warp;
    Carved cv;
    cv.ctx.lm = lm_scratch + slot * lm_stride;
    if (with_nm) {
        cv.simplex.X = x_scratch + slot * (20 * 32) + lane;
        cv.simplex.C = smem + (size_t)warp * (5 * 32) + lane;
    } else {
        cv.simplex.X = nullptr;
        cv.simplex.C = nullptr;
    }
    OpWord *ops = reinterpret_cast<OpWord *>(smem + (with_nm ? (size_t)n_warps * (5 * 32) : 0));
    cv.queue = reinterpret_cast<int *>(ops + pb.n_ops);
    for (int i = tid; i < pb.n_ops; i += nthr) ops[i] = P.ops[pb.ops_off + i];
    cv.ctx.D = P.D + pb.d_off;
    cv.ctx.offs = P.offs + pb.offs_off;
    cv.ctx.ops = ops;
    cv.ctx.n_pairs = pb.n_pairs;
    cv.ctx.n_ops = pb.n_ops;
    cv.ctx.p_uu0 = pb.p_uu0;
    cv.ctx.p_mm0 = pb.p_mm0;
    cv.ctx.eqp = pb.eqp;
    cv.ctx.penw = pb.penw;
    return cv;
}

__device__ __forceinline__ void store_fit(abfit_fit *dst, const abfit_fit &r)
{
    // 64-byte record written as four 16-byte vector stores (dst is 64-byte aligned)
    double2 *d = reinterpret_cast<double2 *>(dst);
    d[0] = make_double2(r.theta[0], r.theta[1]);
    d[1] = make_double2(r.theta[2], r.theta[3]);
    d[2] = make_double2(r.cost, r.lse);
    d[3] = make_double2(__hiloint2double(r.evals, r.iters), __hiloint2double(r.start_id, r.status));
}

__device__ __forceinline__ void lane_nm_reset(LaneNM &L)
{
    L.phase = PH_IDLE;
    L.k = 0; L.ord = 0; L.iters = 0; L.evals = 0; L.status = 0; L.fit_id = -1; L.fr = 0.0;
    L.xt[0] = L.xt[1] = L.xt[2] = L.xt[3] = 0.0;
}

// take `count` consecutive fit ids for the idle lanes in mask m; returns this lane's id (or >= end)
__device__ __forceinline__ int queue_take(int *queue, unsigned m, int lane, int end, bool &drained)
{
    int base = end;
    if (!drained) {
        if (lane == 0) base = atomicAdd(queue, __popc(m));
        base = __shfl_sync(FULL, base, 0);
        if (base >= end) drained = true;  // warp-uniform
    }
    return base + __popc(m & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------------------------
// Tail hand-off.  Once a block's queue is drained its warps thin out: Nelder-Mead run lengths spread 4x, so
// a warp keeps executing full-width instructions for a handful of long fits (simulation on the C4 shape:
// 89 % of the issued lanes do useful work with 3 warps per block, 95 % with hand-off).  A warp with few
// active lanes therefore hands them to a sibling warp of the same block that has room, and exits.
//
// Mailbox protocol (one shared int `mb`, one published active-lane count per warp; all offers are
// targeted, every take goes through a CAS, only the donor withdraws):
//   donor     drained, 0 < active <= HANDOFF_MAX, mb == 0, a live sibling r with nact[r] + active <= 32:
//             CAS mb 0 -> OFFER(donor, r, count); park the lane states in its own (now idle) lane_mem;
//             fence; wait: mb == 0 -> taken, exit;  nact[r] < 0 (r has exited) -> CAS OFFER -> 0, resume.
//   receiver  sees OFFER targeted at it (every loop trip, and once more after announcing its exit):
//             CAS OFFER -> TAKING; idle lanes load the parked states and copy the simplices; mb = 0.
// Active-lane counts only fall after the drain, so the room the donor saw is still there when the receiver
// looks.  The moved state is the complete LaneNM + simplex, so results are unchanged bit for bit.
// ---------------------------------------------------------------------------------
constexpr int HANDOFF_MAX = 16;
constexpr int MB_OFFER = 1 << 30, MB_TAKING = 1 << 29;
__device__ __forceinline__ int mb_offer(int donor, int target, int count) { return MB_OFFER | donor | (target << 4) | (count << 8); }

struct HandoffCtx {
    volatile int *mb;            // mailbox word
    volatile signed char *nact;  // [n_warps] published active-lane counts (-1: exited)
    double *lm_base;             // shared: start of warp 0's per-warp region
    int per_warp;                // doubles per warp region
    int n_lane;                  // doubles of lane state (simplex X follows)
};

__device__ __forceinline__ void handoff_park(const LaneNM &L, double *ent, int lane)
{
    ent[0] = L.xt[0]; ent[1] = L.xt[1]; ent[2] = L.xt[2]; ent[3] = L.xt[3];
    ent[4] = L.fr;
    ent[5] = __hiloint2double(L.phase, L.k);
    ent[6] = __hiloint2double((int)L.ord, L.iters);
    ent[7] = __hiloint2double(L.evals, L.status);
    ent[8] = __hiloint2double(L.fit_id, lane);
}
__device__ __forceinline__ int handoff_unpark(LaneNM &L, const double *ent)
{
    L.xt[0] = ent[0]; L.xt[1] = ent[1]; L.xt[2] = ent[2]; L.xt[3] = ent[3];
    L.fr = ent[4];
    L.phase = __double2hiint(ent[5]); L.k = __double2loint(ent[5]);
    L.ord = (uint32_t)__double2hiint(ent[6]); L.iters = __double2loint(ent[6]);
    L.evals = __double2hiint(ent[7]); L.status = __double2loint(ent[7]);
    L.fit_id = __double2hiint(ent[8]);
    return __double2loint(ent[8]);  // the donor lane that owns the simplex
}

// receiver side: returns true when lanes were taken over
__device__ __forceinline__ bool handoff_try_take(const HandoffCtx &H, int warp, int lane, LaneNM &L, const LaneSimplex &S)
{
    const unsigned idle = __ballot_sync(FULL, L.phase == PH_IDLE);
    int v = 0;
    if (lane == 0) {
        v = *H.mb;
        if (!((v & MB_OFFER) && ((v >> 4) & 15) == warp && ((v >> 8) & 63) <= __popc(idle) &&
              atomicCAS(const_cast<int *>(H.mb), v, MB_TAKING) == v))
            v = 0;
    }
    v = __shfl_sync(FULL, v, 0);
    if (!v) return false;
    __threadfence_block();
    const int donor = v & 15, count = (v >> 8) & 63;
    const int rank = __popc(idle & ((1u << lane) - 1u));
    if (L.phase == PH_IDLE && rank < count) {
        const double *dbase = H.lm_base + (size_t)donor * H.per_warp;
        const int dl = handoff_unpark(L, dbase + rank * 16);
        const double *dX = dbase + H.n_lane * 32 + dl;  // donor lane's simplex: X[20], C[5] at stride 32
#pragma unroll
        for (int q = 0; q < 20; ++q) S.X[q * 32] = dX[q * 32];
#pragma unroll
        for (int q = 0; q < 5; ++q) S.C[q * 32] = dX[(20 + q) * 32];
    }
    const unsigned now = __ballot_sync(FULL, L.phase != PH_IDLE);
    if (lane == 0) {
        H.nact[warp] = (signed char)__popc(now);  // published before the mailbox is released: the next donor sees it
        __threadfence_block();
        *H.mb = 0;
    }
    return true;
}

// donor side: returns true when the lanes were taken (the warp is empty now), false when it keeps them
__device__ __forceinline__ bool handoff_try_give(const HandoffCtx &H, int warp, int n_warps, int lane, LaneNM &L,
                                                 unsigned amask, double *my_lm)
{
    const int n_act = __popc(amask);
    int offer = 0;
    if (lane == 0 && *H.mb == 0) {
        int target = -1, best = 0;
        for (int r = 0; r < n_warps; ++r) {
            const int a = H.nact[r];
            if (r != warp && a > best && a + n_act <= 32) {
                best = a;
                target = r;
            }
        }
        if (target >= 0) {
            offer = mb_offer(warp, target, n_act);
            if (atomicCAS(const_cast<int *>(H.mb), 0, MB_TAKING) != 0) offer = 0;  // reserved while the states are parked
            else H.nact[warp] = 0;  // not a target for anybody while it is giving its lanes away
        }
    }
    offer = __shfl_sync(FULL, offer, 0);
    if (!offer) return false;
    if (L.phase != PH_IDLE) handoff_park(L, my_lm + __popc(amask & ((1u << lane) - 1u)) * 16, lane);
    __syncwarp();
    __threadfence_block();
    const int target = (offer >> 4) & 15;
    int taken = 0;  // 1 taken, 2 withdrawn
    if (lane == 0) {
        atomicExch(const_cast<int *>(H.mb), offer);
        for (;;) {
            const int v = *H.mb;
            if (v == 0) {
                taken = 1;
                break;
            }
            if (v == offer && H.nact[target] < 0 && atomicCAS(const_cast<int *>(H.mb), offer, 0) == offer) {
                taken = 2;
                break;
            }
            __nanosleep(256);
        }
    }
    taken = __shfl_sync(FULL, taken, 0);
    if (taken == 1) {
        L.phase = PH_IDLE;  // the fits live on in the receiver
        return true;
    }
    return false;  // withdrawn: registers still hold the states
}

// ---------------------------------------------------------------------------------
// multi-start Nelder-Mead
// ---------------------------------------------------------------------------------
template <bool D_SHARED, bool X_GLOBAL, bool BIG>
__global__ void __launch_bounds__(128, X_GLOBAL ? 4 : 3)
k_fit_starts(DevicePools P, const WorkItem *__restrict__ items, const double *__restrict__ simplices,
             int n_starts, NMParams nm, abfit_fit *__restrict__ all_out,
             unsigned long long *__restrict__ evals_per_prob, double *x_scratch, double *lm_scratch,
             size_t lm_stride)
{
    constexpr bool HANDOFF = !X_GLOBAL && !BIG;  // the simplex of a moved lane is copied between shared regions
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const WorkItem it = items[blockIdx.x];
    const DevProblem pb = P.probs[it.prob];
    Carved cv = BIG ? carve_big(pb, P, true, x_scratch, lm_scratch, lm_stride)
                    : carve_and_stage<D_SHARED>(pb, P, X_GLOBAL ? 5 : 25, x_scratch);
    if (threadIdx.x == 0) {
        cv.queue[0] = it.first;
        cv.queue[1] = 0;           // mailbox
        cv.queue[2] = 0x20202020;  // published active-lane counts: everybody full
    }
    __syncthreads();
    const WarpCtx &c = cv.ctx;
    const LaneSimplex &S = cv.simplex;
    const DBroadcast Dat{c.D};
    HandoffCtx H;
    H.mb = cv.queue + 1;
    H.nact = reinterpret_cast<volatile signed char *>(cv.queue + 2);
    H.per_warp = (pb.n_lane + 25) * 32;
    H.lm_base = c.lm - (size_t)warp * H.per_warp;
    H.n_lane = pb.n_lane;

    LaneNM L;
    lane_nm_reset(L);
    const int end = it.first + it.count;
    bool drained = false;
    unsigned long long my_evals = 0;

    for (;;) {
        // ---- refill idle lanes from the block's chunk ----
        const bool need = (L.phase == PH_IDLE);
        const unsigned m = __ballot_sync(FULL, need);
        if (m && !drained) {
            const int idx = queue_take(cv.queue, m, lane, end, drained);
            if (need && idx < end) {
                const double *sx = simplices + ((size_t)it.prob * n_starts + idx) * 20;
#pragma unroll
                for (int q = 0; q < 20; ++q) S.X[q * 32] = sx[q];
                nm_begin(L, S, idx);
            }
        }
        unsigned amask = __ballot_sync(FULL, L.phase != PH_IDLE);
        if (HANDOFF && drained && n_warps > 1 && pb.n_lane * 32 >= HANDOFF_MAX * 16) {
            if (lane == 0) H.nact[warp] = (signed char)__popc(amask);
            if (amask != FULL && handoff_try_take(H, warp, lane, L, S)) {
                amask = __ballot_sync(FULL, L.phase != PH_IDLE);
            } else if (amask && __popc(amask) <= HANDOFF_MAX && handoff_try_give(H, warp, n_warps, lane, L, amask, c.lm)) {
                amask = 0;
            }
        }
        if (!amask) {
            if (HANDOFF && n_warps > 1 && pb.n_lane * 32 >= HANDOFF_MAX * 16) {
                // announce the exit, then look once more: an offer posted in between is either taken here or
                // withdrawn by its donor (whoever wins the CAS)
                if (lane == 0) H.nact[warp] = -1;
                __threadfence_block();
                if (handoff_try_take(H, warp, lane, L, S)) continue;
            }
            break;
        }
        const bool active = (L.phase != PH_IDLE);
        if (active) {
            const double f =
                objective(c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            abfit_fit res;
            if (nm_advance(L, S, nm, f, res, amask)) {
                my_evals += (unsigned long long)res.evals;
                store_fit(all_out + (size_t)it.prob * n_starts + res.start_id, res);
            }
        }
    }
    // FLOP accounting: objective evaluations actually executed (excludes the LSE pass)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(FULL, my_evals, o);
    if (lane == 0 && evals_per_prob) atomicAdd(evals_per_prob + it.prob, my_evals);
}

// ---------------------------------------------------------------------------------
// best-of-starts (warp-shuffle argmin) + predicted divergence / residuals of the best
// ---------------------------------------------------------------------------------
template <bool D_SHARED, bool BIG>
__global__ void __launch_bounds__(32)
k_select(DevicePools P, int n_starts, const abfit_fit *__restrict__ all, abfit_fit *__restrict__ best_out,
         double *__restrict__ pred, double *__restrict__ resid, int32_t *__restrict__ prob_status,
         double *lm_scratch, size_t lm_stride)
{
    const int lane = threadIdx.x;
    const int p = blockIdx.x;
    const DevProblem pb = P.probs[p];
    Carved cv = BIG ? carve_big(pb, P, false, nullptr, lm_scratch, lm_stride) : carve_and_stage<D_SHARED>(pb, P, 0);
    __syncwarp();
    const WarpCtx &c = cv.ctx;

    // src/ab_neutral.rs:83-101: ascending stable sort by LSE, first element wins.
    // Ties go to the lowest start id; a NaN anywhere makes the reference panic (:100).
    const abfit_fit *mine = all + (size_t)p * n_starts;
    double best_lse = 0.0;
    int best_id = -1;
    int bad = 0;
    for (int s = lane; s < n_starts; s += 32) {
        const double lse = mine[s].lse;
        const int st = mine[s].status;
        if (st < 0 || lse != lse) {
            bad = 1;
            continue;
        }
        if (best_id < 0 || lse < best_lse) {
            best_lse = lse;
            best_id = s;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ol = __shfl_down_sync(FULL, best_lse, o);
        const int oi = __shfl_down_sync(FULL, best_id, o);
        bad |= __shfl_down_sync(FULL, bad, o);
        if (oi >= 0 && (best_id < 0 || ol < best_lse || (ol == best_lse && oi < best_id))) {
            best_lse = ol;
            best_id = oi;
        }
    }
    best_id = __shfl_sync(FULL, best_id, 0);
    bad = __shfl_sync(FULL, bad, 0);
    if (lane == 0 && prob_status) prob_status[p] = (bad || best_id < 0) ? ABFIT_ERR_NAN : 0;

    if (best_id < 0) {
        if (lane == 0) {
            abfit_fit r;
            r.theta[0] = r.theta[1] = r.theta[2] = r.theta[3] = nan("");
            r.cost = r.lse = nan("");
            r.iters = r.evals = 0;
            r.status = ABFIT_FIT_NAN;
            r.start_id = -1;
            store_fit(best_out + p, r);
        }
        return;
    }
    const abfit_fit b = mine[best_id];
    if (lane == 0) store_fit(best_out + p, b);
    if (!pred && !resid) return;

    // src/ab_neutral.rs:108-135 (every lane computes the same dt table; pairs are strided)
    model_divergence(c, lane, b.theta[0], b.theta[1], b.theta[2]);
    __syncwarp();
    for (int i = lane; i < c.n_pairs; i += 32) {
        const double pr = b.theta[3] + c.lm[(c.offs[i] >> 8) * 32 + lane];
        if (pred) pred[pb.pair_off + i] = pr;
        if (resid) resid[pb.pair_off + i] = c.D[i] - pr;
    }
}

// ---------------------------------------------------------------------------------
// bootstrap refits: each lane owns a replicate with its own D* column
//   D*_i = pred_i + resid[idx_i]      (src/boot_model.rs:50-57)
// The column is built cooperatively by the warp into a [n_pairs][32] scratch tile
// (coalesced on the lane axis), then read once per evaluation.
// ---------------------------------------------------------------------------------
template <bool BIG>
__global__ void __launch_bounds__(32)
k_fit_boot(DevicePools P, const WorkItem *__restrict__ items, int n_boot, const abfit_fit *__restrict__ best,
           const double *__restrict__ pred, const double *__restrict__ resid,
           const int32_t *__restrict__ resample_idx, const double *__restrict__ vary,
           double *__restrict__ dstar_scratch, long long scratch_stride, NMParams nm,
           double *__restrict__ rows_out, abfit_fit *__restrict__ fits_out,
           unsigned long long *__restrict__ evals_per_prob, int *__restrict__ err_flag, double *x_scratch,
           double *lm_scratch, size_t lm_stride)
{
    const int lane = threadIdx.x;
    const WorkItem it = items[blockIdx.x];
    const DevProblem pb = P.probs[it.prob];
    Carved cv = BIG ? carve_big(pb, P, true, x_scratch, lm_scratch, lm_stride) : carve_and_stage<false>(pb, P, 25);
    __syncwarp();
    const WarpCtx &c = cv.ctx;
    const LaneSimplex &S = cv.simplex;
    double *tile = dstar_scratch + (size_t)blockIdx.x * (size_t)scratch_stride;
    const DLaneColumn Dat{tile + lane};
    const double *predp = pred + pb.pair_off;
    const double *residp = resid + pb.pair_off;
    const int32_t *idxp = resample_idx + (size_t)pb.pair_off * n_boot;  // [n_boot][n_pairs] of this problem
    const abfit_fit bm = best[it.prob];

    LaneNM L;
    lane_nm_reset(L);
    int next = it.first;
    const int end = it.first + it.count;
    unsigned long long my_evals = 0;

    for (;;) {
        const bool need = (L.phase == PH_IDLE);
        const unsigned m = __ballot_sync(FULL, need);
        if (m && next < end) {
            const int rank = __popc(m & ((1u << lane) - 1u));
            const int idx = next + rank;
            const bool take = need && idx < end;
            // cooperative build of the D* column of every lane that takes a replicate
            unsigned tm = __ballot_sync(FULL, take);
            while (tm) {
                const int tl = __ffs(tm) - 1;
                tm &= tm - 1;
                const int b = __shfl_sync(FULL, idx, tl);
                const int32_t *ib = idxp + (size_t)b * pb.n_pairs;
                for (int i = lane; i < pb.n_pairs; i += 32) {
                    uint32_t ix = (uint32_t)ib[i];
                    if (ix >= (uint32_t)pb.n_pairs) {  // reported by download_boot; keeps the gather in bounds
                        ix = 0u;
                        *err_flag = 1;
                    }
                    tile[(size_t)i * 32 + tl] = predp[i] + residp[ix];
                }
            }
            __syncwarp();
            if (take) {
                // simplex = [best, vary x 4]  (src/boot_model.rs:69-75)
                const double *vv = vary + ((size_t)it.prob * n_boot + idx) * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) S.X[q * 32] = bm.theta[q];
#pragma unroll
                for (int q = 0; q < 16; ++q) S.X[(4 + q) * 32] = vv[q];
                nm_begin(L, S, idx);
            }
            next += __popc(m);
        }
        const bool active = (L.phase != PH_IDLE);
        const unsigned amask = __ballot_sync(FULL, active);
        if (!amask) break;
        if (active) {
            const double f =
                objective(c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            abfit_fit res;
            if (nm_advance(L, S, nm, f, res, amask)) {
                my_evals += (unsigned long long)res.evals;
                const size_t o = (size_t)it.prob * n_boot + res.start_id;
                // src/boot_model.rs:86-91
                double *row = rows_out + o * 7;
                row[0] = res.theta[0];
                row[1] = res.theta[1];
                row[2] = res.theta[2];
                row[3] = res.theta[3];
                row[4] = p_mm_est(res.theta[0], res.theta[1]);
                row[5] = p_um_est(res.theta[0], res.theta[1]);
                row[6] = p_uu_est(res.theta[0], res.theta[1]);
                if (fits_out) store_fit(fits_out + o, res);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(FULL, my_evals, o);
    if (lane == 0 && evals_per_prob) atomicAdd(evals_per_prob + it.prob, my_evals);
}

// ---------------------------------------------------------------------------------
// bootstrap refits, index-tile variant (n_pairs <= 8191): D* is never materialised; each lane keeps the
// u16 resample indices of its replicate in an L2-resident tile and gathers resid from shared memory
// (see DGather).  Cuts the per-evaluation L2 traffic of k_fit_boot by 4x.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_fit_boot_gather(DevicePools P, const WorkItem *__restrict__ items, int n_boot, const abfit_fit *__restrict__ best,
                  const double *__restrict__ pred, const double *__restrict__ resid,
                  const int32_t *__restrict__ resample_idx, const double *__restrict__ vary,
                  uint2 *__restrict__ idx_scratch, long long scratch_stride, NMParams nm,
                  double *__restrict__ rows_out, abfit_fit *__restrict__ fits_out,
                  unsigned long long *__restrict__ evals_per_prob, int *__restrict__ err_flag, double *x_scratch)
{
    const int lane = threadIdx.x;
    const WorkItem it = items[blockIdx.x];
    const DevProblem pb = P.probs[it.prob];
    Carved cv = carve_and_stage<false>(pb, P, x_scratch ? 5 : 25, x_scratch, boot_gather_lead(pb.n_pairs));
    // resid / pred of this window at the start of shared memory
    extern __shared__ double smem_lead[];
    const int npad = (pb.n_pairs + 1) & ~1;
    double *sresid = smem_lead;
    double *spred = smem_lead + npad;
    for (int i = lane; i < pb.n_pairs; i += 32) {
        spred[i] = pred[pb.pair_off + i];
        sresid[i] = resid[pb.pair_off + i];
    }
    __syncwarp();
    const WarpCtx &c = cv.ctx;
    const LaneSimplex &S = cv.simplex;
    uint2 *tile = idx_scratch + (size_t)blockIdx.x * (size_t)scratch_stride + lane;
    const DGather Dat{tile, spred, reinterpret_cast<const char *>(sresid)};
    const int32_t *idxp = resample_idx + (size_t)pb.pair_off * n_boot;  // [n_boot][n_pairs] of this problem
    const abfit_fit bm = best[it.prob];
    const int ng4 = (pb.n_pairs + 3) >> 2;

    LaneNM L;
    lane_nm_reset(L);
    int next = it.first;
    const int end = it.first + it.count;
    unsigned long long my_evals = 0;

    for (;;) {
        const bool need = (L.phase == PH_IDLE);
        const unsigned m = __ballot_sync(FULL, need);
        if (m && next < end) {
            const int idx = next + __popc(m & ((1u << lane) - 1u));
            if (need && idx < end) {
                // pack this replicate's indices into the lane's tile column (read back by this lane only)
                const int32_t *ib = idxp + (size_t)idx * pb.n_pairs;
                for (int g = 0; g < ng4; ++g) {
                    uint32_t v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        v[q] = (4 * g + q < pb.n_pairs) ? (uint32_t)ib[4 * g + q] : 0u;
                        if (v[q] >= (uint32_t)pb.n_pairs) {  // reported by download_boot; keeps the gather in bounds
                            v[q] = 0u;
                            *err_flag = 1;
                        }
                        v[q] *= 8u;  // byte offset into resid (n_pairs <= 8191)
                    }
                    tile[(size_t)g * 32] = make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16));
                }
                // simplex = [best, vary x 4]  (src/boot_model.rs:69-75)
                const double *vv = vary + ((size_t)it.prob * n_boot + idx) * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) S.X[q * 32] = bm.theta[q];
#pragma unroll
                for (int q = 0; q < 16; ++q) S.X[(4 + q) * 32] = vv[q];
                nm_begin(L, S, idx);
            }
            next += __popc(m);
        }
        const bool active = (L.phase != PH_IDLE);
        const unsigned amask = __ballot_sync(FULL, active);
        if (!amask) break;
        if (active) {
            const double f =
                objective(c, Dat, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            abfit_fit res;
            if (nm_advance(L, S, nm, f, res, amask)) {
                my_evals += (unsigned long long)res.evals;
                const size_t o = (size_t)it.prob * n_boot + res.start_id;
                // src/boot_model.rs:86-91
                double *row = rows_out + o * 7;
                row[0] = res.theta[0];
                row[1] = res.theta[1];
                row[2] = res.theta[2];
                row[3] = res.theta[3];
                row[4] = p_mm_est(res.theta[0], res.theta[1]);
                row[5] = p_um_est(res.theta[0], res.theta[1]);
                row[6] = p_uu_est(res.theta[0], res.theta[1]);
                if (fits_out) store_fit(fits_out + o, res);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_evals += __shfl_down_sync(FULL, my_evals, o);
    if (lane == 0 && evals_per_prob) atomicAdd(evals_per_prob + it.prob, my_evals);
}

// ---------------------------------------------------------------------------------
// warp-per-fit Nelder-Mead (abfit_wide.cuh) for pedigrees with thousands of pairs: one block = one
// warp = one fit at a time.  Every lane carries the same Nelder-Mead state (the state machine of
// abfit_nm.cuh run redundantly, so the control flow is warp-uniform by construction) and the
// 32 lanes share the work of each objective evaluation.  BOOT: bootstrap replicates
// (src/boot_model.rs:41-100) — the warp first materialises D*_i = pred_i + resid[idx_i] as a row of
// the scratch area, then fits it like an observed column.
// ---------------------------------------------------------------------------------
struct WideBoot {
    const abfit_fit *best;
    const double *pred, *resid;
    const int32_t *resample_idx;
    const double *vary;
    double *dstar;  // [block][stride]
    long long stride;
    double *rows_out;
    int *err_flag;
};

template <bool BOOT>
__global__ void __launch_bounds__(32)
k_fit_wide(DevicePools P, const WorkItem *__restrict__ items, const double *__restrict__ simplices, int n_per_prob,
           NMParams nm, abfit_fit *__restrict__ fits_out, unsigned long long *__restrict__ evals_per_prob, WideBoot B)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    const WorkItem it = items[blockIdx.x];
    const DevProblem pb = P.probs[it.prob];
    WideCtx c;
    double *after = wide_carve(c, smem, pb.tmax, pb.n_trip);
    LaneSimplex S;
    S.X = after + lane;
    S.C = S.X + 20 * 32;
    uint32_t *trip = reinterpret_cast<uint32_t *>(after + 25 * 32);
    for (int i = lane; i < pb.n_trip; i += 32) trip[i] = P.wtrip[pb.wtrip_off + i];
    c.trip = trip;
    c.tid = P.wtid + pb.wtid_off;
    c.D = P.D + pb.d_off;
    c.n_pairs = pb.n_pairs;
    c.n_trip = pb.n_trip;
    c.tmax = pb.tmax;
    c.p_uu0 = pb.p_uu0;
    c.p_mm0 = pb.p_mm0;
    c.eqp = pb.eqp;
    c.penw = pb.penw;
    __syncwarp();
    double *drow = BOOT ? B.dstar + (size_t)blockIdx.x * (size_t)B.stride : nullptr;
    unsigned long long my_evals = 0;

    for (int f = it.first; f < it.first + it.count; ++f) {
        if (BOOT) {
            const int32_t *ib = B.resample_idx + (size_t)pb.pair_off * n_per_prob + (size_t)f * pb.n_pairs;
            const double *predp = B.pred + pb.pair_off, *residp = B.resid + pb.pair_off;
            for (int i = lane; i < pb.n_pairs; i += 32) {
                uint32_t ix = (uint32_t)ib[i];
                if (ix >= (uint32_t)pb.n_pairs) {  // reported by download_boot; keeps the gather in bounds
                    ix = 0u;
                    *B.err_flag = 1;
                }
                drow[i] = predp[i] + residp[ix];  // src/boot_model.rs:50-57
            }
            c.D = drow;
            // simplex = [best, vary x 4]  (src/boot_model.rs:69-75)
            const abfit_fit bm = B.best[it.prob];
            const double *vv = B.vary + ((size_t)it.prob * n_per_prob + f) * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) S.X[q * 32] = bm.theta[q];
#pragma unroll
            for (int q = 0; q < 16; ++q) S.X[(4 + q) * 32] = vv[q];
        } else {
            const double *sx = simplices + ((size_t)it.prob * n_per_prob + f) * 20;
#pragma unroll
            for (int q = 0; q < 20; ++q) S.X[q * 32] = sx[q];
        }
        __syncwarp();
        LaneNM L;
        lane_nm_reset(L);
        nm_begin(L, S, f);
        for (;;) {
            const double v = objective_wide(c, lane, L.xt[0], L.xt[1], L.xt[2], L.xt[3], L.phase != PH_LSE);
            abfit_fit res;
            if (nm_advance(L, S, nm, v, res, FULL)) {
                if (lane == 0) {
                    my_evals += (unsigned long long)res.evals;
                    const size_t o = (size_t)it.prob * n_per_prob + res.start_id;
                    if (BOOT) {
                        double *row = B.rows_out + o * 7;  // src/boot_model.rs:86-91
                        row[0] = res.theta[0];
                        row[1] = res.theta[1];
                        row[2] = res.theta[2];
                        row[3] = res.theta[3];
                        row[4] = p_mm_est(res.theta[0], res.theta[1]);
                        row[5] = p_um_est(res.theta[0], res.theta[1]);
                        row[6] = p_uu_est(res.theta[0], res.theta[1]);
                    }
                    if (fits_out) store_fit(fits_out + o, res);
                }
                break;
            }
        }
        __syncwarp();  // every lane is done with this replicate's D* row and simplex
    }
    if (lane == 0 && evals_per_prob && my_evals) atomicAdd(evals_per_prob + it.prob, my_evals);
}

// ---------------------------------------------------------------------------------
// objective only (test hook / CostFunction seam)
// ---------------------------------------------------------------------------------
template <bool D_SHARED, bool BIG>
__global__ void __launch_bounds__(32)
k_cost_batch(DevicePools P, const WorkItem *__restrict__ items, const double *__restrict__ theta,
             double *__restrict__ cost_out, double *__restrict__ lse_out, double *lm_scratch, size_t lm_stride)
{
    const int lane = threadIdx.x;
    const WorkItem it = items[blockIdx.x];
    const DevProblem pb = P.probs[it.prob];
    Carved cv = BIG ? carve_big(pb, P, false, nullptr, lm_scratch, lm_stride) : carve_and_stage<D_SHARED>(pb, P, 0);
    __syncwarp();
    const DBroadcast Dat{cv.ctx.D};
    if (lane < it.count) {
        const double *th = theta + (size_t)(it.first + lane) * 4;
        const double a = th[0], b = th[1], w = th[2], ic = th[3];
        cost_out[it.first + lane] = objective(cv.ctx, Dat, lane, a, b, w, ic, true);
        if (lse_out) lse_out[it.first + lane] = objective(cv.ctx, Dat, lane, a, b, w, ic, false);
    }
}

template <bool BIG>
__global__ void __launch_bounds__(32)
k_model_div(DevicePools P, const double *__restrict__ theta4, double *__restrict__ dt_out,
            double *__restrict__ puu_out, double *lm_scratch, size_t lm_stride)
{
    const int lane = threadIdx.x;
    const DevProblem pb = P.probs[0];
    Carved cv = BIG ? carve_big(pb, P, false, nullptr, lm_scratch, lm_stride) : carve_and_stage<false>(pb, P, 0);
    __syncwarp();
    const WarpCtx &c = cv.ctx;
    model_divergence(c, lane, theta4[0], theta4[1], theta4[2]);
    __syncwarp();
    for (int i = lane; i < c.n_pairs; i += 32) dt_out[i] = c.lm[(c.offs[i] >> 8) * 32 + lane];
    if (lane == 0 && puu_out) *puu_out = p_uu_est(theta4[0], theta4[1]);
}

// ---------------------------------------------------------------------------------
// FP64 roofline micro-benchmark: 8 independent DFMA chains per thread
// ---------------------------------------------------------------------------------
// Model::vary x 4 per bootstrap replicate (src/boot_model.rs:69-75, src/structs.rs:100-128), one thread per coordinate
__global__ void k_gen_vary(uint64_t seed, uint64_t first_problem_id, int n_probs, int n_boot,
                           const abfit_fit *__restrict__ best, double *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_prob = (size_t)n_boot * 16;
    if (i >= (size_t)n_probs * per_prob) return;
    const int p = (int)(i / per_prob);
    const size_t r = i - (size_t)p * per_prob;
    const int b = (int)(r >> 4), v = (int)((r >> 2) & 3), j = (int)(r & 3);
    out[i] = vary_coordinate(seed, first_problem_id + (uint64_t)p, (uint64_t)b, v, j, best[p].theta[j]);
}

__global__ void k_fp64_peak(int iters, double *sink)
{
    double a0 = 1.0 + threadIdx.x * 1e-9, a1 = a0 + 1e-9, a2 = a0 + 2e-9, a3 = a0 + 3e-9;
    double a4 = a0 + 4e-9, a5 = a0 + 5e-9, a6 = a0 + 6e-9, a7 = a0 + 7e-9;
    const double m = 0.9999999, d = 1e-7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = __fma_rn(a0, m, d); a1 = __fma_rn(a1, m, d); a2 = __fma_rn(a2, m, d); a3 = __fma_rn(a3, m, d);
            a4 = __fma_rn(a4, m, d); a5 = __fma_rn(a5, m, d); a6 = __fma_rn(a6, m, d); a7 = __fma_rn(a7, m, d);
        }
    }
    const double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 123.456) sink[0] = s;  // never true; keeps the chains alive
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
int max_dynamic_smem(int device)
{
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
    return v;
}

template <class K>
static int prep_kernel(K kernel, size_t smem_bytes, bool max_shared = true)
{
    if (smem_bytes > 48 * 1024)
        ABFIT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    // prefer the largest shared-memory carve-out: these kernels keep all per-lane state in shared (the
    // warp-per-fit kernels stream D and the triple ids through L1 instead and leave the split to the driver)
    if (max_shared)
        ABFIT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    return 0;
}

template <bool D_SHARED, bool X_GLOBAL, bool BIG>
static int launch_fit_starts_t(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items, int n_warps,
                               const double *simplices, int n_starts, NMParams nm, abfit_fit *all_out,
                               unsigned long long *evals_per_prob, size_t smem_bytes, double *x_scratch,
                               const BigScratch &big)
{
    if (int rc = prep_kernel(k_fit_starts<D_SHARED, X_GLOBAL, BIG>, smem_bytes)) return rc;
    if (getenv("ABFIT_DEV_VERBOSE")) {
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_fit_starts<D_SHARED, X_GLOBAL, BIG>, 32 * n_warps, smem_bytes);
        fprintf(stderr, "[abfit] k_fit_starts<D_SHARED=%d,X_GLOBAL=%d,BIG=%d>: %d blocks x %d warps, %zu B smem/block, %d blocks/SM\n",
                (int)D_SHARED, (int)X_GLOBAL, (int)BIG, n_items, n_warps, smem_bytes, nb);
    }
    k_fit_starts<D_SHARED, X_GLOBAL, BIG><<<n_items, 32 * n_warps, smem_bytes, st>>>(
        P, items, simplices, n_starts, nm, all_out, evals_per_prob, x_scratch, big.lm, big.lm_stride);
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_fit_starts(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items, int n_warps,
                      const double *simplices, int n_starts, NMParams nm, abfit_fit *all_out,
                      unsigned long long *evals_per_prob, size_t smem_bytes, bool d_in_shared, double *x_scratch,
                      const BigScratch &big)
{
    if (n_items <= 0) return 0;
    if (const char *pad = getenv("ABFIT_DEV_SMEM_PAD")) smem_bytes += (size_t)atoi(pad);  // occupancy experiments
#define ABFIT_FS(DS, XG, BG)                                                                                         \
    launch_fit_starts_t<DS, XG, BG>(st, P, items, n_items, n_warps, simplices, n_starts, nm, all_out, evals_per_prob, \
                                    smem_bytes, x_scratch, big)
    if (big.lm) return ABFIT_FS(false, true, true);
    if (d_in_shared) return x_scratch ? ABFIT_FS(true, true, false) : ABFIT_FS(true, false, false);
    return x_scratch ? ABFIT_FS(false, true, false) : ABFIT_FS(false, false, false);
#undef ABFIT_FS
}

int launch_select(cudaStream_t st, const DevicePools &P, int n_probs, int n_starts, const abfit_fit *all,
                  abfit_fit *best_out, double *pred, double *resid, int32_t *prob_status, size_t smem_bytes,
                  bool d_in_shared, const BigScratch &big)
{
    if (n_probs <= 0) return 0;
    if (big.lm) {
        if (int rc = prep_kernel(k_select<false, true>, smem_bytes)) return rc;
        k_select<false, true><<<n_probs, 32, smem_bytes, st>>>(P, n_starts, all, best_out, pred, resid, prob_status,
                                                               big.lm, big.lm_stride);
    } else if (d_in_shared) {
        if (int rc = prep_kernel(k_select<true, false>, smem_bytes)) return rc;
        k_select<true, false><<<n_probs, 32, smem_bytes, st>>>(P, n_starts, all, best_out, pred, resid, prob_status,
                                                               nullptr, 0);
    } else {
        if (int rc = prep_kernel(k_select<false, false>, smem_bytes)) return rc;
        k_select<false, false><<<n_probs, 32, smem_bytes, st>>>(P, n_starts, all, best_out, pred, resid, prob_status,
                                                                nullptr, 0);
    }
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_fit_boot(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items, int n_boot,
                    const abfit_fit *best, const double *pred, const double *resid, const int32_t *resample_idx,
                    const double *vary, double *dstar_scratch, int64_t scratch_stride, NMParams nm,
                    double *rows_out, abfit_fit *fits_out, unsigned long long *evals_per_prob, size_t smem_bytes,
                    int *err_flag, const BigScratch &big)
{
    if (n_items <= 0) return 0;
    if (big.lm) {
        if (int rc = prep_kernel(k_fit_boot<true>, smem_bytes)) return rc;
        k_fit_boot<true><<<n_items, 32, smem_bytes, st>>>(P, items, n_boot, best, pred, resid, resample_idx, vary,
                                                          dstar_scratch, (long long)scratch_stride, nm, rows_out,
                                                          fits_out, evals_per_prob, err_flag, big.x, big.lm, big.lm_stride);
    } else {
        if (int rc = prep_kernel(k_fit_boot<false>, smem_bytes)) return rc;
        k_fit_boot<false><<<n_items, 32, smem_bytes, st>>>(P, items, n_boot, best, pred, resid, resample_idx, vary,
                                                           dstar_scratch, (long long)scratch_stride, nm, rows_out,
                                                           fits_out, evals_per_prob, err_flag, nullptr, nullptr, 0);
    }
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_fit_boot_gather(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items, int n_boot,
                           const abfit_fit *best, const double *pred, const double *resid, const int32_t *resample_idx,
                           const double *vary, void *idx_scratch, int64_t scratch_stride, NMParams nm,
                           double *rows_out, abfit_fit *fits_out, unsigned long long *evals_per_prob, size_t smem_bytes,
                           int *err_flag, double *x_scratch)
{
    if (n_items <= 0) return 0;
    if (int rc = prep_kernel(k_fit_boot_gather, smem_bytes)) return rc;
    k_fit_boot_gather<<<n_items, 32, smem_bytes, st>>>(P, items, n_boot, best, pred, resid, resample_idx, vary,
                                                       static_cast<uint2 *>(idx_scratch), (long long)scratch_stride, nm,
                                                       rows_out, fits_out, evals_per_prob, err_flag, x_scratch);
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_fit_starts_wide(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items,
                           const double *simplices, int n_starts, NMParams nm, abfit_fit *all_out,
                           unsigned long long *evals_per_prob, size_t smem_bytes)
{
    if (n_items <= 0) return 0;
    if (int rc = prep_kernel(k_fit_wide<false>, smem_bytes, false)) return rc;
    if (getenv("ABFIT_DEV_VERBOSE"))
        fprintf(stderr, "[abfit] k_fit_wide<starts>: %d blocks x 1 warp, %zu B smem/block\n", n_items, smem_bytes);
    k_fit_wide<false><<<n_items, 32, smem_bytes, st>>>(P, items, simplices, n_starts, nm, all_out, evals_per_prob,
                                                       WideBoot{});
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_fit_boot_wide(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items, int n_boot,
                         const abfit_fit *best, const double *pred, const double *resid, const int32_t *resample_idx,
                         const double *vary, double *dstar_scratch, int64_t scratch_stride, NMParams nm,
                         double *rows_out, abfit_fit *fits_out, unsigned long long *evals_per_prob, size_t smem_bytes,
                         int *err_flag)
{
    if (n_items <= 0) return 0;
    if (int rc = prep_kernel(k_fit_wide<true>, smem_bytes, false)) return rc;
    if (getenv("ABFIT_DEV_VERBOSE"))
        fprintf(stderr, "[abfit] k_fit_wide<boot>: %d blocks x 1 warp, %zu B smem/block\n", n_items, smem_bytes);
    WideBoot B{best, pred, resid, resample_idx, vary, dstar_scratch, (long long)scratch_stride, rows_out, err_flag};
    k_fit_wide<true><<<n_items, 32, smem_bytes, st>>>(P, items, nullptr, n_boot, nm, fits_out, evals_per_prob, B);
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_cost_batch(cudaStream_t st, const DevicePools &P, const WorkItem *items, int n_items,
                      const double *theta, double *cost_out, double *lse_out, size_t smem_bytes, bool d_in_shared,
                      const BigScratch &big)
{
    if (n_items <= 0) return 0;
    if (big.lm) {
        if (int rc = prep_kernel(k_cost_batch<false, true>, smem_bytes)) return rc;
        k_cost_batch<false, true><<<n_items, 32, smem_bytes, st>>>(P, items, theta, cost_out, lse_out, big.lm, big.lm_stride);
    } else if (d_in_shared) {
        if (int rc = prep_kernel(k_cost_batch<true, false>, smem_bytes)) return rc;
        k_cost_batch<true, false><<<n_items, 32, smem_bytes, st>>>(P, items, theta, cost_out, lse_out, nullptr, 0);
    } else {
        if (int rc = prep_kernel(k_cost_batch<false, false>, smem_bytes)) return rc;
        k_cost_batch<false, false><<<n_items, 32, smem_bytes, st>>>(P, items, theta, cost_out, lse_out, nullptr, 0);
    }
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_model_divergence(cudaStream_t st, const DevicePools &P, const double *theta4, double *dt_out,
                            double *puu_out, size_t smem_bytes, const BigScratch &big)
{
    if (big.lm) {
        if (int rc = prep_kernel(k_model_div<true>, smem_bytes)) return rc;
        k_model_div<true><<<1, 32, smem_bytes, st>>>(P, theta4, dt_out, puu_out, big.lm, big.lm_stride);
    } else {
        if (int rc = prep_kernel(k_model_div<false>, smem_bytes)) return rc;
        k_model_div<false><<<1, 32, smem_bytes, st>>>(P, theta4, dt_out, puu_out, nullptr, 0);
    }
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_gen_vary(cudaStream_t st, uint64_t seed, uint64_t first_problem_id, int n_probs, int n_boot,
                    const abfit_fit *best, double *out)
{
    const size_t n = (size_t)n_probs * n_boot * 16;
    if (n == 0) return 0;
    k_gen_vary<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(seed, first_problem_id, n_probs, n_boot, best, out);
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

int launch_fp64_peak(cudaStream_t st, int blocks, int threads, int iters, double *sink)
{
    k_fp64_peak<<<blocks, threads, 0, st>>>(iters, sink);
    ABFIT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace abfit
