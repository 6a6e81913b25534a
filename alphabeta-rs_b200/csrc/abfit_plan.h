// abfit_plan.h — host-side compilation of pedigrees into the per-window micro-op programs
// the kernels interpret (see abfit_model.cuh), and the block/warp work decomposition.
#pragma once
#include "abfit_internal.h"

namespace abfit {

struct HostPlan {
    std::vector<DevProblem> probs;
    std::vector<double> D;         // every problem padded to an even count (16-byte alignment)
    std::vector<uint32_t> offs;    // 256 * lane_mem index of each pair's dt1t2, padded to 4 per problem
    std::vector<OpWord> ops;
    std::vector<uint32_t> wtrip;   // distinct triples of each program (warp-per-fit kernels, abfit_wide.cuh)
    std::vector<uint32_t> wtid;    // triple id of every pair
    std::vector<double> flops;     // algorithmic FLOPs per objective evaluation
    std::vector<double> fp64_instr;  // FP64 instructions one lane executes per objective evaluation (program-derived)
    std::vector<int32_t> n_triples, tmax;
    std::vector<uint8_t> d_has_nan;
    int32_t max_pairs = 0;
    int64_t total_pairs = 0;
    std::vector<uint32_t> keys;    // scratch of compile_problems: t0 | (t1 - t0) << 8 | (t2 - t0) << 16 per pedigree row
    // empties the plan but keeps the storage: a context's workspace plan is refilled by every one-shot call, and
    // returning 50 MB to the allocator only to fault the same pages in again cost milliseconds per call
    void clear()
    {
        probs.clear();
        D.clear();
        offs.clear();
        ops.clear();
        wtrip.clear();
        wtid.clear();
        flops.clear();
        fp64_instr.clear();
        n_triples.clear();
        tmax.clear();
        d_has_nan.clear();
        keys.clear();
        max_pairs = 0;
        total_pairs = 0;
    }
};

// Validates and compiles n_probs problems. Returns 0 or a negative abfit_status (message via set_error).
int compile_problems(const abfit_problem *probs, int n_probs, HostPlan &hp);

// kernel configuration of one batch for a device with `smem_cap` bytes of opt-in shared memory per block
struct LaunchShape {
    int n_warps = 1;        // warps per block of the multi-start kernel
    bool d_shared = true;   // D column staged in shared memory (else broadcast from L1/L2)
    bool x_global = false;  // simplex vertices in a global scratch area (20 * 32 doubles per warp) instead of shared
    bool big = false;       // per-lane model state does not fit in shared memory: everything per-lane in global scratch
    int n_lane_max = 0;     // doubles of per-lane model state of the largest problem
    size_t smem_fit = 0;    // k_fit_starts
    size_t smem_boot = 0;   // k_fit_boot (1 warp, D* comes from the scratch tile)
    size_t smem_boot_gather = 0;  // k_fit_boot_gather (1 warp + pred/resid of the window), 0 = not usable
    bool boot_x_global = false;   // k_fit_boot_gather keeps the simplex vertices in global scratch (more resident warps)
    size_t smem_aux = 0;    // k_select / k_cost_batch / k_model_div (1 warp, no simplex)
    bool d_shared_aux = true;
    // warp-per-fit kernels (abfit_wide.cuh) replace the global-scratch variant of the Nelder-Mead kernels whenever
    // their per-warp tables fit in shared memory: one block = one warp
    bool wide = false;
    size_t smem_wide = 0;
};
size_t smem_need_wide(const DevProblem &pb);
std::vector<WorkItem> make_items_wide(const HostPlan &hp, int count_per_prob, int n_sm, bool skip_nan);
int choose_launch_shape(const HostPlan &hp, size_t smem_cap, size_t smem_per_sm, int fits_per_prob, LaunchShape &out);

// largest y with sqrt(y) < sd_tol (-1 when no y >= 0 qualifies): lets the kernels evaluate argmin's
// `sd < sd_tolerance` termination test without taking the square root (see NMParams::var_thr)
double nm_var_threshold(double sd_tol);
// spread of the sorted simplex costs above which the termination test certainly fails (NMParams::range_thr)
double nm_range_threshold(double var_thr);
NMParams nm_params(int max_iters, double sd_tol, uint32_t flags);

// chunks of `count_per_prob` fits per problem; every block gets at least 32 * n_warps * 2 fits when it can
std::vector<WorkItem> make_items(const HostPlan &hp, int count_per_prob, int n_sm, int n_warps, bool skip_nan);

// Items of the continuous-scheduling kernels (persistent warps pulling from a global cursor): guided self-scheduling —
// an item takes 1/(2 P) of the fits that are still unassigned (P = resident warps), at most `max_chunk` and one
// window, at least 32 — so the bulk of a batch runs in long items and the warps run out of work within about one
// fit of each other.
std::vector<WorkItem> make_items_guided(const HostPlan &hp, int count_per_prob, int resident_warps, int max_chunk,
                                        bool skip_nan, int p_first = 0, int p_end = -1 /* windows [p_first, p_end) */);

}  // namespace abfit
