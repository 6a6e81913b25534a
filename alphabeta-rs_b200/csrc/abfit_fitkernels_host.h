// abfit_fitkernels_host.h — the shared-memory slot sizes of the continuous-scheduling kernels (abfit_fitkernels.cuh),
// for host code that sizes launches without including the device bodies.
#pragma once
namespace abfit {
#if defined(__CUDACC__)
#define ABFIT_HD_INLINE __host__ __device__ inline
#else
#define ABFIT_HD_INLINE inline
#endif
// window slots per warp: a lane starts a fit of the open item's window while its neighbours finish fits of up to
// V2_SLOTS - 1 earlier windows; with four, an item as short as one fit per lane is reopened without waiting
#ifndef ABFIT_V2_SLOTS
#define ABFIT_V2_SLOTS 4
#endif
constexpr int V2_SLOTS = ABFIT_V2_SLOTS;  // (a generated source may define another count: occupancy experiments)
// multi-start slot: [D npad][p_uu0, p_mm0, eqp, penw]
ABFIT_HD_INLINE int v2_fit_slot_doubles(int n_pairs) { return ((n_pairs + 1) & ~1) + 4; }
// bootstrap slot: [resid npad][pred npad][p_uu0, p_mm0, eqp, penw][best theta 4]; two slots (an item is a whole
// window's replicates; measured: four slots with the predictions read through L1 instead were 5 % slower — the
// kernel is bound by the load/store unit, and 12 resident warps need the slots to stay within 12 KB)
constexpr int V2_BOOT_SLOTS = 2;
// warps per block of the continuous-scheduling kernels.  The warps of a block are independent (own queue state, own
// slots); a block-wide barrier at the top of every evaluation only keeps them in PHASE, so that they walk through the
// (tens of KB of straight-line) objective code together and share its instruction-cache lines.
#ifndef ABFIT_V2_WARPS
#define ABFIT_V2_WARPS 1
#endif
constexpr int V2_WARPS = ABFIT_V2_WARPS;
// warps per block of the multi-start kernel when its warps MERGE during the drain (abfit_fitkernels.cuh, "drain
// merging"): 1 = off (one-warp blocks).  The warps run out of phase (no barrier per evaluation); barriers are used only
// by the merge rounds at the end of a launch.  A block ends with V2_FIT_CTL_BYTES of control words behind the warps'
// regions.
#ifndef ABFIT_V2_FIT_WARPS
#define ABFIT_V2_FIT_WARPS 1
#endif
constexpr int V2_FIT_WARPS = ABFIT_V2_FIT_WARPS;
constexpr int V2_FIT_CTL_BYTES = 128;
ABFIT_HD_INLINE int v2_boot_slot_doubles(int n_pairs) { return 2 * ((n_pairs + 1) & ~1) + 8; }
}  // namespace abfit
