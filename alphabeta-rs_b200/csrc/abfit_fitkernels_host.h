// abfit_fitkernels_host.h — the shared-memory slot sizes of the continuous-scheduling kernels (abfit_fitkernels.cuh),
// for host code that sizes launches without including the device bodies.
#pragma once
namespace abfit {
#if defined(__CUDACC__)
#define ABFIT_HD_INLINE __host__ __device__ inline
#else
#define ABFIT_HD_INLINE inline
#endif
// multi-start slot: [D npad][p_uu0, p_mm0, eqp, penw]
ABFIT_HD_INLINE int v2_fit_slot_doubles(int n_pairs) { return ((n_pairs + 1) & ~1) + 4; }
// bootstrap slot: [resid npad][pred npad][p_uu0, p_mm0, eqp, penw][best theta 4]
ABFIT_HD_INLINE int v2_boot_slot_doubles(int n_pairs) { return 2 * ((n_pairs + 1) & ~1) + 8; }
}  // namespace abfit
