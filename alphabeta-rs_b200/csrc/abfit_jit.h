// abfit_jit.h — run-time specialisation of the throughput kernels for a batch that shares one pedigree
// program (see abfit_jit.cu).
#pragma once
#include <string>

#include "abfit_fitkernels_host.h"
#include "abfit_plan.h"

namespace abfit {

constexpr int JIT_SPEC_NLANE = 8;    // per-lane shared doubles of a specialised kernel: the hand-off's parking area
constexpr int JIT_MAX_PAIRS = 1024;  // unrolled pair loop: ~5.5 instructions per pair
constexpr int JIT_MAX_SLOTS = 160;   // per-lane values of the program, all kept in registers

struct EmbeddedSource {
    const char *name;
    const char *text;
};

struct JitModule {
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fit_starts = nullptr;
    cudaKernel_t fit_boot_gather = nullptr;
    int slots = 4;  // window slots per warp (sched 2)
    int warps = 1;  // warps per block (sched 2): independent warps kept in phase by a block barrier per evaluation
    int fit_warps = 1;       // warps per block of the multi-start kernel (= warps unless drain merging is on)
    bool fit_merge = false;  // multi-start kernel built with drain merging (ABFIT_V2_FIT_WARPS > 1)
    int sched = 2;  // 1: block-per-item bodies, 2: continuous lane scheduling (persistent one-warp blocks)
    double compile_seconds = 0.0;
    bool from_disk_cache = false;
};

// CUDA C++ source of the kernels specialised for problem p's program (pure function of the program)
std::string jit_generate_source(const HostPlan &hp, int p, int sched);
int jit_default_sched();
// can this batch run on specialised kernels at all (one shared program, small enough, regular launch shape)?
bool jit_eligible(const HostPlan &hp, const LaunchShape &shape, std::string *why);
// NVRTC -> sm_100a cubin (disk cache under $ABFIT_CACHE_DIR, default ~/.cache/abfit).  No GPU needed.
int jit_compile(const std::string &source, std::string &cubin, std::string &log, double *seconds, bool *from_disk);
// compiled + loaded module for problem p's program (process-wide cache); note receives the compiler log on failure
int jit_get_module(const HostPlan &hp, int p, const JitModule **out, std::string *note);
bool jit_is_cached(const HostPlan &hp, int p);
std::string jit_last_note();  // why the last specialisation attempt of this process failed (empty: none did)

size_t jit_smem_fit(const DevProblem &pb, int n_warps);
size_t jit_smem_boot_gather(const DevProblem &pb, bool x_global);
int jit_launch_fit_starts(const JitModule *m, cudaStream_t st, DevicePools P, const WorkItem *items, int n_items,
                          int n_warps, const double *simplices, int n_starts, NMParams nm, abfit_fit *all_out,
                          unsigned long long *evals_per_prob, size_t smem_bytes);
int jit_launch_fit_boot_gather(const JitModule *m, cudaStream_t st, DevicePools P, const WorkItem *items, int n_items,
                               int n_boot, const abfit_fit *best, const double *pred, const double *resid,
                               const int32_t *resample_idx, const double *vary, void *idx_scratch,
                               int64_t scratch_stride, NMParams nm, double *rows_out, abfit_fit *fits_out,
                               unsigned long long *evals_per_prob, size_t smem_bytes, int *err_flag, double *x_scratch);

// continuous lane scheduling: persistent one-warp blocks, items opened from a global cursor (cursor: device int)
size_t jit_smem_fit_v2(const JitModule *m, const DevProblem &pb);
size_t jit_smem_boot_v2(const JitModule *m, const DevProblem &pb);
int jit_resident_warps(const JitModule *m, bool boot, const DevProblem &pb, int n_sm);  // grid of a full machine
int jit_launch_fit_starts_v2(const JitModule *m, cudaStream_t st, DevicePools P, const WorkItem *items, int n_items,
                             int64_t n_fits, int grid_max, int *cursor, const double *simplices, int n_starts, NMParams nm,
                             abfit_fit *all_out, unsigned long long *evals_per_prob, size_t smem_bytes,
                             const int *sx_ready = nullptr /* device int: windows whose start simplices have been uploaded */);
int jit_launch_fit_boot_gather_v2(const JitModule *m, cudaStream_t st, DevicePools P, const WorkItem *items, int n_items,
                                  int64_t n_fits, int grid_max, int *cursor, int n_boot, const abfit_fit *best,
                                  const double *pred, const double *resid, const int32_t *resample_idx, const double *vary,
                                  void *idx_scratch, int64_t scratch_stride, NMParams nm, double *rows_out,
                                  abfit_fit *fits_out, unsigned long long *evals_per_prob, size_t smem_bytes, int *err_flag);

}  // namespace abfit
