// shims.h — CUDA intrinsics used by abfit_model.cuh / abfit_nm.cuh, for the host build (tests only)
#pragma once
#include <algorithm>
#include <cmath>
#include <cuda_runtime.h>
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
static inline double __ldcg(const double *p) { return *p; }
static inline void __syncwarp(unsigned) {}
using std::min;
using std::sqrt;
