"""Pure-Python twin of libabfit's seeded input generators (csrc/abfit_rng.h, abfit_gen_* in csrc/abfit_api.cu).

TEST / BENCH INFRASTRUCTURE.  north_star: "random starts and bootstrap resamples are generated on the host from the
same seeds and fed to both implementations".  The reference arm of bench.py uses this twin so that its process never
maps the product library; tests/test_abi.py pins it bit for bit against the C generators.

Integer hashing is done on numpy uint64 (wrapping); the two libm calls of the C code (`std::pow`) go through
math.pow, i.e. the same glibc function, so the doubles are identical."""
import math

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(z):
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
        return z ^ (z >> np.uint64(31))


def u01(seed, stream, problem, item, sub):
    """abfit_rng.h::u01 — item / sub may be arrays (broadcast)"""
    with np.errstate(over="ignore"):
        h = _mix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1342543DE82EF95)))
        h = _mix64(h ^ np.uint64(problem))
        h = _mix64(h ^ np.asarray(item, dtype=np.uint64))
        h = _mix64(h ^ np.asarray(sub, dtype=np.uint64))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def _uniform(lo, hi, u):
    return lo + (hi - lo) * u


def gen_start_simplices(seed, problem_id, n_starts, max_divergence):
    """abfit_gen_start_simplices: [n_starts][5][4] (Model::new, src/structs.rs:78-96)"""
    mx = max_divergence if max_divergence > 0.0 else 0.1
    s = np.arange(n_starts, dtype=np.uint64)[:, None, None]
    sub = (np.arange(5, dtype=np.uint64)[None, :, None] * np.uint64(4)) + np.arange(4, dtype=np.uint64)[None, None, :]
    u = u01(seed, 1, problem_id, s, sub)
    out = np.empty((n_starts, 5, 4))
    e = _uniform(-9.0, -2.0, u[..., :2])
    out[..., :2] = np.array([math.pow(10.0, x) for x in e.ravel()]).reshape(e.shape)
    out[..., 2] = _uniform(0.0, 0.1, u[..., 2])
    out[..., 3] = _uniform(0.0, mx, u[..., 3])
    return out


def gen_resample_idx(seed, problem_id, n_boot, n_pairs):
    """abfit_gen_resample_idx: [n_boot][n_pairs] int32"""
    b = np.arange(n_boot, dtype=np.uint64)[:, None]
    i = np.arange(n_pairs, dtype=np.uint64)[None, :]
    k = (u01(seed, 3, problem_id, b, i) * float(n_pairs)).astype(np.int32)
    return np.minimum(k, n_pairs - 1).astype(np.int32)


def gen_vary_vertices(seed, problem_id, n_boot, best_theta):
    """abfit_gen_vary_vertices: [n_boot][4][4] (Model::vary, src/structs.rs:100-128)"""
    best = np.asarray(best_theta, dtype=np.float64)
    n = np.where(best == 0.0, 0.1, best)
    a = np.abs(n)
    lo, hi = n - a * 0.1, n + a * 0.1
    swap = lo >= hi
    lo, hi = np.where(swap, hi, lo), np.where(swap, lo, hi)
    b = np.arange(n_boot, dtype=np.uint64)[:, None, None]
    sub = (np.arange(4, dtype=np.uint64)[None, :, None] * np.uint64(4)) + np.arange(4, dtype=np.uint64)[None, None, :]
    u = u01(seed, 2, problem_id, b, sub)
    return _uniform(lo[None, None, :], hi[None, None, :], u)
