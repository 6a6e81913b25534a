"""The device-side model / objective / Nelder-Mead headers (alphabeta-rs_b200/csrc/abfit_model.cuh,
abfit_nm.cuh) and the host pedigree compiler (abfit_plan.cu), built FOR THE CPU (tests/host_emul) and
checked bit for bit against the oracle.  No GPU needed: this is how the `-m "not gpu"` suite covers
the micro-op compiler, the software-pipelined pair loop and the NM state machine, and how a compiler
problem on the GPU side is told apart from a source problem."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

EMUL_DIR = os.path.join(ROOT, "tests", "host_emul")


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["make", "-C", EMUL_DIR, "-s", "-B", "libemul.so"])
    return C.CDLL(os.path.join(EMUL_DIR, "libemul.so"))


def emul_cost(emul, ab, ped, p0uu, eqp, w, theta, wide=False):
    arr = ab._pack_problems([ab.Problem(ped, p0uu, eqp, w)])
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, 4)
    cost, lse = np.empty(len(theta)), np.empty(len(theta))
    rc = (emul.emul_cost_wide if wide else emul.emul_cost)(arr, theta.ctypes.data_as(C.c_void_p), len(theta), cost.ctypes.data_as(C.c_void_p),
                        lse.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return cost, lse


def emul_fit(emul, ab, ped, p0uu, sx, max_iters=10000, flags=0, dstar=None, wide=False):
    arr = ab._pack_problems([ab.Problem(ped, p0uu, p0uu, 1.0)])
    sx = np.ascontiguousarray(sx, dtype=np.float64)
    n = sx.size // 20
    out = np.zeros(n, dtype=ab.FIT_DTYPE)
    dp = None if dstar is None else np.ascontiguousarray(dstar, dtype=np.float64).ctypes.data_as(C.c_void_p)
    rc = (emul.emul_fit_wide if wide else emul.emul_fit)(arr, sx.ctypes.data_as(C.c_void_p), n, dp, max_iters, C.c_double(ab.DBL_EPSILON), flags,
                       out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return out


def random_pedigree(rng, n_pairs, tmax, style):
    """time structures that exercise every micro-op: identity operands (t1 == t0, t2 == t0, all equal),
    exponent 1 (G itself), stored powers, deferred d-vectors (t0 larger than both exponents)."""
    rows = []
    for _ in range(n_pairs):
        if style == "lineage":  # (0, a, b), a <= b: the usual mutation-accumulation pedigree
            a, b = sorted(rng.integers(0, tmax + 1, 2))
            t0 = 0
        elif style == "sibling":  # common ancestor late in the pedigree
            t0 = int(rng.integers(0, tmax))
            a, b = t0 + rng.integers(0, 3, 2)
        else:  # anything valid, unsorted t1 / t2
            t0 = int(rng.integers(0, tmax + 1))
            a, b = rng.integers(t0, tmax + 1, 2)
        rows.append([t0, a, b, rng.uniform(0, 0.3)])
    return np.array(rows, dtype=np.float64)


@pytest.mark.parametrize("style", ["lineage", "sibling", "any"])
@pytest.mark.parametrize("tmax", [1, 4, 33, 127])
def test_compiled_program_cost_is_bit_exact(emul, ab, oracle, style, tmax):
    rng = np.random.default_rng(hash((style, tmax)) % 2**32)
    for n_pairs in (1, 3, 4, 5, 37, 120, 131, 260):
        ped = random_pedigree(rng, n_pairs, tmax, style)
        B = 40
        theta = np.stack([10 ** rng.uniform(-7, -1.5, B), 10 ** rng.uniform(-7, -1.5, B), rng.uniform(-0.1, 0.3, B),
                          rng.uniform(0, 0.1, B)], axis=1)
        theta[3] = [-2e-4, 3e-3, -0.2, 0.01]
        cost, lse = emul_cost(emul, ab, ped, 0.8, 0.7, 1.3, theta)
        wcost, wlse = emul_cost(emul, ab, ped, 0.8, 0.7, 1.3, theta, wide=True)  # warp-per-fit formulation
        pb = oracle.Problem(ped, 0.8, 0.7, 1.3)
        for i in range(B):
            assert cost[i] == oracle.cost(pb, theta[i]), (style, tmax, n_pairs, i)
            assert lse[i] == oracle.lse(pb, theta[i], flags=oracle.FAST_DIVERGENCE), (style, tmax, n_pairs, i)
        assert np.array_equal(wcost, cost, equal_nan=True) and np.array_equal(wlse, lse, equal_nan=True)


@pytest.mark.parametrize("wide", [False, True])
def test_golden_cost_kat_through_the_device_headers(emul, ab, oracle, wide):
    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    cost, _ = emul_cost(emul, ab, ped, 0.75, 0.5, 0.7, [[0.0001179555, 0.0001180614, 0.03693534, 0.003023981]], wide=wide)
    assert cost[0] == 0.0006700888539608879  # src/structs.rs:233


@pytest.mark.parametrize("flags", [0, 1, 2])
def test_nelder_mead_state_machine_matches_oracle(emul, ab, oracle, flags):
    """flags: 0 = argmin 0.8.1 with stall early exit, 1 = shrink on failed contraction, 2 = literal (no early exit)"""
    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    ped6 = np.loadtxt(os.path.join(GOLDEN, "pedigree_generated.txt"), skiprows=1)
    max_iters = 400 if flags == 2 else 10000
    oflags = oracle.FAST_DIVERGENCE | {0: oracle.EARLY_EXIT_ON_STALL, 1: oracle.SHRINK_ON_FAILED_CONTRACTION, 2: 0}[flags]
    for p, u in ((ped, 0.8), (ped6, 0.655)):
        sx = ab.gen_start_simplices(7, 3, 24, float(p[:, 3].max()))
        g = emul_fit(emul, ab, p, u, sx, max_iters=max_iters, flags=flags)
        gw = emul_fit(emul, ab, p, u, sx, max_iters=max_iters, flags=flags, wide=True)
        rc, best, allr, pred, resid = oracle.ab_neutral(oracle.Problem(p, u, u, 1.0), sx, max_iters=max_iters, flags=oflags,
                                                        n_threads=4)
        assert rc == 0
        for f in ("theta", "cost", "lse", "iters", "evals", "status"):
            assert np.array_equal(g[f], allr[f]), (flags, f)
            assert np.array_equal(gw[f], allr[f]), ("wide", flags, f)


def test_bootstrap_column_access_matches_oracle(emul, ab, oracle):
    """per-lane D* columns (the bootstrap kernel's access pattern) through the same objective"""
    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    u = 0.8
    sx = ab.gen_start_simplices(1, 0, 16, float(ped[:, 3].max()))
    rc, best, _, pred, resid = oracle.ab_neutral(oracle.Problem(ped, u, u, 1.0), sx,
                                                 flags=oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL, n_threads=4)
    n_boot = 12
    idx = ab.gen_resample_idx(5, 0, n_boot, len(ped))
    vary = ab.gen_vary_vertices(5, 0, n_boot, best["theta"])
    dstar = pred[None, :] + resid[idx]
    simplices = np.concatenate([np.broadcast_to(best["theta"], (n_boot, 1, 4)), vary], axis=1)
    g = emul_fit(emul, ab, ped, u, simplices, max_iters=1000, dstar=dstar)
    gw = emul_fit(emul, ab, ped, u, simplices, max_iters=1000, dstar=dstar, wide=True)
    assert np.array_equal(gw["theta"], g["theta"]) and np.array_equal(gw["evals"], g["evals"])
    rc, rows, fits = oracle.boot_model(oracle.Problem(ped, u, u, 1.0), best["theta"], pred, resid, idx, vary,
                                       flags=oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL, n_threads=4)
    assert rc == 0
    assert np.array_equal(g["theta"], fits["theta"]) and np.array_equal(g["evals"], fits["evals"])


def emul_shape(emul, ab, peds, fits_per_prob):
    arr = ab._pack_problems([ab.Problem(p, 0.8, 0.8, 1.0) for p in peds])
    out = (C.c_int64 * 10)()
    rc = emul.emul_launch_shape(arr, len(peds), fits_per_prob, out)
    assert rc == 0, rc
    keys = ("n_warps", "d_shared", "x_global", "big", "wide", "boot_x_global", "smem_fit", "smem_wide", "smem_boot_gather", "n_items")
    return dict(zip(keys, [int(v) for v in out]))


def test_launch_shapes_for_the_baseline_configs(emul, ab, oracle, monkeypatch):
    """host-side choice of kernel variant and block shape (abfit_plan.cu) for the shapes of BASELINE.json"""
    for k in ("ABFIT_DEV_NWARPS", "ABFIT_DEV_XGLOBAL", "ABFIT_DEV_BIG", "ABFIT_DEV_WIDE", "ABFIT_DEV_BOOT_XGLOBAL",
              "ABFIT_DEV_BOOT_TILE", "ABFIT_DEV_CHUNK"):
        monkeypatch.delenv(k, raising=False)
    ped351 = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    # C4: thousands of windows sharing one program -> 3-warp blocks, 4 per SM, one block per window; bootstrap with
    # the vertices in global scratch because there are more windows than resident bootstrap blocks
    c4 = emul_shape(emul, ab, [ped351] * 2000, 1000)
    assert (c4["n_warps"], c4["d_shared"], c4["x_global"], c4["big"], c4["wide"]) == (3, 1, 0, 0, 0)
    assert 4 * (c4["smem_fit"] + 1024) <= 228 * 1024 < 5 * (c4["smem_fit"] + 1024)
    assert c4["n_items"] == 2000 and c4["boot_x_global"] == 1 and 0 < c4["smem_boot_gather"] < 21 * 1024
    # a few windows: resident all at once, the bootstrap keeps its vertices in shared memory
    assert emul_shape(emul, ab, [ped351] * 500, 1000)["boot_x_global"] == 0
    # C1 / C2: one small pedigree x 1000 starts -> one-warp blocks spread over the SMs
    ped6 = np.loadtxt(os.path.join(GOLDEN, "pedigree_generated.txt"), skiprows=1)
    c1 = emul_shape(emul, ab, [ped6], 1000)
    assert c1["n_warps"] == 1 and c1["big"] == 0 and c1["n_items"] == 32
    # C5: 19 900 pairs, 590 triples -> per-lane state does not fit; warp-per-fit, one fit per block
    samples = [(l, g) for l in range(10) for g in range(1, 21)]
    rows = [[min(g1, g2) if l1 == l2 else 0, g1, g2, 0.1] for i, (l1, g1) in enumerate(samples) for (l2, g2) in samples[i + 1:]]
    c5 = emul_shape(emul, ab, [np.array(rows, dtype=np.float64)], 1000)
    assert c5["big"] == 1 and c5["wide"] == 1 and c5["n_items"] == 1000 and c5["smem_wide"] < 24 * 1024


# ---------------------------------------------------------------------------------------------
# run-time specialised objective (csrc/abfit_jit.cu): the GENERATED source, built for the CPU
# ---------------------------------------------------------------------------------------------
def build_spec(ab, ped, tmp_path, tag):
    """abfit_jit_dump -> generated CUDA C++ -> g++ (shims, -ffp-contract=off) -> ctypes"""
    src = os.path.join(str(tmp_path), f"spec_{tag}.cu")
    so = os.path.join(str(tmp_path), f"libspec_{tag}.so")
    ab.jit_dump(ab.Problem(ped, 0.8, 0.8, 1.0), source_path=src)
    csrc = os.path.join(ROOT, "alphabeta-rs_b200", "csrc")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                           "-I/usr/local/cuda/include", "-I" + csrc, "-I" + EMUL_DIR, "-Wno-attributes", "-Wno-unknown-pragmas",
                           f'-DSPEC_SOURCE="{src}"', "-include", os.path.join(EMUL_DIR, "shims.h"), "-o", so,
                           os.path.join(EMUL_DIR, "emul_spec.cpp")])
    return C.CDLL(so)


def spec_cost(lib, ab, ped, p0uu, eqp, w, theta):
    arr = ab._pack_problems([ab.Problem(ped, p0uu, eqp, w)])
    theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, 4)
    cost, lse = np.empty(len(theta)), np.empty(len(theta))
    assert lib.spec_cost(arr, theta.ctypes.data_as(C.c_void_p), len(theta), cost.ctypes.data_as(C.c_void_p),
                         lse.ctypes.data_as(C.c_void_p)) == 0
    return cost, lse


def spec_fit(lib, ab, emul, ped, p0uu, sx, max_iters=10000, flags=0, dstar=None):
    arr = ab._pack_problems([ab.Problem(ped, p0uu, p0uu, 1.0)])
    sx = np.ascontiguousarray(sx, dtype=np.float64)
    n = sx.size // 20
    out = np.zeros(n, dtype=ab.FIT_DTYPE)
    dp = None if dstar is None else np.ascontiguousarray(dstar, dtype=np.float64).ctypes.data_as(C.c_void_p)
    emul.emul_var_threshold.restype = C.c_double
    thr = emul.emul_var_threshold(C.c_double(ab.DBL_EPSILON))
    emul.emul_range_threshold.restype = C.c_double
    rthr = emul.emul_range_threshold(C.c_double(thr))
    assert lib.spec_fit(arr, sx.ctypes.data_as(C.c_void_p), n, dp, max_iters, C.c_double(ab.DBL_EPSILON), flags,
                        C.c_double(thr), C.c_double(rthr), out.ctypes.data_as(C.c_void_p)) == 0
    return out


@pytest.mark.parametrize("style,tmax,n_pairs", [("lineage", 33, 131), ("sibling", 33, 120), ("any", 4, 37), ("any", 127, 260),
                                                ("sibling", 1, 5), ("lineage", 4, 3)])
def test_specialised_objective_source_is_bit_exact(emul, ab, oracle, tmp_path, style, tmax, n_pairs):
    """the code generator against the oracle on time structures that exercise every micro-op (identity operands,
    G itself, stored powers, deferred d-vectors), pair counts that are not multiples of 4, negative / NaN thetas"""
    rng = np.random.default_rng(hash((style, tmax, n_pairs)) % 2**32)
    ped = random_pedigree(rng, n_pairs, tmax, style)
    lib = build_spec(ab, ped, tmp_path, f"{style}{tmax}_{n_pairs}")
    B = 60
    theta = np.stack([10 ** rng.uniform(-7, -1.5, B), 10 ** rng.uniform(-7, -1.5, B), rng.uniform(-0.1, 0.3, B),
                      rng.uniform(0, 0.1, B)], axis=1)
    theta[3] = [-2e-4, 3e-3, -0.2, 0.01]
    theta[4] = [0.0, 0.0, 0.0, 0.0]  # 0/0 in p_uu_est: NaN cost, finite LSE
    cost, lse = spec_cost(lib, ab, ped, 0.8, 0.7, 1.3, theta)
    pb = oracle.Problem(ped, 0.8, 0.7, 1.3)
    for i in range(B):
        wc = oracle.cost(pb, theta[i])
        assert cost[i] == wc or (np.isnan(cost[i]) and np.isnan(wc)), (style, tmax, n_pairs, i)
        assert lse[i] == oracle.lse(pb, theta[i], flags=oracle.FAST_DIVERGENCE), (style, tmax, n_pairs, i)
    icost, ilse = emul_cost(emul, ab, ped, 0.8, 0.7, 1.3, theta)  # and == the interpreter it replaces
    assert np.array_equal(icost, cost, equal_nan=True) and np.array_equal(ilse, lse)


def test_specialised_objective_golden_kat_and_fits(emul, ab, oracle, tmp_path):
    """the C4 pedigree (data/pedigree.txt): cost KAT of src/structs.rs:233, whole Nelder-Mead runs and bootstrap
    replicates through the generated objective, all bit-identical to the oracle"""
    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    lib = build_spec(ab, ped, tmp_path, "c4")
    cost, _ = spec_cost(lib, ab, ped, 0.75, 0.5, 0.7, [[0.0001179555, 0.0001180614, 0.03693534, 0.003023981]])
    assert cost[0] == 0.0006700888539608879
    u = 0.8
    sx = ab.gen_start_simplices(7, 3, 24, float(ped[:, 3].max()))
    fl = oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL
    g = spec_fit(lib, ab, emul, ped, u, sx)
    rc, best, allr, pred, resid = oracle.ab_neutral(oracle.Problem(ped, u, u, 1.0), sx, flags=fl, n_threads=4)
    assert rc == 0
    for f in ("theta", "cost", "lse", "iters", "evals", "status"):
        assert np.array_equal(g[f], allr[f]), f
    n_boot = 10
    idx = ab.gen_resample_idx(5, 0, n_boot, len(ped))
    vary = ab.gen_vary_vertices(5, 0, n_boot, best["theta"])
    simplices = np.concatenate([np.broadcast_to(best["theta"], (n_boot, 1, 4)), vary], axis=1)
    gb = spec_fit(lib, ab, emul, ped, u, simplices, max_iters=1000, dstar=pred[None, :] + resid[idx])
    rc, rows, fits = oracle.boot_model(oracle.Problem(ped, u, u, 1.0), best["theta"], pred, resid, idx, vary, flags=fl,
                                       n_threads=4)
    assert rc == 0 and np.array_equal(gb["theta"], fits["theta"]) and np.array_equal(gb["evals"], fits["evals"])


def test_specialised_kernels_compile_for_sm_100a_without_a_gpu(ab, oracle, tmp_path, monkeypatch):
    """NVRTC cross-compiles the generated source + the embedded device headers here; the cubin holds both kernels"""
    monkeypatch.setenv("ABFIT_CACHE_DIR", str(tmp_path / "cache"))
    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    cubin = str(tmp_path / "spec.cubin")
    secs = ab.jit_dump(ab.Problem(ped, 0.8, 0.8, 1.0), cubin_path=cubin)
    assert secs > 0 and os.path.getsize(cubin) > 10000
    blob = open(cubin, "rb").read()
    assert b"abfit_jit_fit_starts" in blob and b"abfit_jit_fit_boot_gather" in blob
    # second request: served from the disk cache
    assert ab.jit_dump(ab.Problem(ped, 0.8, 0.8, 1.0), cubin_path=cubin) == 0.0
