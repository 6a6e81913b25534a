"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/abfit.h declares, and refuses to compute without a CUDA device (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "abfit.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(abfit_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol(ab):
    lib = ctypes.CDLL(ab.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/abfit.h but not exported"
    assert sorted(ab.EXPORTED_SYMBOLS) == syms


def test_version(ab):
    assert "sm_100a" in ab.version()


def test_generators_are_deterministic_and_in_range(ab):
    a = ab.gen_start_simplices(1, 2, 50, 0.02)
    b = ab.gen_start_simplices(1, 2, 50, 0.02)
    assert a.shape == (50, 5, 4) and np.array_equal(a, b)
    assert not np.array_equal(a, ab.gen_start_simplices(1, 3, 50, 0.02))
    # Model::new (src/structs.rs:78-96)
    assert np.all((a[..., 0] >= 1e-9) & (a[..., 0] < 1e-2)) and np.all((a[..., 1] >= 1e-9) & (a[..., 1] < 1e-2))
    assert np.all((a[..., 2] >= 0) & (a[..., 2] < 0.1)) and np.all((a[..., 3] >= 0) & (a[..., 3] < 0.02))
    # a prefix of a longer draw is the same draw (counter based: shards can be generated independently)
    assert np.array_equal(ab.gen_start_simplices(1, 2, 80, 0.02)[:50], a)
    # max_divergence <= 0 -> 0.1 (src/structs.rs:80-83)
    z = ab.gen_start_simplices(1, 2, 50, 0.0)
    assert np.all(z[..., 3] < 0.1) and z[..., 3].max() > 0.05
    # Model::vary (src/structs.rs:100-128)
    th = np.array([1e-4, -2e-3, 0.0, 0.5])
    v = ab.gen_vary_vertices(9, 0, 64, th)
    assert v.shape == (64, 4, 4)
    assert np.all(np.abs(v[..., 0] - 1e-4) <= 1e-5) and np.all(np.abs(v[..., 1] + 2e-3) <= 2e-4)
    assert np.all((v[..., 2] >= 0.09) & (v[..., 2] <= 0.11)) and np.all(np.abs(v[..., 3] - 0.5) <= 0.05)
    idx = ab.gen_resample_idx(9, 0, 16, 351)
    assert idx.shape == (16, 351) and idx.min() >= 0 and idx.max() < 351 and len(np.unique(idx)) > 300


def test_analyze_matches_oracle(ab, oracle):
    rng = np.random.default_rng(11)
    rows = np.abs(rng.normal(1.0, 0.2, (100, 7)))
    assert np.array_equal(ab.analyze(rows), oracle.analyze(rows))
    rows = np.abs(rng.normal(1.0, 0.2, (1000, 7)))
    assert np.array_equal(ab.analyze(rows), oracle.analyze(rows))


def test_no_cpu_fallback(ab):
    """Without a GPU the library must fail loudly, not compute on the host."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(ab.AbfitError) as e:
        ab.Context(0)
    assert e.value.code == ab.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under the product package may reference it"""
    pkg = os.path.join(ROOT, "alphabeta-rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "abref" not in txt and "oracle/" not in txt.replace("oracle/.", ""), os.path.join(dirpath, f)


def test_python_twin_of_the_seeded_generators_is_bit_identical(ab):
    """oracle/gen_py.py feeds the reference arm of bench.py (so that process never maps libabfit.so); the same seeds
    must give the same doubles / indices as the C generators the GPU arm uses (north_star: identical inputs)"""
    from oracle import gen_py as g

    for pid in (0, 5, 9999, 12345678901):
        for md in (0.0123, 0.0, -1.0):
            assert ab.gen_start_simplices(0xAB0B200, pid, 130, md).tobytes() == g.gen_start_simplices(0xAB0B200, pid, 130, md).tobytes()
        for n in (6, 351):
            assert ab.gen_resample_idx(0xAB0B200, pid, 33, n).tobytes() == g.gen_resample_idx(0xAB0B200, pid, 33, n).tobytes()
        for th in ([1e-4, 2e-3, 0.05, 0.01], [-1e-4, 0.0, -0.3, 0.0]):
            assert ab.gen_vary_vertices(0xAB0B200, pid, 21, th).tobytes() == g.gen_vary_vertices(0xAB0B200, pid, 21, th).tobytes()


def test_literal_sort_picks_the_same_winner(oracle):
    """ABREF_LITERAL_SORT (the reference's sort_by with the re-evaluating comparator, src/ab_neutral.rs:83-101) and the
    direct stable argmin agree on the winner; the flag only adds the reference's serial work"""
    import os
    from conftest import GOLDEN
    from oracle import gen_py as g

    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree.txt"))
    pb = oracle.Problem(ped, 0.8, 0.8, 1.0)
    sx = g.gen_start_simplices(7, 0, 45, float(ped[:, 3].max()))
    fl = oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL
    a = oracle.ab_neutral(pb, sx, flags=fl, n_threads=4)
    b = oracle.ab_neutral(pb, sx, flags=fl | oracle.LITERAL_SORT, n_threads=4)
    assert a[0] == b[0] == 0 and a[1]["start_id"] == b[1]["start_id"] and np.array_equal(a[3], b[3])


def test_rust_sys_crate_names_only_declared_entry_points():
    """rust/abfit-sys (untested courtesy binding: no Rust toolchain here) must at least name real entry points"""
    src = open(os.path.join(ROOT, "rust", "abfit-sys", "src", "lib.rs")).read()
    rust = set(re.findall(r"pub fn (abfit_[a-z0-9_]+)\s*\(", src))
    assert len(rust) >= 20 and rust <= set(declared_symbols()), sorted(rust - set(declared_symbols()))
