"""Result files: the product's host writers against a literal restatement of the
reference's writers and against the one golden file the current reference code writes itself
(data/pedigree_generated.txt, src/pedigree.rs:345-358)."""
import os
import struct

import numpy as np

from conftest import GOLDEN


def test_rust_f64_display(ab, oracle):
    kat = [(0.0, "0"), (-0.0, "-0"), (1.0, "1"), (4.0, "4"), (0.1, "0.1"), (0.10931174089068826, "0.10931174089068826"),
           (1e-7, "0.0000001"), (1e21, "1000000000000000000000"), (123456.789, "123456.789"), (float("nan"), "NaN"),
           (float("inf"), "inf"), (float("-inf"), "-inf"), (5e-324, "0." + "0" * 323 + "5"), (2.5e-5, "0.000025"),
           (1.7976931348623157e308, "17976931348623157" + "0" * 292), (-1.5, "-1.5")]
    for v, want in kat:
        assert oracle.rust_f64(v) == want, v
        assert ab.format_f64(v) == want, v
    rng = np.random.default_rng(3)
    vals = np.concatenate([rng.normal(size=300), 10 ** rng.uniform(-300, 300, 300), -10 ** rng.uniform(-20, 20, 100),
                           rng.integers(-10**15, 10**15, 100).astype(np.float64)])
    for v in vals:
        assert ab.format_f64(float(v)) == oracle.rust_f64(float(v)), v
        assert float(ab.format_f64(float(v))) == float(v)  # round trip


def test_pedigree_file_is_byte_identical_to_the_reference_output(ab, oracle, tmp_path):
    ped = oracle.load_pedigree_file(os.path.join(GOLDEN, "pedigree_generated.txt")) if False else \
        np.loadtxt(os.path.join(GOLDEN, "pedigree_generated.txt"), skiprows=1)
    golden = open(os.path.join(GOLDEN, "pedigree_generated.txt"), "rb").read()
    assert oracle.pedigree_file_text(ped).encode() == golden
    path = os.path.join(tmp_path, "pedigree.txt")
    ab.write_pedigree(path, ped)
    assert open(path, "rb").read() == golden


def test_analysis_npy_and_results_files(ab, oracle, tmp_path):
    rng = np.random.default_rng(8)
    rows = np.abs(rng.normal(1e-3, 2e-4, (200, 7)))
    an = ab.analyze(rows)
    assert np.array_equal(an, oracle.analyze(rows))
    p = os.path.join(tmp_path, "analysis.txt")
    ab.write_analysis(p, an)
    assert open(p).read() == oracle.analysis_file_text(an) == ab.format_analysis(an)
    assert open(p).read().count("\n") == 24
    # raw.npy: readable by numpy, '<f8', C order, 64-byte aligned header
    for arr in (rows, rng.normal(size=(5, 7, 3)), np.zeros((0, 7))):
        q = os.path.join(tmp_path, "raw.npy")
        ab.write_npy(q, arr)
        raw = open(q, "rb").read()
        assert raw[:8] == b"\x93NUMPY\x01\x00" and (10 + struct.unpack("<H", raw[8:10])[0]) % 64 == 0
        back = np.load(q)
        assert back.dtype == np.dtype("<f8") and back.shape == arr.shape and np.array_equal(back, arr)
    # metaprofile results.txt
    n = 6
    best = np.zeros(n, dtype=ab.FIT_DTYPE)
    best["theta"][:, 0] = 10 ** rng.uniform(-5, -3, n)
    best["theta"][:, 1] = 10 ** rng.uniform(-4, -2, n)
    ans = np.stack([ab.analyze(np.abs(rng.normal(1e-3, 2e-4, (50, 7)))) for _ in range(n)])
    cg = rng.integers(0, 1000, n)
    region = np.array([0, 0, 1, 1, 2, 2])
    obs = rng.uniform(0, 1, n)
    obs[3] = np.nan
    r = os.path.join(tmp_path, "results.txt")
    ab.write_metaprofile_results(r, "run 1", cg, region, best, ans, obs)
    assert open(r).read() == oracle.metaprofile_results_text("run 1", cg, region, best["theta"], ans, obs)
    for i in range(n):
        assert ab.steady_state(best["theta"][i, 0], best["theta"][i, 1]) == oracle.steady_state(best["theta"][i, 0], best["theta"][i, 1])
