"""Input side: the product's host parsers and pedigree builder against the oracle's Python
restatement, which is pinned on the reference's parser tests (src/methylation_site.rs:517-593) and golden files."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, resolve_golden

STRAND = {"+": 1, "-": -1, "*": 0}


def same_site(got, want, oracle):
    if want is None:
        return got is None
    if got is None:
        return False
    return (got["chromosome"] == oracle.chromosome_id(str(want["chromosome"])) and got["start"] == want["start"]
            and got["end"] == want["end"] and got["strand"] == STRAND[want["strand"]]
            and got["posteriormax"] == want["posteriormax"] and got["status"] == want["status"] and got["meth_lvl"] == want["meth_lvl"])


def test_reference_parser_cases(ab, oracle):
    # src/methylation_site.rs:517-593
    ok9 = "1\t23151\t+\tCG\t0\t8\t0.9999\tU\t0.0025"
    s = ab.parse_methylome_line(ok9)
    assert s == {"chromosome": 1, "start": 23151, "end": 23152, "strand": 1, "posteriormax": 0.9999, "status": 0, "meth_lvl": 0.0025}
    assert ab.parse_methylome_line(ok9, invert_strand=True)["strand"] == -1
    assert ab.parse_methylome_line("1\t23151\t+\tCHH\t0\t8\t0.9999\tU\t0.0025") is None  # CHH skipped
    assert ab.parse_methylome_line("X\t23151\t+\tCG\t0\t8\t0.9999\tU\t0.0025") is None    # chromosome X rejected
    assert ab.parse_methylome_line("chr1\t1\t4\t1") == {"chromosome": 1, "start": 1, "end": 4, "strand": 0, "posteriormax": 0.0,
                                                          "status": 0, "meth_lvl": 0.0}  # bedGraph
    for bad in ("", "seqnames\tstart\tstrand\tcontext", "1\t-5\t+\tCG\t0\t8\t0.9\tU\t0.1", "1\t5\t+\tCG\t0\t8\t0x1p3\tU\t0.1",
                "1\t5\t+\tCG\t0\t8\t 0.9\tU\t0.1", "1\t5\t+\tCG\t0\t8\t0.9\t\t0.1", "1\t4294967296\t+\tCG\t0\t8\t0.9\tU\t0.1"):
        assert ab.parse_methylome_line(bad) is None and oracle.parse_methylome_line(bad) is None, bad


def test_every_golden_methylome_line_parses_like_the_oracle(ab, oracle):
    n = 0
    for d in ("methylome", "desired_output"):
        for name in sorted(os.listdir(os.path.join(GOLDEN, d))):
            if not (name.endswith(".txt") and ("methylome" in name or d == "methylome")):
                continue
            cg = 0
            for line in open(os.path.join(GOLDEN, d, name)).read().split("\n"):
                want = oracle.parse_methylome_line(line)
                assert same_site(ab.parse_methylome_line(line), want, oracle), (name, line)
                cg += want is not None
                n += 1
            if d == "methylome":
                assert cg == 500  # src/methylation_site.rs:553-593
    assert n > 10000


@pytest.mark.gpu
def test_pedigree_build_matches_oracle_and_golden(ab, ctx, oracle, monkeypatch, tmp_path):
    """C1: data/nodelist.txt + edgelist.txt -> data/pedigree_generated.txt (bit-exact), and the 13-sample desired_output set"""
    # the nodelists name their files relative to the reference's repo root / data dir: recreate that layout
    os.makedirs(os.path.join(tmp_path, "data"))
    os.symlink(os.path.join(GOLDEN, "methylome"), os.path.join(tmp_path, "data", "methylome"))
    for f in os.listdir(os.path.join(GOLDEN, "desired_output")):
        os.symlink(os.path.join(GOLDEN, "desired_output", f), os.path.join(tmp_path, f))
    monkeypatch.chdir(tmp_path)
    ped, p0uu, info = ab.build_pedigree(ctx, os.path.join(GOLDEN, "nodelist.txt"), os.path.join(GOLDEN, "edgelist.txt"), 0.99)
    want, want_p0, _ = oracle.build_pedigree(os.path.join(GOLDEN, "nodelist.txt"), os.path.join(GOLDEN, "edgelist.txt"), 0.99,
                                             resolve_golden)
    assert np.array_equal(ped, want) and p0uu == want_p0 and info["n_samples"] == 4 and info["n_sites"] == 500
    out = os.path.join(tmp_path, "pedigree.txt")
    ab.write_pedigree(out, ped)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, "pedigree_generated.txt"), "rb").read()
    ped2, p02, info2 = ab.build_pedigree(ctx, os.path.join(GOLDEN, "desired_output", "nodelist.fn"),
                                         os.path.join(GOLDEN, "desired_output", "edgelist.fn"), 0.99)
    want2, want_p02, _ = oracle.build_pedigree(os.path.join(GOLDEN, "desired_output", "nodelist.fn"),
                                               os.path.join(GOLDEN, "desired_output", "edgelist.fn"), 0.99, resolve_golden)
    assert ped2.shape == (78, 4) and np.array_equal(ped2, want2) and p02 == want_p02


@pytest.mark.gpu
def test_alphabeta_cli_end_to_end(ab, ctx, oracle, tmp_path):
    """the `alphabeta` binary on the repo's example (BASELINE configs[0] / configs[1] with 200 iterations): same files as the
    reference's formats, numbers identical to the oracle pipeline fed the same seeded inputs"""
    os.makedirs(os.path.join(tmp_path, "data"))
    os.symlink(os.path.join(GOLDEN, "methylome"), os.path.join(tmp_path, "data", "methylome"))
    outdir = os.path.join(tmp_path, "out")
    os.makedirs(outdir)
    exe = os.path.join(ROOT, "alphabeta-rs_b200", "alphabeta")
    n = 200
    r = subprocess.run([exe, "-n", os.path.join(GOLDEN, "nodelist.txt"), "-e", os.path.join(GOLDEN, "edgelist.txt"), "-i", str(n),
                        "-o", outdir, "--seed", "12345"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert open(os.path.join(outdir, "pedigree.txt"), "rb").read() == open(os.path.join(GOLDEN, "pedigree_generated.txt"), "rb").read()
    # the oracle pipeline on the same seeded inputs
    ped, p0uu, _ = oracle.build_pedigree(os.path.join(GOLDEN, "nodelist.txt"), os.path.join(GOLDEN, "edgelist.txt"), 0.99, resolve_golden)
    sx = ab.gen_start_simplices(12345, 0, n, float(ped[:, 3].max()))
    flags = oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL
    rc, best, _, pred, resid = oracle.ab_neutral(oracle.Problem(ped, p0uu, p0uu, 1.0), sx, flags=flags, n_threads=8)
    idx = ab.gen_resample_idx(12345, 0, n, len(ped))
    vary = ab.gen_vary_vertices(12345, 0, n, best["theta"])
    rc2, rows, _ = oracle.boot_model(oracle.Problem(ped, p0uu, p0uu, 1.0), best["theta"], pred, resid, idx, vary, flags=flags, n_threads=8)
    assert rc == 0 and rc2 == 0
    assert np.array_equal(np.load(os.path.join(outdir, "raw.npy")), rows)
    assert open(os.path.join(outdir, "analysis.txt")).read() == oracle.analysis_file_text(oracle.analyze(rows))
    want_block = ("##########\nResults:\n\nModel:\n\tAlpha: %s\n\tBeta: %s\n\tWeight: %s\n\tIntercept: %s\n" % tuple(
        oracle.rust_f64(float(x)) for x in best["theta"])) + oracle.analysis_file_text(oracle.analyze(rows)) + "\n" + \
        "Estimated steady state %s\nObserved steady state methylation %s\n##########\n" % (
            oracle.rust_f64(oracle.steady_state(best["theta"][0], best["theta"][1])), oracle.rust_f64(1.0 - p0uu))
    assert want_block in r.stdout
