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


def _fuzz_lines(rng, n):
    """methylome-like lines with the grammar's corner cases mixed in (numbers Rust accepts and rejects, field counts of
    all six formats, separators, prefixes)"""
    nums = ["0", "1", "0.9999", ".5", "5.", "+0.25", "-0.0", "1e-3", "1E+2", "1e", "e5", "0x10", "inf", "-Infinity", "NaN", "nan ",
            "1e999", "1e-999", "4.9e-324", "0.1234567890123456789", "12345678901234567890", "1_0", "", " 1", "1.2.3", "+", "-",
            "+.e1", "9007199254740993", "0.30000000000000004", "2.2250738585072011e-308"]
    ints = ["0", "7", "+12", "-3", "4294967295", "4294967296", "00012", "1.0", "", "12a"]
    chroms = ["1", "5", "chr2", "chrchr3", "chr", "M", "C", "chrM", "X", "255", "256", "0", "+4", "01"]
    ctx = ["CG", "CHH", "CHG", "cg", "CG "]
    out = []
    for _ in range(n):
        k = rng.integers(0, 6)
        pick = lambda a: a[rng.integers(0, len(a))]
        if k == 0:
            f = [pick(chroms), pick(ints), pick(["+", "-", "*", ""]), pick(ctx), pick(ints), pick(ints), pick(nums), pick(["U", "M", "I", "", "Q", "MU"]), pick(nums)]
        elif k == 1:
            f = [pick(chroms), pick(ints), pick(["+", "-"]), pick(ctx), pick(ints), pick(ints), pick(nums), pick(["U", "M", "I"]), pick(nums), "CGA"]
        elif k == 2:
            f = [pick(chroms), pick(ints), pick(ints), "x", "y", pick(["+", "-"]), pick(ints), pick(ints), pick(nums), pick(["U", "M", "I"]), pick(nums)]
            f[3] = pick(ctx)
        elif k == 3:
            f = [pick(chroms), pick(ints), pick(ints), pick(nums)]
        elif k == 4:
            f = [pick(chroms) for _ in range(rng.integers(0, 14))]
        else:
            f = [pick(chroms), pick(ints), "+", "CG", "1", "2", "0.5", "U", "0.5"]
        sep = "\t" if rng.random() < 0.85 else pick([" ", "\t\t", ","])
        out.append(sep.join(f))
    return out


def test_parser_fuzz_against_the_oracle(ab, oracle):
    rng = np.random.default_rng(4242)
    lines = _fuzz_lines(rng, 6000)
    accepted = 0
    for line in lines:
        want = oracle.parse_methylome_line(line)
        got = ab.parse_methylome_line(line)
        if want is None:
            assert got is None, line
            continue
        accepted += 1
        assert got is not None, line
        for k in ("start", "end", "status"):
            assert got[k] == want[k], (line, k)
        for k in ("posteriormax", "meth_lvl"):  # same bits, NaN included
            assert np.float64(got[k]).tobytes() == np.float64(want[k]).tobytes() or (np.isnan(got[k]) and np.isnan(want[k])), (line, k)
    assert 500 < accepted < 5500


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_pedigree_graph_random_trees_against_the_oracle(ab, oracle, tmp_path, seed):
    """abfit_pedigree_graph (nodelist / edgelist parsing, shortest path, t0 rule: src/pedigree.rs:99-135,264-337) on
    random lineage trees — branching founders, unmeasured intermediate nodes, generation gaps, separators mixed, an
    isolated measured node, edges naming unknown nodes — against the oracle's restatement: same measured files, same
    (i, j, t0, t1, t2) rows in the same order"""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(4, 40))
    gen = [0]
    parent = [-1]
    for k in range(1, n):
        q = int(rng.integers(0, k))
        parent.append(q)
        gen.append(gen[q] + int(rng.integers(1, 4)))
    meth = [rng.random() < 0.6 for _ in range(n)]
    meth[-1] = meth[-2] = True
    sep = lambda: ["\t", ",", " "][int(rng.integers(0, 3))]
    lines = ["filename,node,gen,meth"]
    for k in range(n):
        s1 = sep()
        lines.append(s1.join([f"/data/m_{k}.txt" if meth[k] else "-", f"N{k}", str(gen[k]), "Y" if meth[k] else "N"]))
        if rng.random() < 0.1:
            lines.append("")  # blank lines count towards the ids
    lines.append("\t".join([f"/data/m_lonely.txt", "LONELY", "7", "Y"]))  # measured, but no edge reaches it
    lines.append("bad line")
    edges = ["from,to"]
    for k in range(1, n):
        a, b = (f"N{parent[k]}", f"N{k}") if rng.random() < 0.7 else (f"N{k}", f"N{parent[k]}")
        edges.append(sep().join([a, b]))
    edges.append("N0,GHOST")
    nl, el = os.path.join(tmp_path, "nodes.txt"), os.path.join(tmp_path, "edges.txt")
    open(nl, "w").write("\n".join(lines) + "\n")
    open(el, "w").write("\n".join(edges) + "\n")
    files, pairs = ab.pedigree_graph(nl, el)
    nodes = oracle.parse_nodelist(open(nl).read())
    meas = [x for x in nodes if x["meth"]]
    want = oracle.pedigree_pairs(meas, oracle.parse_edgelist(open(el).read(), nodes))
    assert files == [x["file"] for x in meas]
    assert len(want) > 0 and pairs.shape == (len(want), 5)
    assert np.array_equal(pairs, np.array(want, dtype=np.float64))


def test_annotation_parser_fuzz_against_the_oracle(ab, oracle):
    """Gene::from_annotation_file_line (src/genes.rs:166-216): every golden annotation line, and 4000 lines with the
    field count, the separators, the strand column and the numbers perturbed, product == oracle restatement"""
    rng = np.random.default_rng(515)
    golden = [l for l in open(os.path.join(GOLDEN, "annotation.bed")).read().split("\n")]
    chroms = ["1", "5", "chr2", "chrchr3", "chr", "M", "C", "chrM", "X", "255", "256", "+4", ""]
    ints = ["0", "7", "+12", "-3", "4294967295", "4294967296", "00012", "1.0", "", "12a", "300", "90000"]
    strands = ["+", "-", "*", ".", "", "++", "sense"]
    lines = list(golden)
    for _ in range(4000):
        pick = lambda a: a[rng.integers(0, len(a))]
        k = rng.integers(0, 4)
        if k == 0:
            f = [pick(chroms), pick(ints), pick(ints), "AT1G01", "gbM", pick(strands)]
        elif k == 1:
            f = [pick(chroms), pick(ints), pick(ints), pick(ints), pick(strands), "AT1G01"]
        elif k == 2:
            f = [pick(chroms + ints + strands) for _ in range(rng.integers(0, 9))]
        else:
            f = golden[rng.integers(0, len(golden) - 1)].split("\t")
            if f and rng.random() < 0.5:
                f[rng.integers(0, len(f))] = pick(chroms + ints + strands)
        lines.append(pick(["\t", "\t", " ", ",", "\t "]).join(f))
    n_ok = 0
    for line in lines:
        for inv in (False, True):
            want = oracle.parse_annotation_line(line, inv)
            got = ab.parse_annotation_line(line, inv)
            assert got == (None if want is None else tuple(int(x) for x in want)), (line, inv, got, want)
        n_ok += want is not None
    assert n_ok > len(golden) - 5 and n_ok < len(lines) - 1000


def test_parser_f64_is_correctly_rounded(ab):
    """<f64 as FromStr> is correctly rounded; so is Python's float().  20 000 random decimal strings (up to 25
    significant digits, exponents across the whole range incl. subnormals, overflow to inf and underflow to 0) through
    the whole-file parser, compared bit for bit"""
    rng = np.random.default_rng(99)
    texts = []
    for _ in range(20000):
        nd = int(rng.integers(1, 26))
        digits = "".join(str(int(d)) for d in rng.integers(0, 10, nd))
        k = int(rng.integers(0, nd + 1))
        mant = digits[:k] + "." + digits[k:] if rng.random() < 0.8 else digits
        if mant.startswith("."):
            mant = mant if rng.random() < 0.5 else "0" + mant
        form = rng.integers(0, 4)
        if form == 0:
            t = mant
        elif form == 1:
            t = mant + "e" + str(int(rng.integers(-30, 30)))
        elif form == 2:
            e = int(rng.integers(-345, 320))
            t = mant + "E" + ("+" if e >= 0 and rng.random() < 0.3 else "") + str(e)
        else:
            t = mant + "e-" + str(int(rng.integers(290, 330)))
        texts.append(("-" if rng.random() < 0.2 else "+" if rng.random() < 0.1 else "") + t)
    data = "\n".join(f"1\t{i + 1}\t+\tCG\t0\t8\t{t}\tU\t{texts[-1 - i]}" for i, t in enumerate(texts)).encode()
    got = ab.parse_methylome_buffer(data)
    assert len(got["sites"]) == len(texts)
    want = np.array([float(t) for t in texts])
    assert got["posteriormax"].tobytes() == want.tobytes()
    assert got["meth_lvl"].tobytes() == want[::-1].tobytes()
    assert np.isinf(want).any() and (want == 0).any() and ((np.abs(want) < 2.3e-308) & (want != 0)).any()


@pytest.mark.parametrize("newline,trailing,threads", [("\n", True, None), ("\r\n", True, "1"), ("\n", False, "3"), ("\r\n", False, None)])
def test_methylome_buffer_parser_equals_the_line_parser(ab, oracle, monkeypatch, newline, trailing, threads):
    """abfit_parse_methylome_buffer (whole file image, several threads, chunk cuts at line ends) == the per-line parser
    applied to BufRead::lines of the same bytes: golden files, a 12 MB fuzz image, header skipping, CRLF, missing final
    newline, empty input"""
    if threads:
        monkeypatch.setenv("ABFIT_HOST_THREADS", threads)
    rng = np.random.default_rng(77)
    golden = open(os.path.join(GOLDEN, "methylome", "G0.txt")).read().split("\n")
    golden = [l for l in golden if l]
    fuzz = _fuzz_lines(rng, 3000)
    big = (golden + fuzz) * 40  # ~ 230 000 lines, > 8 MB: several chunks per thread
    for lines, skip in ((golden, True), (golden, False), (fuzz, False), (big, True), ([], False), ([""], True), (["1\t5\t+\tCG\t0\t8\t0.9\tU\t0.1"], False)):
        data = newline.join(lines) + (newline if trailing and lines else "")
        got = ab.parse_methylome_buffer(data.encode(), skip_first_line=skip, want_lines=True)
        want, want_lines = [], []
        for i, line in enumerate(lines):
            if skip and i == 0:
                continue
            s = ab.parse_methylome_line(line)
            if s is not None:
                want.append(s)
                want_lines.append(line)
        assert len(got["sites"]) == len(want)
        if not want:
            continue
        assert np.array_equal(got["sites"]["chromosome"], [w["chromosome"] for w in want])
        assert np.array_equal(got["sites"]["start"], [w["start"] for w in want])
        assert np.array_equal(got["sites"]["end"], [w["end"] for w in want])
        assert np.array_equal(got["sites"]["strand"], [w["strand"] for w in want])
        assert np.array_equal(got["status"], [w["status"] for w in want])
        assert np.array_equal(got["posteriormax"], np.array([w["posteriormax"] for w in want]), equal_nan=True)
        assert np.array_equal(got["meth_lvl"], np.array([w["meth_lvl"] for w in want]), equal_nan=True)
        raw = data.encode()
        for j in (0, len(want) // 2, len(want) - 1):
            o, n = int(got["line_off"][j]), int(got["line_len"][j])
            assert raw[o:o + n].decode() == want_lines[j]
    # inverted strand, and a capacity that is too small
    data = "\n".join(golden).encode()
    inv = ab.parse_methylome_buffer(data, invert_strand=True, skip_first_line=True)
    fwd = ab.parse_methylome_buffer(data, skip_first_line=True)
    assert np.array_equal(inv["sites"]["strand"], -fwd["sites"]["strand"]) and len(fwd["sites"]) == 500


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
                        "-o", outdir, "--seed", "12345"], cwd=tmp_path, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, ABFIT_PROGRESS="1"))
    assert r.returncode == 0, r.stdout + r.stderr
    # the two bars of progress::specific (src/progress.rs:5-23) on stderr, finished at n / n
    assert "ABNeutral" in r.stderr and "BootModel" in r.stderr and f"{n:7d}/{n:<7d}" in r.stderr
    # bootstrap.png (src/boot_model.rs:105-109)
    from test_plots import decode_png
    img, _ = decode_png(os.path.join(outdir, "bootstrap.png"))
    assert img.shape == (960, 1280, 3) and ((img[:, :, 0] == 255) & (img[:, :, 1] == 0)).sum() > 100
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


def test_metaprofile_extraction_files_match_the_reference_output(tmp_path):
    """C3 on the shipped data: `metaprofile -m data/methylome -g data/annotation.bed -w 5` -> data/output_metaplot/*
    (no GPU involved: binning, distributions and steady-state files are host work)"""
    exe = os.path.join(ROOT, "alphabeta-rs_b200", "metaprofile")
    r = subprocess.run([exe, "-m", os.path.join(GOLDEN, "methylome"), "-g", os.path.join(GOLDEN, "annotation.bed"), "-w", "5",
                        "-o", str(tmp_path), "--name", "t"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "Done" in r.stdout, r.stdout + r.stderr
    gold = os.path.join(GOLDEN, "output_metaplot")
    for f in ("distribution_G0.txt", "steady_state_methylation.txt"):
        assert open(os.path.join(tmp_path, f), "rb").read() == open(os.path.join(gold, f), "rb").read(), f
    # distributions.txt lists the files in directory order in the reference, in name order here
    assert sorted(open(os.path.join(tmp_path, "distributions.txt")).read().split("\n")) == \
        sorted(open(os.path.join(gold, "distributions.txt")).read().split("\n"))


@pytest.mark.gpu
@pytest.mark.parametrize("ragged", [False, True])
def test_metaprofile_cli_fused_pipeline(ab, ctx, oracle, tmp_path, ragged):
    """`metaprofile ... alphabeta` on the example methylomes with an annotation that covers them: results.txt and raw.npy
    equal the reference's per-window loop restated with the oracle (same seeds).  ragged: two CG lines are taken out of
    one sample's file, so some windows list different numbers of sites per sample — the reference keeps those windows
    with D = 0 for every pair of unequal length (DMatrix::from, src/pedigree.rs:222-230)"""
    meth_dir = os.path.join(GOLDEN, "methylome")
    if ragged:
        meth_dir = os.path.join(tmp_path, "methylome")
        os.makedirs(meth_dir)
        for name in sorted(os.listdir(os.path.join(GOLDEN, "methylome"))):
            lines = open(os.path.join(GOLDEN, "methylome", name)).read().split("\n")
            if name == "G1_2.txt":
                cg = [i for i, l in enumerate(lines) if oracle.parse_methylome_line(l) is not None]
                keep = set(range(len(lines))) - {cg[len(cg) // 3], cg[2 * len(cg) // 3]}
                lines = [l for i, l in enumerate(lines) if i in keep]
            open(os.path.join(meth_dir, name), "w").write("\n".join(lines))
    ann = os.path.join(tmp_path, "ann.bed")
    genes = [(1, 300, 700, "-"), (1, 250, 800, "+"), (2, 1200, 1600, "+"), (2, 1250, 1650, "-"), (3, 600, 900, "*"),
             (4, 1200, 1550, "-"), (4, 1150, 1600, "+"), (5, 300, 900, "*"), ("C", 200, 900, "+"), ("C", 150, 950, "-"),
             ("M", 200, 600, "-"), ("M", 150, 650, "+")]
    open(ann, "w").write("".join(f"{c}\t{s}\t{e}\tg{i}\tgbM\t{sd}\n" for i, (c, s, e, sd) in enumerate(genes)))
    out = os.path.join(tmp_path, "out")
    os.makedirs(out)
    exe = os.path.join(ROOT, "alphabeta-rs_b200", "metaprofile")
    n_it, seed = 60, 777
    r = subprocess.run([exe, "-m", meth_dir, "-g", ann, "-w", "10", "-s", "5", "-c", "200", "-o", out,
                        "--name", "run7", "--iterations", str(n_it), "--seed", str(seed), "alphabeta",
                        "--nodes", os.path.join(GOLDEN, "nodelist.txt"), "--edges", os.path.join(GOLDEN, "edgelist.txt")],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # ---- the reference's flow, restated: per window, pedigree from that window's sites, fit, bootstrap -------------
    ped6, _, info = oracle.build_pedigree(os.path.join(GOLDEN, "nodelist.txt"), os.path.join(GOLDEN, "edgelist.txt"), 0.99, resolve_golden)
    og = [(oracle.chromosome_id(str(c)), s, e, STRAND[sd]) for c, s, e, sd in genes]
    # every sample's file is binned on its own (Windows::extract per file); the files list the same CG positions in
    # different orders, so a window holds different rows of each file
    per_sample = []
    for node in info["nodes"]:
        path = os.path.join(meth_dir, os.path.basename(resolve_golden(node["file"])))
        sites, st, po, me = [], [], [], []
        for line in open(path).read().split("\n")[1:]:
            q = oracle.parse_methylome_line(line)
            if q is not None:
                sites.append((oracle.chromosome_id(str(q["chromosome"])), q["start"], q["end"], STRAND[q["strand"]]))
                st.append(q["status"]); po.append(q["posteriormax"]); me.append(q["meth_lvl"])
        d, assign = oracle.extract_windows(og, sites, window_size=10, window_step=5, cutoff=200)
        per_sample.append((d, assign, np.array(st, dtype=np.uint8), np.array(po), np.array(me)))
    dist = per_sample[0][0]  # G0.txt: first node and first file in name order
    flags = oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL
    results, raws = [], []
    f = 0
    n_ragged = 0
    for w in range(len(dist)):
        cols = [[si for si, ww in ps[1] if ww == w] for ps in per_sample]
        if not any(cols):
            continue
        if len({len(c) for c in cols}) == 1:
            status = np.stack([ps[2][c] for ps, c in zip(per_sample, cols)])
            post = np.stack([ps[3][c] for ps, c in zip(per_sample, cols)])
            meth = np.stack([ps[4][c] for ps, c in zip(per_sample, cols)])
            D, _, _ = oracle.dmatrix(status, post, 0.99)
            p0 = oracle.p0uu(post, meth, 0.99)[0]
        else:  # DMatrix::from: pairs of unequal length keep D = 0; every sample's level comes from its own sites
            n_ragged += 1
            S_ = len(cols)
            D = np.zeros(S_ * (S_ - 1) // 2)
            p = 0
            for i in range(S_):
                for j in range(i + 1, S_):
                    if len(cols[i]) == len(cols[j]):
                        D[p] = oracle.dmatrix(np.stack([per_sample[i][2][cols[i]], per_sample[j][2][cols[j]]]),
                                              np.stack([per_sample[i][3][cols[i]], per_sample[j][3][cols[j]]]), 0.99)[0][0]
                    p += 1
            own = [oracle.p0uu(ps[3][c][None, :], ps[4][c][None, :], 0.99)[0] if c else float("nan") for ps, c in zip(per_sample, cols)]
            acc = 0.0
            for v in own:  # src/pedigree.rs:179-183: sum of (1 - level) in sample order, / S
                acc += v
            p0 = acc / S_
        if np.isnan(D).any() or np.isnan(p0):
            continue
        ped = ped6.copy()
        ped[:, 3] = D
        pb = oracle.Problem(ped, p0, p0, 1.0)
        sx = ab.gen_start_simplices(seed, w, n_it, float(D.max()))
        rc, best, _, pred, resid = oracle.ab_neutral(pb, sx, flags=flags, n_threads=8)
        assert rc == 0
        idx = ab.gen_resample_idx(seed, w, n_it, len(ped))
        vary = ab.gen_vary_vertices(seed, w, n_it, best["theta"])  # keyed by the window id, like the starts and the resamples
        rc2, rows, _ = oracle.boot_model(pb, best["theta"], pred, resid, idx, vary, flags=flags, n_threads=8)
        assert rc2 == 0
        region = 0 if w < 20 else 1 if w < 40 else 2
        results.append((region, best["theta"].copy(), oracle.analyze(rows), 1.0 - p0))
        raws.append(rows)
        f += 1
    assert len(results) >= 20 and (n_ragged > 0) == ragged
    want = oracle.metaprofile_results_text("run7", dist[:len(results)], [r_[0] for r_ in results], [r_[1] for r_ in results],
                                           [r_[2] for r_ in results], [r_[3] for r_ in results])
    if os.environ.get("ABFIT_TEST_DUMP"):
        open(os.path.join(os.environ["ABFIT_TEST_DUMP"], "mp_want.txt"), "w").write(want)
        open(os.path.join(os.environ["ABFIT_TEST_DUMP"], "mp_got.txt"), "w").write(open(os.path.join(out, "results.txt")).read())
        open(os.path.join(os.environ["ABFIT_TEST_DUMP"], "mp_stdout.txt"), "w").write(r.stdout)
    assert open(os.path.join(out, "results.txt")).read() == want
    raw = np.load(os.path.join(out, "raw.npy"))
    assert raw.shape == (n_it, 7, len(results)) and np.array_equal(raw, np.stack(raws, axis=2))
    # metaplot.png (src/cli/metaprofile.rs:113): a valid 1280 x 960 picture with both series
    from test_plots import decode_png
    img, _ = decode_png(os.path.join(out, "metaplot.png"))
    assert img.shape == (960, 1280, 3)
    # --devices: windows sharded over several GPUs (all visible ones; on a one-GPU box three contexts on device 0) give
    # the same files byte for byte
    out2 = os.path.join(tmp_path, "out_multi")
    os.makedirs(out2)
    nd = ab.device_count()
    devs = f"0-{nd - 1}" if nd >= 2 else "0,0,0"
    r2 = subprocess.run([exe, "-m", meth_dir, "-g", ann, "-w", "10", "-s", "5", "-c", "200", "-o", out2,
                         "--name", "run7", "--iterations", str(n_it), "--seed", str(seed), "--devices", devs, "alphabeta",
                         "--nodes", os.path.join(GOLDEN, "nodelist.txt"), "--edges", os.path.join(GOLDEN, "edgelist.txt")],
                        capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0, r2.stdout + r2.stderr
    for fn in ("results.txt", "raw.npy", "distributions.txt", "steady_state_methylation.txt"):
        assert open(os.path.join(out2, fn), "rb").read() == open(os.path.join(out, fn), "rb").read(), fn


def test_metaprofile_write_windows_layout(oracle, tmp_path):
    """--write-windows: the reference's per-window tree (Windows::save src/windows.rs:259-285, setup_output_dir
    src/setup.rs:5-74).  No GPU involved: extraction only, plus the nodelist / edgelist copies of the sub-command's
    set-up checked through a run that stops at the missing device."""
    ann = os.path.join(tmp_path, "ann.bed")
    genes = [(1, 300, 700, "-"), (1, 250, 800, "+"), (2, 1200, 1600, "+"), (3, 600, 900, "*")]
    open(ann, "w").write("".join(f"{c}\t{s}\t{e}\tg{i}\tgbM\t{sd}\n" for i, (c, s, e, sd) in enumerate(genes)))
    out = os.path.join(tmp_path, "out")
    os.makedirs(out)
    exe = os.path.join(ROOT, "alphabeta-rs_b200", "metaprofile")
    r = subprocess.run([exe, "-m", os.path.join(GOLDEN, "methylome"), "-g", ann, "-w", "10", "-s", "5", "-c", "200", "-o", out,
                        "--write-windows"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    header = ("seqnames\tstart\tstrand\tcontext\tcounts.methylated\tcounts.total\tposteriorMax\tstatus\trc.meth.lvl\t"
              "context.trinucleotide\n")
    og = [(oracle.chromosome_id(str(c)), s, e, STRAND[sd]) for c, s, e, sd in genes]
    for fn in sorted(os.listdir(os.path.join(GOLDEN, "methylome"))):
        lines, sites = [], []
        for line in open(os.path.join(GOLDEN, "methylome", fn)).read().split("\n")[1:]:
            q = oracle.parse_methylome_line(line)
            if q is not None:
                lines.append(line)
                sites.append((oracle.chromosome_id(str(q["chromosome"])), q["start"], q["end"], STRAND[q["strand"]]))
        dist, assign = oracle.extract_windows(og, sites, window_size=10, window_step=5, cutoff=200)
        assert len(dist) == 60 and sum(dist) > 0
        for w in range(60):
            region, i = ("upstream", w) if w < 20 else ("gene", w - 20) if w < 40 else ("downstream", w - 40)
            want = header + "\n".join(lines[si] for si, ww in assign if ww == w)
            assert open(os.path.join(out, region, str(i * 5), fn)).read() == want, (fn, w)
    assert sorted(os.listdir(os.path.join(out, "gene")), key=int) == [str(5 * i) for i in range(20)]
    # set-up of the sub-command: every directory gets the edgelist and the nodelist, '/'-led tab-separated lines re-pathed
    nl = os.path.join(tmp_path, "nodes.tsv")
    open(nl, "w").write("filename\tnode\tgen\tmeth\n/abs/dir/G0.txt\t0_0\t0\tY\n./rel/G1_2.txt\t1_2\t1\tY\nx\t1_0\t1\tN\n")
    el = os.path.join(tmp_path, "edges.tsv")
    open(el, "w").write("from\tto\n0_0\t1_0\n1_0\t1_2\n")
    out2 = os.path.join(tmp_path, "out2")
    os.makedirs(out2)
    subprocess.run([exe, "-m", os.path.join(GOLDEN, "methylome"), "-g", ann, "-w", "10", "-s", "50", "-o", out2, "--write-windows",
                    "--device", "99", "alphabeta", "--nodes", nl, "--edges", el], capture_output=True, text=True, timeout=300)
    d = os.path.join(out2, "downstream", "50")
    assert open(os.path.join(d, "edgelist.txt")).read() == open(el).read()
    assert open(os.path.join(d, "nodelist.txt")).read() == (
        f"filename\tnode\tgen\tmeth\n{d}/G0.txt\t0_0\t0\tY\n./rel/G1_2.txt\t1_2\t1\tY\nx\t1_0\t1\tN\n\n")
