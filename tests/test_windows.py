"""Site -> window assignment.

1. The oracle's restatement is pinned on the reference's own enabled tests
   (src/methylation_site.rs:599-979, src/windows.rs:361-395) and on data/output_metaplot (all-zero
   distributions for data/methylome + data/annotation.bed).
2. The product's host implementation (abfit_place_sites, C ABI) must agree with the oracle bit for bit:
   same distribution, same (site, window) list — on the golden files and on random genomes with
   overlapping genes, both strands, unknown strands, absolute / relative windows, gene-length cutoffs.
No GPU involved: this part of the path is integer / f64-compare logic on the host."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def site(loc, strand, chrom=1):
    return (chrom, loc, loc, strand)  # MethylationSite::new(location, strand): start == end == location


def place(oracle, s, gene, counts, **kw):
    return oracle.place_in_windows(s, gene, counts, kw["window_size"], kw["window_step"], kw["cutoff"], kw["absolute"])


# ---- the reference's unit tests, restated --------------------------------------------------------
def test_ref_windows_new_counts(oracle):
    # src/windows.rs:361-395
    assert oracle.window_counts(512, 256, 2048, 4096, True) == (8, 16, 8)
    assert oracle.window_counts(5, 1, 2048, 4096, False) == (100, 100, 100)


def test_ref_place_site_relative_acting_like_absolute(oracle):
    # src/methylation_site.rs:661-713
    kw = dict(window_size=2, window_step=1, cutoff=100, absolute=False)
    counts = oracle.window_counts(2, 1, 100, 100, False)
    within, up, down = (1, 100, 200, 1), (1, 200, 300, 1), (1, 0, 100, 1)
    for i in range(1, 100):
        cg = site(i + 100, 1)
        assert (0, i) in place(oracle, cg, up, counts, **kw)
        assert (1, i) in place(oracle, cg, within, counts, **kw)
        assert (2, i) in place(oracle, cg, down, counts, **kw)


@pytest.mark.parametrize("strand", [1, -1])
def test_ref_place_site_relative(oracle, strand):
    # src/methylation_site.rs:714-768 (sense) and :872-927 (antisense: windows reversed)
    kw = dict(window_size=2, window_step=1, cutoff=1000, absolute=False)
    counts = oracle.window_counts(2, 1, 1000, 1000, False)
    assert counts[0] == 100
    within, up, down = (1, 1000, 2000, strand), (1, 2000, 3000, strand), (1, 0, 1000, strand)
    for i in range(1, 1000):
        cg = site(i + 1000, strand)
        w = i // 10 if strand > 0 else (999 - i) // 10
        # for the antisense case the reference's "upstream"/"downstream" genes swap roles
        regions = {0: up, 1: within, 2: down} if strand > 0 else {0: down, 1: within, 2: up}
        for region, gene in regions.items():
            assert (region, w) in place(oracle, cg, gene, counts, **kw), (i, region)


def test_ref_place_site(oracle):
    # src/methylation_site.rs:769-818
    kw = dict(window_size=2, window_step=1, cutoff=2048, absolute=False)
    counts = oracle.window_counts(2, 1, 2048, 1000, False)
    gene = (1, 100, 200, 1)
    expect = {80: [(0, 98), (0, 99)], 100: [(1, 0)], 123: [(1, 21), (1, 22), (1, 23)], 200: [(1, 99)], 201: [(2, 0)],
              712: [(2, 24)], 1224: [(2, 49)], 2248: [(2, 99)]}
    for loc, want in expect.items():
        got = place(oracle, site(loc, 1), gene, counts, **kw)
        for w in want:
            assert w in got, (loc, w, got)


def test_ref_place_site_absolute_2(oracle):
    # src/methylation_site.rs:819-871
    kw = dict(window_size=2, window_step=1, cutoff=2048, absolute=True)
    counts = oracle.window_counts(2, 1, 2048, 100, True)
    gene = (1, 100, 200, 1)
    expect = {80: [(0, 2026), (0, 2027), (0, 2028)], 100: [(1, 0)], 123: [(1, 21), (1, 22), (1, 23)], 200: [(1, 99)],
              201: [(2, 0), (2, 1)], 220: [(2, 18), (2, 19), (2, 20)]}
    for loc, want in expect.items():
        got = place(oracle, site(loc, 1), gene, counts, **kw)
        for w in want:
            assert w in got, (loc, w, got)


def test_ref_place_site_absolute_invert(oracle):
    # src/methylation_site.rs:928-978
    kw = dict(window_size=2, window_step=1, cutoff=1000, absolute=True)
    counts = oracle.window_counts(2, 1, 1000, 1000, True)
    assert counts[0] == 1000
    within, up, down = (1, 1000, 2000, 1), (1, 2000, 3000, 1), (1, 0, 1000, 1)
    for i in range(1, 1000):
        cg = site(i + 1000, 1)
        assert (0, i) in place(oracle, cg, up, counts, **kw)
        assert (1, i) in place(oracle, cg, within, counts, **kw)
        assert (2, i) in place(oracle, cg, down, counts, **kw)


# ---- golden metaprofile fixture -------------------------------------------------------------------
def load_golden_genes(oracle):
    genes = []
    for line in open(os.path.join(GOLDEN, "annotation.bed")).read().split("\n"):
        g = oracle.parse_annotation_line(line)
        if g is not None:
            genes.append(g)
    return genes


def load_golden_sites(oracle, name):
    sites = []
    lines = open(os.path.join(GOLDEN, "methylome", name)).read().split("\n")
    for line in lines[1:]:  # Windows::extract skips the header row (src/windows.rs:322)
        s = oracle.parse_methylome_line(line)
        if s is not None:
            sites.append((oracle.chromosome_id(str(s["chromosome"])), s["start"], s["end"], {"+": 1, "-": -1, "*": 0}[s["strand"]]))
    return sites


def test_golden_metaprofile_distribution_is_all_zero(ab, oracle):
    """C3 on the shipped data: data/output_metaplot/distributions.txt (60 windows, all 0)"""
    genes = load_golden_genes(oracle)
    assert len(genes) > 5000
    want = {}
    for line in open(os.path.join(GOLDEN, "output_metaplot", "distributions.txt")).read().split("\n"):
        if line:
            f = line.split(";")
            want[f[0]] = [int(x) for x in f[1:] if x != ""]
    for name, w in want.items():
        sites = load_golden_sites(oracle, name)
        assert len(sites) == 500  # src/methylation_site.rs:553-593
        dist, _ = oracle.extract_windows(genes, sites, window_size=5, window_step=0, cutoff=2048)
        assert dist == w and len(w) == 60
        got, asite, awin = ab.place_sites(genes, sites, window_size=5, window_step=0, cutoff=2048)
        assert got.tolist() == w and len(asite) == 0


# ---- product vs oracle -----------------------------------------------------------------------------
def random_case(rng, n_genes, n_sites, span, unknown_frac):
    genes, sites = [], []
    for _ in range(n_genes):
        c = int(rng.integers(1, 4))
        st = int(rng.integers(0, span))
        ln = int(rng.integers(0, span // 20 + 2))
        strand = 0 if rng.random() < unknown_frac else int(rng.choice([1, -1]))
        genes.append((c, st, st + ln, strand))
    locs = np.sort(rng.integers(0, span + span // 10, n_sites))
    for loc in locs:
        c = int(rng.integers(1, 5))  # chromosome 4 has no genes
        strand = 0 if rng.random() < unknown_frac else int(rng.choice([1, -1]))
        sites.append((c, int(loc), int(loc) + int(rng.integers(0, 2)), strand))
    return genes, sites


@pytest.mark.parametrize("absolute", [False, True])
@pytest.mark.parametrize("cutoff_gene_length", [False, True])
def test_product_matches_oracle_on_random_genomes(ab, oracle, absolute, cutoff_gene_length):
    rng = np.random.default_rng(11 + 2 * absolute + cutoff_gene_length)
    for (size, step, cutoff) in ((5, 1, 200), (5, 0, 64), (7, 3, 100), (1, 1, 50), (512, 256, 2048), (3, 5, 0)):
        genes, sites = random_case(rng, 40, 1500, 20000, 0.2)
        mgl = max(g[2] - g[1] for g in genes) if absolute else 100
        kw = dict(window_size=size, window_step=step, cutoff=cutoff, max_gene_length=mgl, absolute=absolute,
                  cutoff_gene_length=cutoff_gene_length)
        want_dist, want_assign = oracle.extract_windows(genes, sites, **kw)
        dist, asite, awin = ab.place_sites(genes, sites, **kw)
        assert tuple(ab.window_counts(**kw)) == oracle.window_counts(size, step, cutoff, mgl, absolute)
        assert dist.tolist() == want_dist, kw
        assert list(zip(asite.tolist(), awin.tolist())) == want_assign, kw
        assert sum(want_dist) > 0 or cutoff == 0 or absolute or not want_dist


def test_overlapping_genes_keep_the_cached_gene(ab, oracle):
    """the previously matched gene is reused while the site is still inside it (src/windows.rs:331-333)"""
    genes = [(1, 100, 1000, 1), (1, 300, 400, 1), (1, 350, 2000, 1)]
    sites = [(1, p, p, 1) for p in range(90, 2100, 7)]
    kw = dict(window_size=5, window_step=1, cutoff=10, max_gene_length=100, absolute=False)
    want_dist, want_assign = oracle.extract_windows(genes, sites, **kw)
    dist, asite, awin = ab.place_sites(genes, sites, **kw)
    assert dist.tolist() == want_dist and list(zip(asite.tolist(), awin.tolist())) == want_assign
    assert sum(want_dist) > 100


def test_degenerate_arguments(ab):
    with pytest.raises(ab.AbfitError):
        ab.window_counts(window_size=0, window_step=0)
    # zero-length gene in relative mode: position = 0/0 = NaN compares false with every bound -> no placement
    dist, asite, _ = ab.place_sites([(1, 50, 50, 1)], [(1, 50, 50, 1)], window_size=5, window_step=1, cutoff=10)
    assert dist.sum() == 0 and len(asite) == 0
