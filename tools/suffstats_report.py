"""EXPERIMENT report: the multi-start fit with the objective evaluated from per-triple
sufficient statistics (ABFIT_EXPERIMENT_SUFFSTATS=1: O(distinct triples) per evaluation instead of O(pairs)) against
the exact kernels, which are bit-identical to the oracle (tests/test_gpu_parity.py).  Never the default: the sum is
regrouped, so the Nelder-Mead trajectories differ.  For every window of the C4 synthetic metaprofile (bench.py's
generator, the same seeded start simplices on both sides) the best-of-starts result is compared at north_star's
tolerances: RSS 1e-9 relative (the exact objective evaluated at the experiment's best theta), alpha / beta 1e-6.
  python tools/suffstats_report.py [windows=2000] [starts=1000] [replicates=100]  -> one JSON line
(ABFIT_EXPERIMENT_SUFFSTATS=1: multi-start kernel; =2: the bootstrap refits too, compared replicate by replicate from the
same exact best models)
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from __graft_entry__ import _load_product
ab = _load_product()

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
NS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
shape = bench.load_shape()
peds, p0 = bench.synth_windows(W, 0, shape)
probs = [ab.Problem(peds[i], float(p0[i]), float(p0[i]), 1.0) for i in range(W)]
sx = np.stack([ab.gen_start_simplices(bench.SEED, i, NS, float(peds[i][:, 3].max())) for i in range(W)])
ctx = ab.Context(0)

def run(suff):
    if suff:
        os.environ["ABFIT_EXPERIMENT_SUFFSTATS"] = "1"
    else:
        os.environ.pop("ABFIT_EXPERIMENT_SUFFSTATS", None)
    os.environ["ABFIT_JIT"] = "1"
    b = ctx.batch(probs)
    b.upload_starts(sx)
    assert b.uses_specialised_kernels()
    b.run_fit()  # warm-up (module compile)
    ms = []
    for _ in range(3):
        b.run_fit()
        ms.append(b.timing()["fit_ms"])
    res = b.download_fit(want_all=True)
    ev = b.timing()["evals_fit"]
    b.close()
    return res, float(np.median(ms)), ev

exact, ms_exact, ev_exact = run(False)
suff, ms_suff, ev_suff = run(True)
os.environ.pop("ABFIT_EXPERIMENT_SUFFSTATS", None)
be, bs = exact.best, suff.best
# the exact objective at the experiment's best thetas (interpreter kernel)
_, lse_at_suff = ctx.cost_batch(probs, bs["theta"], np.arange(W, dtype=np.int32))
rel = lambda a, b: np.abs(a - b) / np.maximum(np.abs(b), 1e-300)
d_rss = rel(lse_at_suff, be["lse"])
d_a, d_b = rel(bs["theta"][:, 0], be["theta"][:, 0]), rel(bs["theta"][:, 1], be["theta"][:, 1])
ok = (d_rss <= 1e-9) & (d_a <= 1e-6) & (d_b <= 1e-6)
better = lse_at_suff < be["lse"]
# per start: how far apart do the two objectives drive the same start?
ea, sa = exact.all, suff.all
same_iters = float(np.mean(ea["iters"] == sa["iters"]))
out = {
    "experiment": "sufficient-statistics multi-start objective (ABFIT_EXPERIMENT_SUFFSTATS=1), never the default",
    "windows": W, "starts": NS,
    "fit_ms_exact": ms_exact, "fit_ms_suffstats": ms_suff, "speedup_multi_start_kernel": ms_exact / ms_suff,
    "Mfits_per_s_exact": W * NS / ms_exact / 1e3, "Mfits_per_s_suffstats": W * NS / ms_suff / 1e3,
    "evals_exact": int(ev_exact), "evals_suffstats": int(ev_suff),
    "windows_within_tolerance": int(ok.sum()), "fraction_within_tolerance": float(ok.mean()),
    "windows_rss_within_1e-9": int((d_rss <= 1e-9).sum()), "windows_alpha_within_1e-6": int((d_a <= 1e-6).sum()),
    "windows_beta_within_1e-6": int((d_b <= 1e-6).sum()),
    "windows_where_experiment_found_lower_exact_rss": int(better.sum()),
    "same_winning_start": float(np.mean(be["start_id"] == bs["start_id"])),
    "starts_with_identical_iteration_count": same_iters,
    "rss_rel_diff_quantiles_50_90_99_max": [float(np.quantile(d_rss, q)) for q in (0.5, 0.9, 0.99, 1.0)],
    "alpha_rel_diff_quantiles_50_90_99_max": [float(np.quantile(d_a, q)) for q in (0.5, 0.9, 0.99, 1.0)],
    "beta_rel_diff_quantiles_50_90_99_max": [float(np.quantile(d_b, q)) for q in (0.5, 0.9, 0.99, 1.0)],
    "verdict": "stays an experiment" if not ok.all() else "all windows within tolerance on this data set",
}
# ---- mode 2: the bootstrap refits too (per-replicate statistics), from the EXACT best models on both sides -------------
NB = int(sys.argv[3]) if len(sys.argv) > 3 else 100
N = len(shape)
idx = np.concatenate([ab.gen_resample_idx(bench.SEED, i, NB, N).ravel() for i in range(W)])
vary = np.stack([ab.gen_vary_vertices(bench.SEED, i, NB, exact.best[i]["theta"]) for i in range(W)])

def run_boot(mode):
    if mode:
        os.environ["ABFIT_EXPERIMENT_SUFFSTATS"] = "2"
    else:
        os.environ.pop("ABFIT_EXPERIMENT_SUFFSTATS", None)
    os.environ["ABFIT_JIT"] = "1"
    b = ctx.batch(probs)
    b.upload_boot(idx, vary, best=exact.best, pred=exact.pred, resid=exact.resid)
    b.run_boot()
    ms = []
    for _ in range(3):
        b.run_boot()
        ms.append(b.timing()["boot_ms"])
    rows, _ = b.download_boot()
    b.close()
    return rows, float(np.median(ms))

rows_e, bms_e = run_boot(False)
rows_s, bms_s = run_boot(True)
os.environ.pop("ABFIT_EXPERIMENT_SUFFSTATS", None)
da, db = rel(rows_s[:, :, 0], rows_e[:, :, 0]), rel(rows_s[:, :, 1], rows_e[:, :, 1])
an_e = np.stack([ab.analyze(rows_e[i]) for i in range(W)])
an_s = np.stack([ab.analyze(rows_s[i]) for i in range(W)])
ci = rel(an_s[:, 16:20], an_e[:, 16:20])  # q0.025 / q0.975 of alpha and beta
out["bootstrap"] = {
    "replicates": NB, "boot_ms_exact": bms_e, "boot_ms_suffstats": bms_s, "speedup_bootstrap_kernel": bms_e / bms_s,
    "replicates_alpha_beta_within_1e-6": float(np.mean((da <= 1e-6) & (db <= 1e-6))),
    "replicates_bit_identical": float(np.mean((rows_s[:, :, :4] == rows_e[:, :, :4]).all(axis=2))),
    "alpha_rel_diff_quantiles_50_99_max": [float(np.quantile(da, q)) for q in (0.5, 0.99, 1.0)],
    "ci_bounds_rel_diff_quantiles_50_99_max": [float(np.quantile(ci, q)) for q in (0.5, 0.99, 1.0)],
    "windows_with_all_ci_bounds_within_1e-6": int((ci <= 1e-6).all(axis=1).sum()),
}
out["step_ms_exact"] = ms_exact + bms_e
out["step_ms_experiment"] = ms_suff + bms_s
print(json.dumps(out))
