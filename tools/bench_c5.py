"""C5 (BASELINE configs[4]) on one GPU: observed divergence of one GPU's site range (200 samples x 625 000 sites)
and the ABneutral fit of the 19 900-pair pedigree (1000 starts + 100 bootstrap replicates). Development aid."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import _load_product
ab = _load_product()
from oracle import abref_py as o
from test_gpu_parity import c5_pedigree
rng = np.random.default_rng(17)
ped = c5_pedigree(rng)
a, b, w, c, p0 = 2e-4, 1e-3, 0.04, 0.002, 0.75
dt, _ = o.divergence(o.Problem(ped, p0, p0, 1.0), a, b, w, o.FAST_DIVERGENCE)
ped[:, 3] = np.maximum(c + dt + rng.normal(0, 5e-4, len(ped)), 0.0)
prob = ab.Problem(ped, p0, p0, 1.0)
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
NB = 100
sx = ab.gen_start_simplices(1, 0, NS, float(ped[:, 3].max()))
idx = ab.gen_resample_idx(1, 0, NB, len(ped))
ctx = ab.Context(0)
bt = ctx.batch([prob])
fl = bt.flops_per_eval(0)
print("C5 pedigree:", ped.shape, fl)
bt.upload_starts(sx[None])
for rep in range(2 if NS <= 4000 else 1):
    bt.run_fit(); tm = bt.timing()
    print("fit %d starts: %.1f ms, evals/fit %.0f, %.2f TFLOP/s" % (NS, tm["fit_ms"], tm["evals_fit"] / NS, tm["evals_fit"] * fl["flops"] / tm["fit_ms"] / 1e9))
res = bt.download_fit(want_all=True)
ev = np.sort(res.all[0]["evals"])
print("evals per fit: median %d, p90 %d, p99 %d, max %d -> %.1f us per evaluation on the longest fit's warp" %
      (ev[len(ev) // 2], ev[int(0.9 * len(ev))], ev[int(0.99 * len(ev))], ev[-1], tm["fit_ms"] * 1e3 / ev[-1]))
vary = ab.gen_vary_vertices(1, 0, NB, res.best[0]["theta"])
bt.upload_boot(idx, vary[None])
bt.run_boot(); tm = bt.timing()
print("boot 100 replicates: %.1f ms, evals/fit %.0f" % (tm["boot_ms"], tm["evals_boot"] / NB))
print("best", res.best[0]["theta"], "lse", res.best[0]["lse"])
ctx.close()
