"""bootstrap kernel alone on the C4 shape: ms per launch for the current ABFIT_DEV_* knobs (development aid)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from __graft_entry__ import _load_product
ab = _load_product()
W, NS, NB = int(sys.argv[1]) if len(sys.argv) > 1 else 4000, 64, 100
shape = bench.load_shape(); N = len(shape)
peds, p0 = bench.synth_windows(W, 0, shape)
probs = [ab.Problem(peds[i], float(p0[i]), float(p0[i]), 1.0) for i in range(W)]
sx = np.stack([ab.gen_start_simplices(bench.SEED, i, NS, float(peds[i][:, 3].max())) for i in range(W)])
idx = np.concatenate([ab.gen_resample_idx(bench.SEED, i, NB, N).ravel() for i in range(W)])
os.environ["ABFIT_JIT"] = "1"
ctx = ab.Context(0)
b = ctx.batch(probs)
b.upload_starts(sx); b.run_fit()
res = b.download_fit()
vary = np.stack([ab.gen_vary_vertices(bench.SEED, i, NB, res.best[i]["theta"]) for i in range(W)])
b.upload_boot(idx, vary)
ms = []
for _ in range(4):
    b.run_boot(); ms.append(b.timing()["boot_ms"])
rows, _ = b.download_boot()
print("knobs", {k: v for k, v in os.environ.items() if k.startswith("ABFIT_DEV")}, "boot ms", ["%.2f" % m for m in ms], "checksum", float(np.nansum(rows)))
