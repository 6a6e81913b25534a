"""Does the bootstrap kernel hide behind the multi-start kernel when both are in flight?  (development aid)
Two contexts (= two streams) on one GPU: the multi-start fit of one half of the windows and the bootstrap of the
other half, first one after the other, then enqueued together."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from probe_c4 import ab, c4_windows

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
NS, NB = 1000, 100
c1, c2 = ab.Context(0), ab.Context(0)
probs = c4_windows(2 * W)
pa, pb = probs[:W], probs[W:]
sxa = np.stack([ab.gen_start_simplices(1, w, NS, float(p.pedigree[:, 3].max())) for w, p in enumerate(pa)])
sxb = np.stack([ab.gen_start_simplices(1, W + w, NS, float(p.pedigree[:, 3].max())) for w, p in enumerate(pb)])
ba, bb = c1.batch(pa), c2.batch(pb)
ba.upload_starts(sxa)
bb.upload_starts(sxb)
bb.run_fit()
res = bb.download_fit()
idx = np.concatenate([ab.gen_resample_idx(1, W + w, NB, p.n_pairs).ravel() for w, p in enumerate(pb)])
vary = np.stack([ab.gen_vary_vertices(1, W + w, NB, res.best[w]["theta"]) for w in range(W)])
bb.upload_boot(idx, vary)
ba.run_fit(); bb.run_boot(); c1.sync(); c2.sync()  # warm
for rep in range(3):
    t0 = time.perf_counter(); ba.run_fit(); c1.sync(); t1 = time.perf_counter(); bb.run_boot(); c2.sync(); t2 = time.perf_counter()
    ba.run_fit(); bb.run_boot(); c1.sync(); c2.sync(); t3 = time.perf_counter()
    bb.run_boot(); ba.run_fit(); c1.sync(); c2.sync(); t4 = time.perf_counter()
    print("fit %.1f ms + boot %.1f ms = %.1f ms one after the other; together %.1f ms (fit first), %.1f ms (boot first)" % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t2 - t0), 1e3 * (t3 - t2), 1e3 * (t4 - t3)))
