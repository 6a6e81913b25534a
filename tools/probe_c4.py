"""Quick C4-shaped throughput probe (development aid; bench.py is the contract)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import _load_product
ab = _load_product()

def c4_windows(W, seed=0xAB0B200):
    shape = np.loadtxt(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "pedigree.txt"), skiprows=1)
    rng = np.random.default_rng(seed)
    probs = []
    for w in range(W):
        ped = shape.copy()
        # cheap stand-in for c + dt + noise: scale the real D column
        ped[:, 3] = np.maximum(shape[:, 3] * rng.uniform(0.5, 1.5) + rng.normal(0, 5e-4, len(shape)), 0)
        u = rng.uniform(0.6, 0.95)
        probs.append(ab.Problem(ped, u, u, 1.0))
    return probs

if __name__ == "__main__":
    W = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    NS = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    NB = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    ctx = ab.Context(0)
    print(ctx.info(), "fp64 peak TFLOP/s", ctx.measure_fp64_peak())
    probs = c4_windows(W)
    t = time.time()
    sx = np.stack([ab.gen_start_simplices(1, w, NS, float(p.pedigree[:, 3].max())) for w, p in enumerate(probs)])
    print("gen starts %.1fs" % (time.time() - t))
    b = ctx.batch(probs)
    fl = b.flops_per_eval(0)
    print("flops/eval", fl)
    t = time.time()
    b.upload_starts(sx)
    print("upload_starts %.2f s; specialised kernels: %s" % (time.time() - t, b.uses_specialised_kernels()))
    for rep in range(3):
        b.run_fit()
        tm = b.timing()
        fits = W * NS
        print("fit: %.1f ms  %.3f Mfits/s  evals/fit %.0f  %.2f TFLOP/s ; select %.2f ms" % (
            tm["fit_ms"], fits / tm["fit_ms"] / 1e3, tm["evals_fit"] / fits, tm["evals_fit"] * fl["flops"] / tm["fit_ms"] / 1e9, tm["select_ms"]))
    res = b.download_fit(want_all=True)
    st = res.all["status"]
    print("status counts", {int(k): int((st == k).sum()) for k in np.unique(st)}, "iters median", np.median(res.all["iters"]))
    idx = np.concatenate([ab.gen_resample_idx(1, w, NB, p.n_pairs).ravel() for w, p in enumerate(probs)])
    vary = np.stack([ab.gen_vary_vertices(1, w, NB, res.best[w]["theta"]) for w in range(W)])
    b.upload_boot(idx, vary)
    for rep in range(3):
        b.run_boot()
        tm = b.timing()
        fits = W * NB
        print("boot: %.1f ms  %.3f Mfits/s  evals/fit %.0f  %.2f TFLOP/s" % (
            tm["boot_ms"], fits / tm["boot_ms"] / 1e3, tm["evals_boot"] / fits, tm["evals_boot"] * fl["flops"] / tm["boot_ms"] / 1e9))
    rows, bf = b.download_boot(want_fits=True)
    st = bf["status"]
    print("boot status counts", {int(k): int((st == k).sum()) for k in np.unique(st)}, "iters median", np.median(bf["iters"]))
