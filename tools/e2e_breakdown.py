"""where does the one-shot host-buffer path spend its time? (development aid)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from __graft_entry__ import _load_product
ab = _load_product()
W, NS, NB = int(sys.argv[1]), 1000, 100
shape = bench.load_shape(); N = len(shape)
peds, p0uu = bench.synth_windows(W, 0, shape)
probs = [ab.Problem(peds[i], float(p0uu[i]), float(p0uu[i]), 1.0) for i in range(W)]
import torch
sx = torch.empty((W, NS, 5, 4), dtype=torch.float64, pin_memory=True).numpy()
idx = torch.empty((W, NB, N), dtype=torch.int32, pin_memory=True).numpy()
vary = torch.empty((W, NB, 4, 4), dtype=torch.float64, pin_memory=True).numpy()
for i in range(W):
    sx[i] = ab.gen_start_simplices(1, i, NS, float(peds[i, :, 3].max())); idx[i] = ab.gen_resample_idx(1, i, NB, N)
ctx = ab.Context(0)
T = lambda: time.perf_counter()
for rep in range(2):
    t = [T()]
    packed = ab._pack_problems(probs); t.append(T())
    b = ctx.batch(probs); t.append(T())
    b.upload_starts(sx); b.sync(); t.append(T())
    b.run_fit(); b.sync(); t.append(T())
    res = b.download_fit(); t.append(T())
    for i in range(W): vary[i] = ab.gen_vary_vertices(1, i, NB, res.best[i]["theta"])
    t.append(T())
    b.upload_boot(idx, vary); b.sync(); t.append(T())
    b.run_boot(); b.sync(); t.append(T())
    rows, _ = b.download_boot(); t.append(T())
    b.close(); t.append(T())
    names = "pack create upload_starts run_fit download_fit gen_vary upload_boot run_boot download_boot destroy".split()
    print(" ".join(f"{n}={1e3*(t[i+1]-t[i]):.1f}ms" for i, n in enumerate(names)), "total=%.1fms" % (1e3*(t[-1]-t[0])))
