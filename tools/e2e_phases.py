"""host-side phases of the one-shot call abfit_alphabeta_batch (ABFIT_DEV_VERBOSE prints them): development aid"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from __graft_entry__ import _load_product
ab = _load_product()
import torch
W, NS, NB = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 1000, 100
shape = bench.load_shape(); N = len(shape)
from concurrent.futures import ThreadPoolExecutor
pool = ThreadPoolExecutor(16)
peds, probs, sx_t, idx_t = bench.make_inputs(ab, torch, W, 0, NS, NB, shape, pool)
sx, idx = sx_t.numpy(), idx_t.numpy()
ctx = ab.Context(0)
for rep in range(3):
    if rep == 2:
        os.environ["ABFIT_DEV_VERBOSE"] = "1"
    t = time.perf_counter()
    out = ctx.alphabeta_batch(probs, sx, idx, bench.SEED)
    print("call %.1f ms" % (1e3 * (time.perf_counter() - t)), flush=True)
