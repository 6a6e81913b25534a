"""BASELINE configs[4] on N GPUs of one box (development aid; bench.py is the contract):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/c5_multi.py [L=5000000] [n_starts=1000]

200 synthetic methylomes x L CG sites: every rank generates its contiguous site range on its GPU
(multi.site_shard), computes exact integer partial sums of the observed divergence (abfit_divergence_device),
rank 0 adds them (multi.combine_site_shards) into the 19 900-pair pedigree; then the ABneutral fit of that one
pedigree with its starts sharded over the ranks (multi.start_shard, warp-per-fit kernels) and the reference's
best-of-starts rule applied to the per-rank winners on the host (multi.best_of_shards).  No data-path collective:
torch.distributed only carries the gathers, the broadcast of the pedigree and the timing reductions."""
import importlib.util
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from __graft_entry__ import _load_product  # noqa: E402

ab = _load_product()
spec = importlib.util.spec_from_file_location("abfit_multi", os.path.join(ROOT, "alphabeta-rs_b200", "multi.py"))
multi = importlib.util.module_from_spec(spec)
spec.loader.exec_module(multi)


def c5_times(lineages=10, generations=20):
    samples = [(l, g) for l in range(lineages) for g in range(1, generations + 1)]
    rows = []
    for i in range(len(samples)):
        for j in range(i + 1, len(samples)):
            (l1, g1), (l2, g2) = samples[i], samples[j]
            rows.append([min(g1, g2) if l1 == l2 else 0, g1, g2, 0.0])
    return np.array(rows, dtype=np.float64)


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    n_starts = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("gloo")  # host-side gathers only
    lineages, generations = 10, 20
    S = lineages * generations
    first, count = multi.site_shard(L, rank, world)

    # ---- this rank's site range: 10 lineages drifting away from one founder ------------------------------
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)
    founder = (torch.rand(count, device=dev, generator=g) < 0.25).to(torch.uint8) * 2
    status = torch.empty((S, count), dtype=torch.uint8, device=dev)
    for l in range(lineages):
        cur = founder.clone()
        for t in range(generations):
            r = torch.rand(count, device=dev, generator=g)
            cur = torch.where((cur == 0) & (r < 2e-4), torch.full_like(cur, 2),
                              torch.where((cur == 2) & (r < 1e-3), torch.zeros_like(cur), cur))
            status[l * generations + t] = cur
    post = torch.where(torch.rand((S, count), device=dev, generator=g) < 0.9,
                       torch.full((), 0.9999, device=dev, dtype=torch.float64),
                       torch.rand((S, count), device=dev, generator=g, dtype=torch.float64) * 0.49 + 0.5)
    meth = (status.to(torch.float64) * 0.5 + 0.05 * torch.randn((S, count), device=dev, generator=g, dtype=torch.float64)).clamp_(0, 1)
    torch.cuda.synchronize()
    ctx = ab.Context(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- observed divergence: site-sharded, exact integer partial sums ------------------------------------
    ctx.dmatrix_device(status.data_ptr(), post.data_ptr(), meth.data_ptr(), S, count)  # warm-up
    barrier()
    t0 = time.perf_counter()
    out = ctx.dmatrix_device(status.data_ptr(), post.data_ptr(), meth.data_ptr(), S, count)
    barrier()
    t_div = time.perf_counter() - t0
    parts = [multi.gather_to_root(np.asarray(out[k][0])[None]) for k in ("diff", "cnt", "methsum", "nvalid")]
    ped = c5_times(lineages, generations)
    meta = np.zeros(2)
    if rank == 0:
        D, p0uu, d, c = multi.combine_site_shards(*parts)
        ped[:, 3] = D
        meta[:] = [float(p0uu), float(np.nanmax(D))]
    if world > 1:
        tp, tm = torch.from_numpy(ped), torch.from_numpy(meta)
        dist.broadcast(tp, 0)
        dist.broadcast(tm, 0)
    p0uu, max_d = float(meta[0]), float(meta[1])

    # ---- one pedigree, its starts sharded over the ranks ----------------------------------------------------
    sx_all = ab.gen_start_simplices(0xAB0B200, 0, n_starts, max_d)
    s0, sc = multi.start_shard(n_starts, rank, world)
    prob = ab.Problem(ped, p0uu, p0uu, 1.0)
    bt = ctx.batch([prob])
    bt.upload_starts(np.ascontiguousarray(sx_all[s0:s0 + sc])[None])
    bt.run_fit()  # warm-up (module load, scratch allocation)
    barrier()
    t0 = time.perf_counter()
    bt.run_fit()
    res = bt.download_fit()
    barrier()
    t_fit = time.perf_counter() - t0
    tm = bt.timing()
    cands = multi.gather_to_root(res.best[:1].copy())
    firsts = multi.gather_to_root(np.array([s0], dtype=np.int64))
    times = multi.gather_to_root(np.array([[t_div, t_fit, out["kernel_ms"][0], out["kernel_ms"][1], tm["fit_ms"]]]))
    if rank == 0:
        win, rec = multi.best_of_shards(cands, firsts)
        mx = times.max(axis=0)
        print(json.dumps({
            "workload": f"C5: {S} samples x {L} sites, {len(ped)} pairs, {n_starts} starts", "n_gpus": world,
            "divergence_call_ms": 1e3 * mx[0], "pack_kernel_ms": mx[2], "pair_kernel_ms": mx[3],
            "fit_call_ms": 1e3 * mx[1], "fit_kernel_ms": mx[4], "p0uu": p0uu,
            "best": {"alpha": float(rec["theta"][0]), "beta": float(rec["theta"][1]), "lse": float(rec["lse"]),
                     "start_id": int(rec["start_id"]), "rank": int(win)}}))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
