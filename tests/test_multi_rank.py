"""N > 1 host logic on CPU: two gloo ranks shard windows / site ranges and gather on rank 0.
The numerical work of each shard is done by the ORACLE here (this container has no GPU; the product
refuses to compute without one) — what is under test is the product's sharding, the seeded input
generators being independent of the sharding, the gather in window order and the exact combination of
site-range partial sums."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from conftest import load_product
    from oracle import abref_py as o

    ab = load_product()
    import importlib.util

    spec = importlib.util.spec_from_file_location("abfit_multi", os.path.join(ROOT, "alphabeta-rs_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        # ---- windows sharded over ranks -----------------------------------------------------
        W, n_starts, seed = 5, 6, 0xAB0B200
        ped = np.loadtxt(os.path.join(GOLDEN, "pedigree_generated.txt"), skiprows=1)
        rng = np.random.default_rng(3)
        peds = [np.column_stack([ped[:, :3], np.abs(ped[:, 3] + rng.normal(0, 0.01, len(ped)))]) for _ in range(W)]
        first, count = multi.window_shard(W, rank, world)
        local = np.zeros(count, dtype=ab.FIT_DTYPE)
        for j in range(count):
            w = first + j
            sx = ab.gen_start_simplices(seed, w, n_starts, float(peds[w][:, 3].max()))  # keyed by the GLOBAL window id
            rc, best, _, _, _ = o.ab_neutral(o.Problem(peds[w], 0.7, 0.7, 1.0), sx,
                                             flags=o.FAST_DIVERGENCE | o.EARLY_EXIT_ON_STALL)
            assert rc == 0
            for f in ("theta", "cost", "lse", "iters", "evals", "status", "start_id"):
                local[j][f] = best[f]
        allw = multi.gather_to_root(local)
        # ---- site ranges sharded over ranks --------------------------------------------------
        S, L = 4, 1000
        rs = np.random.default_rng(5)
        status = rs.integers(0, 3, (S, L)).astype(np.uint8)
        post = np.where(rs.random((S, L)) < 0.8, 0.995, 0.7)
        meth = rs.random((S, L))
        f0, n = multi.site_shard(L, rank, world)
        _, diff, cnt = o.dmatrix(status[:, f0:f0 + n], post[:, f0:f0 + n], 0.99)
        valid = post[:, f0:f0 + n] >= 0.99
        methsum = np.array([meth[s, f0:f0 + n][valid[s]].sum() for s in range(S)])
        nvalid = valid.sum(axis=1).astype(np.int64)
        parts = [multi.gather_to_root(np.asarray(x)[None]) for x in (diff, cnt, methsum, nvalid)]
        # ---- ONE problem, its starts sharded over ranks (the C5 fit) -----------------------------
        n_big = 9
        sx_all = ab.gen_start_simplices(seed, 77, n_big, float(peds[0][:, 3].max()))
        s0, sc = multi.start_shard(n_big, rank, world)
        rc, best, _, pred, resid = o.ab_neutral(o.Problem(peds[0], 0.7, 0.7, 1.0), sx_all[s0:s0 + sc],
                                                flags=o.FAST_DIVERGENCE | o.EARLY_EXIT_ON_STALL)
        assert rc == 0
        mine = np.zeros(1, dtype=ab.FIT_DTYPE)
        for f in ("theta", "cost", "lse", "iters", "evals", "status", "start_id"):
            mine[0][f] = best[f]
        cands = multi.gather_to_root(mine)
        firsts = multi.gather_to_root(np.array([s0], dtype=np.int64))
        preds = multi.gather_to_root(np.asarray(pred)[None])
        if rank == 0:
            win, rec = multi.best_of_shards(cands, firsts)
            np.savez(os.path.join(out_dir, "bigfit.npz"), rec=rec, pred=preds[win], win=win)
            np.save(os.path.join(out_dir, "fits.npy"), allw)
            D, p0uu, d, c = multi.combine_site_shards(*parts)
            np.savez(os.path.join(out_dir, "div.npz"), D=D, p0uu=p0uu, d=d, c=c)
    finally:
        dist.destroy_process_group()


def test_window_and_site_shards_cover_everything_once():
    import importlib.util

    spec = importlib.util.spec_from_file_location("abfit_multi", os.path.join(ROOT, "alphabeta-rs_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    for n in (0, 1, 7, 8, 10000):
        for world in (1, 2, 3, 8):
            spans = [multi.window_shard(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    for L in (0, 1, 63, 64, 65, 5_000_000):
        for world in (1, 2, 8):
            spans = [multi.site_shard(L, r, world) for r in range(world)]
            assert sum(c for _, c in spans) == L and all(f % 64 == 0 or c == 0 for f, c in spans)


def test_best_of_shards_rule(ab):
    import importlib.util

    spec = importlib.util.spec_from_file_location("abfit_multi", os.path.join(ROOT, "alphabeta-rs_b200", "multi.py"))
    multi = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(multi)
    c = np.zeros(3, dtype=ab.FIT_DTYPE)
    c["lse"] = [2.0, 1.0, 1.0]
    c["start_id"] = [0, 5, 1]
    win, rec = multi.best_of_shards(c, [0, 10, 3])  # global ids 0, 15, 4: the tie goes to id 4
    assert win == 2 and int(rec["start_id"]) == 4
    assert [multi.start_shard(1000, r, 8) for r in (0, 7)] == [(0, 125), (875, 125)]
    c["lse"][0] = np.nan
    with pytest.raises(FloatingPointError):
        multi.best_of_shards(c, [0, 10, 3])


@pytest.mark.timeout(300)
def test_two_gloo_ranks_equal_one_process(tmp_path, oracle, ab):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(tmp_path, "fits.npy"))
    # single process, all windows
    W, n_starts, seed = 5, 6, 0xAB0B200
    ped = np.loadtxt(os.path.join(GOLDEN, "pedigree_generated.txt"), skiprows=1)
    rng = np.random.default_rng(3)
    peds = [np.column_stack([ped[:, :3], np.abs(ped[:, 3] + rng.normal(0, 0.01, len(ped)))]) for _ in range(W)]
    assert len(got) == W
    for w in range(W):
        sx = ab.gen_start_simplices(seed, w, n_starts, float(peds[w][:, 3].max()))
        rc, best, _, _, _ = oracle.ab_neutral(oracle.Problem(peds[w], 0.7, 0.7, 1.0), sx,
                                              flags=oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL)
        assert np.array_equal(got[w]["theta"], best["theta"]) and got[w]["start_id"] == best["start_id"]
    # one problem, starts sharded: same winner (global start id), same prediction as one process
    sx_all = ab.gen_start_simplices(seed, 77, 9, float(peds[0][:, 3].max()))
    rc, best, _, pred, _ = oracle.ab_neutral(oracle.Problem(peds[0], 0.7, 0.7, 1.0), sx_all,
                                             flags=oracle.FAST_DIVERGENCE | oracle.EARLY_EXIT_ON_STALL)
    z = np.load(os.path.join(tmp_path, "bigfit.npz"))
    assert int(z["rec"]["start_id"]) == int(best["start_id"]) and np.array_equal(z["rec"]["theta"], best["theta"])
    assert z["rec"]["lse"] == best["lse"] and np.array_equal(z["pred"], pred)
    S, L = 4, 1000
    rs = np.random.default_rng(5)
    status = rs.integers(0, 3, (S, L)).astype(np.uint8)
    post = np.where(rs.random((S, L)) < 0.8, 0.995, 0.7)
    meth = rs.random((S, L))
    D, diff, cnt = oracle.dmatrix(status, post, 0.99)
    z = np.load(os.path.join(tmp_path, "div.npz"))
    assert np.array_equal(z["d"], diff) and np.array_equal(z["c"], cnt) and np.array_equal(z["D"], D, equal_nan=True)
    want_p0uu = oracle.p0uu(post, meth, 0.99)[0]
    assert abs(float(z["p0uu"]) - want_p0uu) <= 1e-12 * abs(want_p0uu)
