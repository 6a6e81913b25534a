"""Host-side multi-GPU plumbing of the fit path: one process per GPU, windows (or, for
the observed divergence, site ranges) partitioned across ranks, results gathered on the host.  There
is no data-path collective — the shards never exchange data mid-run; `torch.distributed` only carries
the final gather (gloo on CPU in the tests, NCCL or gloo on the GPU box)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def window_shard(n_windows: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous, balanced block of windows for `rank`: [first, first + count).  All starts and bootstrap
    replicates of a window stay on one device (the bootstrap needs that window's best-of-starts)."""
    if world <= 0 or not 0 <= rank < world or n_windows < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_windows, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def site_shard(n_sites: int, rank: int, world: int, align: int = 64) -> Tuple[int, int]:
    """contiguous site range for `rank`, aligned to the 64-site words of the packed bit-planes"""
    words = (n_sites + align - 1) // align
    first_w, count_w = window_shard(words, rank, world)
    first = min(first_w * align, n_sites)
    end = min((first_w + count_w) * align, n_sites)
    return first, end - first


def gather_to_root(local: np.ndarray, root: int = 0, group=None) -> Optional[np.ndarray]:
    """concatenate per-rank arrays (same dtype and trailing shape, any leading length) in rank order on `root`"""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = np.ascontiguousarray(local)
    raw = torch.from_numpy(local.view(np.uint8).reshape(-1).copy())
    sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([raw.numel()], dtype=torch.int64), group=group)
    sizes = [int(s.item()) for s in sizes]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.uint8)
    buf[: raw.numel()] = raw
    out = [torch.zeros(pad, dtype=torch.uint8) for _ in range(world)] if rank == root else None
    dist.gather(buf, out, dst=root, group=group)
    if rank != root:
        return None
    parts = []
    item = local.dtype.itemsize * int(np.prod(local.shape[1:], dtype=np.int64))
    for r in range(world):
        a = out[r][: sizes[r]].numpy().view(local.dtype)
        parts.append(a.reshape((sizes[r] // item,) + local.shape[1:]) if item else a)
    return np.concatenate(parts, axis=0)


def combine_site_shards(diff: Sequence[np.ndarray], cnt: Sequence[np.ndarray], methsum: Sequence[np.ndarray],
                        nvalid: Sequence[np.ndarray]):
    """Observed divergence from per-shard partial sums (abfit_divergence's diff/cnt/methsum/nvalid outputs of
    each site range, in rank order): the integer sums are exact, so D = diff / (2 cnt) (src/pedigree.rs:257) is
    bit-identical for any number of GPUs; methsum partials are added in rank order (p0uu within 1e-12 rel)."""
    d = np.sum(np.stack([np.asarray(x, dtype=np.uint64) for x in diff]), axis=0, dtype=np.uint64)
    c = np.sum(np.stack([np.asarray(x, dtype=np.uint64) for x in cnt]), axis=0, dtype=np.uint64)
    with np.errstate(divide="ignore", invalid="ignore"):
        D = d.astype(np.float64) / (2.0 * c.astype(np.float64))
    ms = np.zeros_like(np.asarray(methsum[0], dtype=np.float64))
    nv = np.zeros_like(np.asarray(nvalid[0], dtype=np.int64))
    for m, n in zip(methsum, nvalid):
        ms = ms + np.asarray(m, dtype=np.float64)
        nv = nv + np.asarray(n, dtype=np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rc = ms / nv.astype(np.float64)
    S = rc.shape[-1]
    p0uu = np.zeros(rc.shape[:-1])
    for s in range(S):  # src/pedigree.rs:179-183: sequential sum over samples, then / S
        p0uu = p0uu + (1.0 - rc[..., s])
    return D, p0uu / float(S), d, c


def start_shard(n_starts: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block of START ids for `rank` — the sharding of one huge problem (the C5 fit is
    1 window x 1000 starts).  The seeded start generator is keyed by the global start id, so rank r fits
    simplices[first : first + count] of the same array a single process would use."""
    return window_shard(n_starts, rank, world)


def best_of_shards(cands: np.ndarray, first_start: Sequence[int]) -> Tuple[int, np.ndarray]:
    """Final argmin over the per-rank winners of a start-sharded fit (one `abfit_fit` record per rank, in rank
    order; `first_start[r]` = first global start id of rank r).  Same rule as one process: smallest penalty-free
    LSE wins and the lowest start id breaks ties (stable sort, src/ab_neutral.rs:83-101); a NaN LSE is the
    reference's panic (`partial_cmp().unwrap()`, :100).  Returns (winning rank, its record with the GLOBAL start id)."""
    cands = np.asarray(cands)
    if len(cands) == 0 or len(cands) != len(first_start):
        raise ValueError("one candidate and one first_start per rank")
    if np.isnan(cands["lse"]).any() or (cands["status"] < 0).any():
        raise FloatingPointError("a shard returned a failed fit / NaN LSE (the reference panics here)")
    gid = cands["start_id"].astype(np.int64) + np.asarray(first_start, dtype=np.int64)
    win = 0
    for r in range(1, len(cands)):
        if cands["lse"][r] < cands["lse"][win] or (cands["lse"][r] == cands["lse"][win] and gid[r] < gid[win]):
            win = r
    rec = cands[win].copy()
    rec["start_id"] = gid[win]
    return win, rec
