"""Observed pairwise divergence (DMatrix::from + p0uu) on the C5 shape — HBM roofline of the packing pass.
  python tools/bench_divergence.py [S] [L]        default 200 samples x 5 000 000 sites / 8 GPUs = 625 000 sites
Inputs are generated on the device (torch), the kernels are timed by the library's own CUDA events; a
slice of the result is checked against the oracle."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_product
ab = _load_product()
from oracle import abref_py as o

S = int(sys.argv[1]) if len(sys.argv) > 1 else 200
L = int(sys.argv[2]) if len(sys.argv) > 2 else 625_000
NWIN = int(sys.argv[3]) if len(sys.argv) > 3 else 1   # > 1: the site axis cut into NWIN equal windows (metaprofile shape)
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(5)
status = (torch.rand((S, L), device=dev, generator=g) * 3).to(torch.uint8).clamp_(max=2)
post = torch.where(torch.rand((S, L), device=dev, generator=g) < 0.9, torch.full((), 0.9999, device=dev, dtype=torch.float64),
                   torch.rand((S, L), device=dev, generator=g, dtype=torch.float64) * 0.49 + 0.5)
meth = torch.rand((S, L), device=dev, generator=g, dtype=torch.float64)
torch.cuda.synchronize()
ctx = ab.Context(0)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6532.9)
alg_bytes = 17.0 * S * L + 24.0 * (S * (S - 1) // 2)
segs = None if NWIN == 1 else [int(round(i * L / NWIN)) for i in range(NWIN + 1)]
for rep in range(4):
    t = time.perf_counter()
    out = ctx.dmatrix_device(status.data_ptr(), post.data_ptr(), meth.data_ptr(), S, L, seg_offsets=segs)
    wall = time.perf_counter() - t
    pk, pr = out["kernel_ms"]
    print(f"S={S} L={L}: pack {pk:.3f} ms = {17.0 * S * L / pk / 1e6:.0f} GB/s ({17.0 * S * L / pk / 1e6 / hbm * 100:.1f}% of {hbm:.0f}), "
          f"pairs+finalise {pr:.3f} ms, whole call {wall * 1e3:.1f} ms, algorithmic {alg_bytes / (pk + pr) / 1e6:.0f} GB/s over both passes, "
          f"launches {out['launches']}")
if NWIN > 1:
    w = NWIN // 2
    a, b = segs[w], segs[w + 1]
    D, diff, cnt = o.dmatrix(status[:, a:b].cpu().numpy(), post[:, a:b].cpu().numpy(), 0.99)
    assert np.array_equal(out["diff"][w], diff) and np.array_equal(out["cnt"][w], cnt)
    print(f"{NWIN} windows of ~{L // NWIN} sites: window {w} bit-exact against the oracle")
    ctx.close()
    sys.exit(0)
# parity sample: first 6 samples, first 20 000 sites through the oracle
s6, l6 = min(S, 6), min(L, 20000)
sub = ctx.dmatrix(status[:s6, :l6].cpu().numpy(), post[:s6, :l6].cpu().numpy(), meth[:s6, :l6].cpu().numpy(), 0.99)
D, diff, cnt = o.dmatrix(status[:s6, :l6].cpu().numpy(), post[:s6, :l6].cpu().numpy(), 0.99)
assert np.array_equal(sub["diff"][0], diff) and np.array_equal(sub["cnt"][0], cnt) and np.array_equal(sub["D"][0], D)
# the full result is consistent with it on the integer side: sums over site ranges add up
half = ctx.dmatrix_device(status.data_ptr(), post.data_ptr(), meth.data_ptr(), S, L, seg_offsets=[0, L // 2 // 64 * 64, L])
assert np.array_equal(half["diff"].sum(axis=0), out["diff"][0]) and np.array_equal(half["cnt"].sum(axis=0), out["cnt"][0])
print("parity ok (oracle slice bit-exact; site-range partial sums add up exactly)")
ctx.close()
