"""debug aid: GPU fit vs host emulation of the same device source (tests/host_emul)"""
import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_product
ab = _load_product()
from oracle import abref_py as o
E = C.CDLL(os.path.join(ROOT, "tests/host_emul/libemul.so"))
GOLD = os.path.join(ROOT, "tests/golden")
ped351 = o.load_pedigree_file(os.path.join(GOLD, "pedigree.txt"))
ped6 = np.loadtxt(os.path.join(GOLD, "pedigree_generated.txt"), skiprows=1)
def emul_fit(ped, u, sx, max_iters=10000, flags=0):
    arr = ab._pack_problems([ab.Problem(ped, u, u, 1.0)])
    n = sx.shape[0]
    out = np.zeros(n, dtype=ab.FIT_DTYPE)
    rc = E.emul_fit(arr, sx.ctypes.data_as(C.c_void_p), n, None, max_iters, C.c_double(ab.DBL_EPSILON), flags, out.ctypes.data_as(C.c_void_p))
    assert rc == 0, rc
    return out
ctx = ab.Context(0)
for name, ped, u in [("ped6", ped6, 0.655), ("ped351", ped351, 0.8)]:
    for n in (1, 2, 32, 33, 100, 300):
        sx = ab.gen_start_simplices(0xAB0B200, 0, n, float(ped[:, 3].max()))
        for mi in (0, 1, 2, 3, 5, 10, 10000):
            g = ctx.fit_batch([ab.Problem(ped, u, u, 1.0)], sx[None], max_iters=mi).all[0]
            e = emul_fit(ped, u, sx, mi)
            bad = [f for f in ("theta", "cost", "lse", "iters", "evals", "status", "start_id") if not np.array_equal(g[f], e[f])]
            msg = ""
            if bad:
                w = np.where((g["cost"] != e["cost"]) | (g["evals"] != e["evals"]))[0]
                msg = f"first bad start {w[:5]} gpu cost {g['cost'][w[:2]]} emul {e['cost'][w[:2]]} evals {g['evals'][w[:2]]} {e['evals'][w[:2]]}"
            print(name, "n", n, "max_iters", mi, "mismatch:", bad, msg)
