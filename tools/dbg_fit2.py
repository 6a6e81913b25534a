import ctypes as C, os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import _load_product
ab = _load_product()
GOLD = os.path.join(ROOT, "tests/golden")
ped6 = np.loadtxt(os.path.join(GOLD, "pedigree_generated.txt"), skiprows=1)
np.set_printoptions(precision=17, linewidth=200)
ctx = ab.Context(0)
u = 0.655
pr = ab.Problem(ped6, u, u, 1.0)
sx = ab.gen_start_simplices(0xAB0B200, 0, 1, float(ped6[:, 3].max()))
print("simplex\n", sx[0])
c, l = ctx.cost_batch([pr], sx[0])
print("cost of vertices", c, "lse", l)
for mi in (0, 1):
    g = ctx.fit_batch([pr], sx[None], max_iters=mi).all[0]
    print("max_iters", mi, g)
