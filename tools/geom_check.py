"""Geometry (in)dependence of the results: the same filter on different numbers of CTAs."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qnmh_sysid2018_b200.engine import Engine  # noqa: E402

T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for lg in [int(a) for a in sys.argv[1].split(",")]:
    N = 1 << lg
    obs = bench.synthetic_obs("lgss", T)
    res = {}
    for G in (148, 100, 64):
        eng = Engine("lgss", "float64", 0)
        eng.set_option("ctas_per_filter", G)
        eng.configure(obs, np.tile(bench.THETA_LGSS, (1, 1)), N, 10, True, True, 0.0, "systematic", seed=1)
        eng.execute()
        eng.synchronize()
        st = eng.fetch_steps()
        out = eng.fetch()
        res[G] = (st["ll_terms"][0].copy(), out["filt"][0].copy(), eng.info("tile_cap"), eng.info("ctas_per_filter"))
        eng.close()
    a = res[148]
    for G in (100, 64):
        b = res[G]
        d = np.abs(a[0] - b[0])
        first = int(np.argmax(d > 1e-13)) if (d > 1e-13).any() else -1
        print("N=2^%d G=148(%d, cap %d) vs G=%d(%d, cap %d): max |d ll_t| %.3e first step %d; d ll_t[:6] %s; max |d filt| %.3e"
              % (lg, a[3], a[2], G, b[3], b[2], d.max(), first, np.array2string(d[:6], precision=2), np.abs(a[1] - b[1]).max()))
