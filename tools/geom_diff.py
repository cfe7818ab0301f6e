"""Where do two geometries of the same filter differ?  (full history compared on the host)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qnmh_sysid2018_b200.engine import Engine  # noqa: E402

lg, T, Ga, Gb = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
N = 1 << lg
obs = bench.synthetic_obs("lgss", T)
hist = {}
for G in (Ga, Gb):
    eng = Engine("lgss", "float64", 0)
    eng.set_option("ctas_per_filter", G)
    eng.configure(obs, np.tile(bench.THETA_LGSS, (1, 1)), N, 3, True, True, 0.0, "systematic", seed=1, store_history=True)
    eng.execute()
    eng.synchronize()
    h = eng.fetch_history()
    st = eng.fetch_steps()
    hist[G] = (h["particles"][0], h["log_weights"][0], h["ancestors"][0], st["ll_terms"][0], st["sum_w"][0], st["max_logw"][0],
               eng.info("ctas_per_filter"), eng.info("tile_cap"))
    print("G=%d -> ctas %d tile_cap %d items %d" % (G, eng.info("ctas_per_filter"), eng.info("tile_cap"), eng.info("items")))
    eng.close()
a, b = hist[Ga], hist[Gb]
for t in range(T + 1):
    da = np.nonzero(a[2][t] != b[2][t])[0]
    dx = np.nonzero(a[0][t] != b[0][t])[0]
    print("t=%d: ancestors differ at %d indices %s (a: %s, b: %s); particles differ at %d; d ll_t %.3e; S a %.17g b %.17g; m a %.17g b %.17g"
          % (t, da.size, da[:6], a[2][t][da[:6]], b[2][t][da[:6]], dx.size, a[3][t] - b[3][t], a[4][t], b[4][t], a[5][t], b[5][t]))
    if dx.size:
        j = dx[0]
        print("    first differing particle n=%d: x %r vs %r, lw %r vs %r, anc %d vs %d" % (j, a[0][t][j], b[0][t][j], a[1][t][j], b[1][t][j], a[2][t][j], b[2][t][j]))
