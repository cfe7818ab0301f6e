"""Step time with degenerate generations (every weight NaN) against a regular run, cluster and persistent kernels."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qnmh_sysid2018_b200.engine import Engine  # noqa: E402

for model, N, T, opts in (("lgss", 2000, 200, {}), ("sv", 1000, 200, {}), ("sv", 16384, 100, {}), ("lgss", 2000, 200, {"cluster": -1})):
    obs = bench.synthetic_obs(model, T)
    theta = bench.THETA_LGSS if model == "lgss" else bench.THETA_SV
    ms = {}
    for name, bad in (("regular", False), ("derailed", True)):
        o = obs.copy()
        if bad:
            o[3::4] = 1e200
        eng = Engine(model, "float64", 0)
        for k, v in opts.items():
            eng.set_option(k, v)
        eng.configure(o, theta, N, 10, True, True, 0.0, "systematic", seed=5)
        for _ in range(3):
            eng.execute()
            eng.synchronize()
        ms[name] = eng.last_kernel_ms()
        ll = eng.fetch_log_like()[0]
        kern = "cluster x%d" % eng.info("cluster") if eng.info("cluster") else "persistent x%d" % eng.info("ctas_per_filter")
        eng.close()
    print("%s N=%d T=%d %s: regular %.3f ms, derailed %.3f ms (ll %s)" % (model, N, T, kern, ms["regular"], ms["derailed"], ll))
