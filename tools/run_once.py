"""Runs the persistent kernel a few times on one configuration (profiling target for ncu)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qnmh_sysid2018_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="lgss")
ap.add_argument("--N", type=int, default=1 << 20)
ap.add_argument("--T", type=int, default=1000)
ap.add_argument("--B", type=int, default=1)
ap.add_argument("--L", type=int, default=10)
ap.add_argument("--dtype", default="float64")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--G", type=int, default=0)
ap.add_argument("--no-smoother", action="store_true")
ap.add_argument("--inject", action="store_true")
ap.add_argument("--opt", action="append", default=[], help="name=value engine option")
a = ap.parse_args()
theta = bench.THETA_LGSS if a.model == "lgss" else bench.THETA_SV
obs = bench.synthetic_obs(a.model, a.T)
eng = Engine(a.model, a.dtype, 0)
if a.G:
    eng.set_option("ctas_per_filter", a.G)
for o in a.opt:
    k, v = o.split("=")
    eng.set_option(k, int(v))
draws = None
if a.inject:
    from oracle import pf_oracle as orc
    draws = orc.random_draws(np.random.default_rng(0), a.N, a.T)
eng.configure(obs, np.tile(theta, (a.B, 1)), a.N, a.L, not a.no_smoother, True, 0.0, "systematic", seed=1, draws=draws)
ms = []
for _ in range(a.reps):
    eng.execute()
    eng.synchronize()
    ms.append(eng.last_kernel_ms())
ll = eng.fetch_log_like()
units = a.B * a.N * a.T
print("N=%d T=%d B=%d %s G=%d groups=%d: kernel ms %s -> %.2f G particle-steps/s, %.2f us/time-step, ll[0]=%.12f"
      % (a.N, a.T, a.B, a.dtype, eng.info("ctas_per_filter"), eng.info("n_groups"),
         ["%.3f" % m for m in ms], units / min(ms) / 1e6, min(ms) * 1e3 / a.T, ll[0]))
prof = eng.debug_profile()
if prof.any():
    cl_names = ["children wait", "A: max/exp/scan", "s store + all-gather wait", "prefix + bounds", "normalise + publish", "B: search", "B: gather + normals", "B: propagate/weight/stores", "B: loop tail", "-", "-", "-", "-", "-", "-", "-"]
    names = cl_names if eng.info("cluster") > 0 else ["A:scan(+rest)", "filt block_sum", "exchange1 wait", "prefix+bounds", "children", "block_max", "exchange2 wait", "xchg_max", "A:load+exp", "X1 arrive", "X2 arrive", "B: search", "-", "B: gather + normals", "B: propagate/weight", "B: stores"]
    per_step = prof / float(a.T) / 1.965e3  # us at 1965 MHz
    for i, nm in enumerate(names):
        print("   %-20s mean %7.2f us  max %7.2f us  min %7.2f us" % (nm, per_step[:, i].mean(), per_step[:, i].max(), per_step[:, i].min()))
    print("   total (mean over CTAs) %.2f us/step" % per_step[:, :13].sum(axis=1).mean())
