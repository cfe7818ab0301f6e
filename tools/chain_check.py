"""LGSS qMH / mh0 chains on the B200 estimator at several particle counts (diagnostic for tests/test_gpu_chains.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "examples"))
import test_gpu_chains as tc  # noqa: E402
from qnmh_sysid2018_b200.lockstep import LockstepEstimatorPool  # noqa: E402

fx = tc._fixture()
obs = fx["lgss_pf_obs"]
for alg in sys.argv[2].split(","):
    for N in [int(v) for v in sys.argv[1].split(",")]:
        K, iters, burn = 4, int(sys.argv[3]), 500
        pool = LockstepEstimatorPool(K, dict(tc.PF, no_particles=N))
        outs = pool.run([tc._chain(lambda: tc._lgss(obs), iters, burn, 7 + k, alg=alg) for k in range(K)])
        draws = np.concatenate([o["params"][burn:] for o in outs])
        print(alg, "N=%d" % N, "pooled mean", np.round(draws.mean(axis=0), 4), "std", np.round(draws.std(axis=0), 4),
              "accept", np.round([o["accept_rate"] for o in outs], 2), "per chain",
              [list(np.round(o["posterior_mean"], 3)) for o in outs], "%.2f ms/launch" % (pool.device_ms / pool.batches), flush=True)
print("ref kalman", fx["lgss_kalman_posterior_mean"], "ref pf", fx["lgss_pf_posterior_mean"])
