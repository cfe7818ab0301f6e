"""Why did the SV chain take 37 ms per iteration when the kernel takes 12?  Time single evaluations."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
from run_qmh import build
from qnmh_sysid2018_b200 import ParticleMethodsB200
m = build("sv", 1000)
for states in (False, True):
    pf = ParticleMethodsB200({"no_particles": 1 << 16, "fixed_lag": 10, "generate_initial_state": True,
                              "estimate_gradient": True, "estimate_hessian": True, "store_states": states})
    for theta in [(0.5, 0.95, 0.25), (0.54, 0.93, 0.32), (0.2, 0.9, 0.4), (0.8, 0.97, 0.2)]:
        m.store_params(theta) if hasattr(m, "store_params") else None
        ts = []
        for _ in range(4):
            t0 = time.perf_counter(); pf.smoother(m); ts.append(time.perf_counter() - t0)
        print("store_states=%s theta=%s: wall ms %s, device ms %.2f, ll %.2f, cluster=%d G=%d" % (
            states, theta, ["%.1f" % (1e3 * t) for t in ts], pf.last_device_ms, pf.results["log_like"],
            pf._engine.info("cluster"), pf._engine.info("ctas_per_filter")))
