"""Trace of the device time per estimator call along an SV qMH chain (which calls are slow, and where)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "examples"))
from run_qmh import build
from mh_driver import ParticleMH
from qnmh_sysid2018_b200 import ParticleMethodsB200
m = build("sv", 1000)
pf = ParticleMethodsB200({"no_particles": 1 << 16, "fixed_lag": 10, "generate_initial_state": True,
                          "estimate_gradient": True, "estimate_hessian": True})
trace = []
orig = pf.smoother
def wrapped(model):
    orig(model)
    trace.append((pf.last_device_ms, pf._engine.last_kernel_ms(), tuple(np.round(model.get_all_params(), 4)), pf.results["log_like"]))
pf.smoother = wrapped
mh = ParticleMH(m, pf, "qmh", step_size=0.5)
mh.run(int(sys.argv[1]) if len(sys.argv) > 1 else 1500, m.get_free_params(), burn_in=0, verbose=0)
t = np.array([r[0] for r in trace])
print("calls %d, device ms: median %.2f, mean %.2f, max %.2f; calls above 20 ms: %d" % (len(t), np.median(t), t.mean(), t.max(), (t > 20).sum()))
slow = [i for i in range(len(t)) if t[i] > 20]
for i in slow[:12]:
    print("  call %5d: execute %.2f ms, kernel %.2f ms, theta %s, ll %.2f" % (i, trace[i][0], trace[i][1], trace[i][2], trace[i][3]))
for i in range(0, len(t), max(1, len(t) // 15)):
    print("  call %5d: execute %.2f ms kernel %.2f theta %s" % (i, trace[i][0], trace[i][1], trace[i][2]))
