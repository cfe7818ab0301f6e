"""Kernel time of one SV evaluation (N = 2^16, T = 1000) over a grid of parameters: which ones are slow?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
from run_qmh import build
from qnmh_sysid2018_b200.engine import Engine
m = build("sv", 1000)
obs = np.asarray(m.obs, dtype=np.float64).reshape(-1)
eng = Engine("sv", "float64", 0)
eng.configure(obs, np.array([0.5, 0.95, 0.25]), 1 << 16, 10, True, True, 0.0, "systematic", seed=1)
rows = []
for mu in (-0.2, 0.1, 0.3, 0.5, 0.8, 1.1):
    for phi in (0.90, 0.93, 0.96):
        for sv in (0.2, 0.3, 0.4):
            eng.update_params(np.array([mu, phi, sv]), seed=1, eval_ids=[3])
            eng.execute(); eng.synchronize()
            rows.append((eng.last_kernel_ms(), mu, phi, sv, eng.fetch_log_like()[0]))
rows.sort(reverse=True)
for r in rows[:8] + rows[-3:]:
    print("kernel %.2f ms  mu %.2f phi %.2f sigma_v %.2f  ll %.2f" % r)
# per-step statistics of the slowest one
ms, mu, phi, sv, _ = rows[0]
eng.update_params(np.array([mu, phi, sv]), seed=1, eval_ids=[3])
eng.execute(); eng.synchronize()
st = eng.fetch_steps()
print({k: np.asarray(v).shape for k, v in st.items()})
prof = eng.debug_profile()
if prof.any():
    per = prof / 1000.0 / 1.965e3
    print("per-phase us/step, mean over CTAs:", np.round(per.mean(axis=0), 2))
    print("per-phase us/step, max over CTAs: ", np.round(per.max(axis=0), 2))
