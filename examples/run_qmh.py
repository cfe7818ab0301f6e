"""End-to-end particle-qMH chains on the B200 estimator (configs C1 and C4 of BASELINE.json).

    python examples/run_qmh.py lgss [--iters 2000] [--particles 2000] [--obs 500] [--alg qmh]
    python examples/run_qmh.py sv   [--iters 1000] [--particles 65536] [--obs 1000]

LGSS follows example 2 of the reference (scripts/example2_lgss_particles.py: true parameters
(0.2, 0.5, 1.0, 0.5), inference on (mu, phi, sigma_v), N = 2000, L = 10); the SV case follows example
3's settings on synthetic data (the Bitcoin series needs network access)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from mh_driver import ParticleMH  # noqa: E402
from qnmh_sysid2018_b200 import (LinearGaussianModel, ParticleMethodsB200,  # noqa: E402
                                 StochasticVolatilityModel, StochasticVolatilityModelLeverage)


def build(kind, no_obs, seed=12345):
    if kind == "lgss":
        m = LinearGaussianModel()
        m.params.update(mu=0.2, phi=0.5, sigma_v=1.0, sigma_e=0.5)
        m.initial_state = 0.0
        est = ("mu", "phi", "sigma_v")
        start = (0.2, 0.5, 1.0)
    else:
        m = StochasticVolatilityModelLeverage() if kind == "sv_leverage" else StochasticVolatilityModel()
        m.params.update(mu=0.5, phi=0.95, sigma_v=0.25)
        if kind == "sv_leverage":
            m.params["rho"] = -0.3
        m.initial_state = 0.5
        est = tuple(m.params.keys())
        start = tuple(m.params.values())
    m.generate_data(np.random.RandomState(seed), no_obs=no_obs)
    m.fix_true_params()
    m.create_inference_model(est)
    m.store_params(start)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model", choices=["lgss", "sv", "sv_leverage"])
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--burn-in", type=int, default=200)
    ap.add_argument("--particles", type=int, default=0)
    ap.add_argument("--obs", type=int, default=0)
    ap.add_argument("--alg", default="qmh", choices=["mh0", "mh1", "qmh"])
    ap.add_argument("--dtype", default="float64")
    ap.add_argument("--step-size", type=float, default=0.5)
    a = ap.parse_args()
    N = a.particles or (2000 if a.model == "lgss" else 1 << 16)
    T = a.obs or (500 if a.model == "lgss" else 1000)
    model = build(a.model, T)
    pf = ParticleMethodsB200({"no_particles": N, "fixed_lag": 10, "generate_initial_state": True,
                              "estimate_gradient": True, "estimate_hessian": True, "dtype": a.dtype})
    mh = ParticleMH(model, pf, a.alg, step_size=a.step_size)
    out = mh.run(a.iters, model.get_free_params(), burn_in=a.burn_in, verbose=max(1, a.iters // 5))
    print("model %s  alg %s  N=%d  T=%d  %d iterations: %.1f s (%.2f ms/iteration, %.2f G particle-steps/s)"
          % (a.model, a.alg, N, T, a.iters, out["seconds"], 1e3 * out["time_per_iteration"],
             N * T / out["time_per_iteration"] / 1e9))
    print("state estimator: %d calls, %.2f ms per call on the host clock, %.2f ms per call on the device"
          % (out["estimator_calls"], 1e3 * out["estimator_seconds"] / max(1, out["estimator_calls"]),
             out["estimator_device_ms_per_call"]))
    print("accept rate %.2f; posterior mean %s, std %s; true %s"
          % (out["accept_rate"], np.round(out["posterior_mean"], 3), np.round(out["posterior_std"], 3),
             [model.true_params[k] for k in model.params_to_estimate]))


if __name__ == "__main__":
    main()
