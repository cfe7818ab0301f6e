"""Round-2 fixtures from the UNMODIFIED reference (build container only; needs /root/reference):

    python tests/golden/make_golden_r2.py [kalman] [chains] [timing]

  kalman  -> tests/golden/kalman_rts_lgss.npz   Kalman RTS smoother, exact score / Hessian estimate
             (state/kalman_methods/standard.py:86-178) for the oracle restatement
  chains  -> tests/golden/mh_chains_ref.npz     short MetropolisHastings chains of the reference
             (parameter/mcmc/metropolis_hastings.py): LGSS qMH on the NumPy particle smoother and on
             the Kalman smoother (example-2 settings, T = 250), SV-leverage qMH on synthetic data
             (example-3 settings); posterior summaries the B200 chains are compared with
  timing  -> profiles/r2_reference_cpu_timing.json   the reference's own CPU path timed here:
             NumPy ParticleMethods.smoother at N = 1000/2000, T = 250/500 and Cython flps_lgss
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

SEED = 87655678
F500 = "linear_gaussian_model_T500_midSNR.csv"


def lgss_model(ref, T, theta=(0.2, 0.5, 1.0, 0.5), data_file=F500, estimate=("mu", "phi", "sigma_v")):
    m = ref["LinearGaussianModel"]()
    m.params["mu"], m.params["phi"], m.params["sigma_v"], m.params["sigma_e"] = theta
    m.no_obs = T
    m.initial_state = 0.0
    m.import_data(file_name=os.path.join(ref["data_dir"], data_file))
    m.fix_true_params()
    m.create_inference_model(params_to_estimate=estimate)
    return m


def svl_model(ref, T, theta=(0.5, 0.95, 0.25, -0.3), seed=12345):
    m = ref["StochasticVolatilityModelLeverage"]()
    m.params["mu"], m.params["phi"], m.params["sigma_v"], m.params["rho"] = theta
    m.no_obs = T
    m.initial_state = theta[0]
    np.random.seed(seed)
    m.generate_data()
    m.fix_true_params()
    m.create_inference_model(params_to_estimate=("mu", "phi", "sigma_v", "rho"))
    return m


def mh_settings(dim, no_iters, burn_in, initial_params, step_size=0.5):
    """example2_lgss_particles.py:39-60 / example3_stochastic_volatility.py:40-60, qMH variant
    (`qn_strategy='bfgs'`, base Hessian 0.01^2 I)."""
    return {'no_iters': no_iters, 'no_burnin_iters': burn_in, 'step_size': step_size,
            'base_hessian': 0.01 ** 2 * np.eye(dim), 'initial_params': tuple(initial_params),
            'verbose': False, 'verbose_wait_enter': False, 'trust_region_size': None,
            'hessian_estimate': None, 'hessian_correction': 'replace',
            'hessian_correction_verbose': False, 'no_iters_between_progress_reports': 100000,
            'qn_memory_length': 20, 'qn_initial_hessian': 'scaled_gradient', 'qn_strategy': 'bfgs',
            'qn_bfgs_curvature_cond': 'damped', 'qn_initial_hessian_scaling': 0.01,
            'qn_initial_hessian_fixed': np.eye(dim) * 0.01 ** 2, 'qn_only_accepted_info': True,
            'qn_accept_all_initial': True}


def kalman(ref):
    rows = {}
    for tag, T, fn, theta in (("T250_mid", 250, F500, (0.2, 0.5, 1.0, 0.5)),
                              ("T1000_good", 1000, "linear_gaussian_model_T1000_goodSNR.csv", (0.2, 0.5, 1.0, 0.1))):
        m = lgss_model(ref, T, theta, fn)
        kf = ref["KalmanMethods"]({"initial_state": 0.0, "initial_cov": 1e-5, "estimate_gradient": True,
                                   "estimate_hessian": True})
        kf.results = {}
        kf.smoother(m)
        r = kf.results
        rows["obs_" + tag] = np.asarray(m.obs).reshape(-1)
        rows["theta_" + tag] = np.asarray(theta)
        rows["ll_" + tag] = float(r["log_like"])
        for key in ("filt_state_est", "filt_state_cov", "pred_state_est", "pred_state_cov", "smo_state_est",
                    "smo_state_cov", "gradient_internal", "hessian_internal", "log_joint_gradient_estimate",
                    "log_joint_hessian_estimate"):
            rows[key + "_" + tag] = np.asarray(r[key], dtype=np.float64).reshape(np.shape(r[key])).squeeze()
        rows["log_prior_gradient_" + tag] = np.array(list(m.log_prior_gradient().values()), dtype=np.float64)
        rows["log_prior_hessian_" + tag] = np.array(list(m.log_prior_hessian().values()), dtype=np.float64)
        print("kalman rts %-11s ll=%.8f score=%s" % (tag, rows["ll_" + tag], rows["gradient_internal_" + tag]))
    np.savez_compressed(os.path.join(HERE, "kalman_rts_lgss.npz"), **rows)


def _chain_summary(mh, burn_in):
    n = mh.settings['no_iters']
    params = np.asarray(mh.params[:n], dtype=np.float64)
    acc = np.asarray(mh.accepted[:n], dtype=np.float64).reshape(-1)
    return dict(params=params, accepted=acc, log_like=np.asarray(mh.log_like[:n], dtype=np.float64).reshape(-1),
                posterior_mean=params[burn_in:].mean(axis=0), posterior_std=params[burn_in:].std(axis=0),
                accept_rate=float(acc[burn_in:].mean()), time_per_iteration=float(mh.time_per_iteration))


def chains(ref):
    rows = {}

    def keep(tag, summ, **meta):
        for k, v in summ.items():
            rows["%s_%s" % (tag, k)] = np.asarray(v)
        for k, v in meta.items():
            rows["%s_%s" % (tag, k)] = np.asarray(v)
        print("%-14s accept %.3f  mean %s  std %s  %.3f s/iteration" % (
            tag, summ["accept_rate"], np.round(summ["posterior_mean"], 4), np.round(summ["posterior_std"], 4),
            summ["time_per_iteration"]), flush=True)

    # LGSS, Kalman smoother (exact likelihood and score): the low-noise posterior
    m = lgss_model(ref, 250)
    np.random.seed(SEED)
    mh = ref["MetropolisHastings"](m, "qmh", mh_settings(3, 6000, 1000, (0.2, 0.5, 1.0)))
    kf = ref["KalmanMethods"]({"initial_state": 0.0, "initial_cov": 1e-5, "estimate_gradient": True,
                               "estimate_hessian": True})
    kf.results = {}
    mh.run(kf)
    keep("lgss_kalman", _chain_summary(mh, 1000), obs=np.asarray(m.obs).reshape(-1), burn_in=1000)

    # LGSS, NumPy particle smoother (the estimator the drop-in replaces), example-2 particle settings
    m = lgss_model(ref, 250)
    np.random.seed(SEED)
    mh = ref["MetropolisHastings"](m, "qmh", mh_settings(3, 1600, 400, (0.2, 0.5, 1.0)))
    pf = ref["ParticleMethods"]({'no_particles': 1000, 'resampling_method': 'systematic', 'fixed_lag': 10,
                                 'initial_state': 0.0, 'generate_initial_state': True,
                                 'estimate_gradient': True, 'estimate_hessian': True})
    pf.results = {}
    mh.run(pf)
    keep("lgss_pf", _chain_summary(mh, 400), obs=np.asarray(m.obs).reshape(-1), burn_in=400,
         no_particles=1000, fixed_lag=10)

    # SV with leverage, synthetic data, example-3 settings (Quandl data is not available offline)
    m = svl_model(ref, 200)
    np.random.seed(SEED)
    mh = ref["MetropolisHastings"](m, "qmh", mh_settings(4, 1200, 300, (0.5, 0.95, 0.25, -0.3)))
    pf = ref["ParticleMethods"]({'no_particles': 500, 'resampling_method': 'systematic', 'fixed_lag': 10,
                                 'initial_state': 0.0, 'generate_initial_state': True,
                                 'estimate_gradient': True, 'estimate_hessian': True})
    pf.results = {}
    mh.run(pf)
    keep("svl_pf", _chain_summary(mh, 300), obs=np.asarray(m.obs).reshape(-1), burn_in=300,
         no_particles=500, fixed_lag=10, theta=np.array([0.5, 0.95, 0.25, -0.3]))
    np.savez_compressed(os.path.join(HERE, "mh_chains_ref.npz"), **rows)


def timing(ref):
    out = {"host": {"cores_used": 1, "cpu_count": os.cpu_count()},
           "note": "unmodified reference through tests/golden/ref_harness.py, single process, build container"}
    rows = []
    for N, T in ((1000, 250), (1000, 500), (2000, 250), (2000, 500), (4096, 500)):
        m = lgss_model(ref, T)
        pf = ref["ParticleMethods"]({'no_particles': N, 'resampling_method': 'systematic', 'fixed_lag': 10,
                                     'initial_state': 0.0, 'generate_initial_state': True,
                                     'estimate_gradient': True, 'estimate_hessian': True})
        pf.results = {}
        np.random.seed(SEED)
        pf.smoother(m)  # warm-up
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            pf.smoother(m)
        dt = (time.perf_counter() - t0) / reps
        rows.append({"impl": "NumPy ParticleMethods.smoother (standard.py:121)", "N": N, "T": T, "L": 10,
                     "seconds": dt, "particle_timesteps_per_s": N * T / dt})
        print(rows[-1], flush=True)
    m = lgss_model(ref, 500)
    cy = ref["ParticleMethodsCythonLGSS"]({'no_particles': 1000, 'fixed_lag': 10, 'estimate_gradient': True,
                                           'estimate_hessian': True})
    cy.results = {}
    cy.smoother(m)
    reps = 5
    t0 = time.perf_counter()
    for _ in range(reps):
        cy.smoother(m)
    dt = (time.perf_counter() - t0) / reps
    rows.append({"impl": "Cython flps_lgss (cython_lgss_helper.pyx:144; N=1000, T=500, L=10 compiled in)",
                 "N": 1000, "T": 500, "L": 10, "seconds": dt, "particle_timesteps_per_s": 1000 * 500 / dt})
    print(rows[-1], flush=True)
    # the O(N T L) NumPy port the bench's CPU arm times, same sizes, for the "port is conservative" claim
    sys.path.insert(0, ROOT)
    from oracle import pf_oracle as orc
    for N, T in ((1000, 500), (2000, 500), (4096, 500)):
        obs = np.asarray(lgss_model(ref, T).obs).reshape(-1)
        draws = orc.random_draws(np.random.default_rng(1), N, T)
        orc.smoother(orc.MODEL_LGSS, (0.2, 0.5, 1.0, 0.5), obs, N, 10, draws)
        t0 = time.perf_counter()
        for _ in range(3):
            orc.smoother(orc.MODEL_LGSS, (0.2, 0.5, 1.0, 0.5), obs, N, 10, draws)
        dt = (time.perf_counter() - t0) / 3
        rows.append({"impl": "oracle port pf_oracle.smoother (O(N T L))", "N": N, "T": T, "L": 10, "seconds": dt,
                     "particle_timesteps_per_s": N * T / dt})
        print(rows[-1], flush=True)
    out["rows"] = rows
    with open(os.path.join(ROOT, "profiles", "r2_reference_cpu_timing.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    what = sys.argv[1:] or ["kalman", "chains", "timing"]
    ref = ref_harness.load_reference()
    if "kalman" in what:
        kalman(ref)
    if "timing" in what:
        timing(ref)
    if "chains" in what:
        chains(ref)
