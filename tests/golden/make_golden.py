"""Generates tests/golden/*.npz by running the UNMODIFIED reference (NumPy path:
state/particle_methods/standard.py + resampling.pyx + models/*.py + base_state_inference.py, and
the Kalman filter state/kalman_methods/standard.py as analytic cross-check) on seeded inputs.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

Each fixture stores the inputs (observations, parameters, settings, legacy numpy seed) and the
reference's outputs.  The random numbers are NOT stored: the legacy `np.random.seed` stream is
frozen across numpy versions, and `oracle.pf_oracle.draws_from_legacy_seed` regenerates the exact
draws in the order the reference consumed them.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_harness  # noqa: E402

SEED = 87655678  # helper_linear_gaussian.py:33


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _lgss_model(ref, T, theta, data_file, estimate=("mu", "phi", "sigma_v")):
    m = ref["LinearGaussianModel"]()
    m.params["mu"], m.params["phi"], m.params["sigma_v"], m.params["sigma_e"] = theta
    m.no_obs = T
    m.initial_state = 0.0
    m.import_data(file_name=os.path.join(ref["data_dir"], data_file))
    m.fix_true_params()
    m.create_inference_model(params_to_estimate=estimate)
    return m


def _sv_model(ref, leverage, T, theta, data_seed=12345):
    cls = ref["StochasticVolatilityModelLeverage" if leverage else "StochasticVolatilityModel"]
    m = cls()
    names = ("mu", "phi", "sigma_v", "rho") if leverage else ("mu", "phi", "sigma_v")
    for k, v in zip(names, theta):
        m.params[k] = v
    m.no_obs = T
    m.initial_state = theta[0]
    if leverage:
        m.obs = np.zeros((T + 1, 1))  # generate_data reads obs[i] before writing it (Q10)
    np.random.seed(data_seed)
    m.generate_data()
    m.fix_true_params()
    m.create_inference_model(params_to_estimate=names)
    return m


def _run(ref, name, model, model_kind, settings, seed=SEED, keep_full=True):
    pf = ref["ParticleMethods"](dict(settings))
    pf.results = {}  # Q12: the class attribute is shared between instances
    np.random.seed(seed)
    pf.smoother(model)
    res = pf.results
    anc = pf.ancestors.astype(np.int32)  # (N, T+1)
    out = dict(
        model_kind=model_kind,
        obs=np.asarray(model.obs, dtype=np.float64).reshape(-1),
        theta=np.asarray(model.get_all_params(), dtype=np.float64),
        param_names=np.array(list(model.params.keys())),
        params_to_estimate=np.array(list(model.params_to_estimate)),
        params_to_estimate_idx=np.asarray(model.params_to_estimate_idx, dtype=np.int64),
        no_particles=settings["no_particles"], fixed_lag=settings["fixed_lag"],
        generate_initial_state=bool(settings.get("generate_initial_state", False)),
        initial_state=float(settings.get("initial_state", 0.0)),
        resampling_method=settings.get("resampling_method", "systematic"),
        seed=seed,
        log_like=float(res["log_like"]),
        filt_state_est=np.asarray(res["filt_state_est"]).reshape(-1),
        smo_state_est=np.asarray(res["smo_state_est"]).reshape(-1),
        state_trajectory=np.asarray(res["state_trajectory"]).reshape(-1),
        gradient_internal=np.asarray(res["gradient_internal"], dtype=np.float64),
        hessian_internal=np.asarray(res["hessian_internal"], dtype=np.float64),
        log_joint_gradient_estimate=np.asarray(res["log_joint_gradient_estimate"], dtype=np.float64),
        log_joint_hessian_estimate=np.asarray(res["log_joint_hessian_estimate"], dtype=np.float64),
        log_prior_gradient=np.array([model.log_prior_gradient()[k] for k in model.params]),
        log_prior_hessian=np.array([model.log_prior_hessian()[k] for k in model.params]),
        ancestors_sha256=_sha(anc.T),  # time-major (T+1, N) int32
        particles_sha256=_sha(pf.particles.T),
        weights_last=pf.weights[:, -1].copy(),
        particles_last=pf.particles[:, -1].copy(),
    )
    if keep_full:
        out["ancestors"] = anc.T.copy()
        out["particles"] = pf.particles.T.copy()
    else:
        out["ancestors_every25"] = anc.T[::25].copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-34s ll=%.12f  grad=%s" % (name, out["log_like"], out["gradient_internal"]))


def _kalman(ref):
    rows = {}
    for tag, T, fn, theta in (
            ("T250_mid", 250, "linear_gaussian_model_T500_midSNR.csv", (0.2, 0.5, 1.0, 0.5)),
            ("T500_mid", 500, "linear_gaussian_model_T500_midSNR.csv", (0.2, 0.5, 1.0, 0.5)),
            ("T1000_mid", 1000, "linear_gaussian_model_T1000_midSNR.csv", (0.2, 0.5, 1.0, 0.5)),
            ("T1000_good", 1000, "linear_gaussian_model_T1000_goodSNR.csv", (0.2, 0.5, 1.0, 0.1))):
        m = _lgss_model(ref, T, theta, fn)
        kf = ref["KalmanMethods"]({"initial_state": 0.0, "initial_cov": 1e-5,
                                   "estimate_gradient": True, "estimate_hessian": True})
        kf.results = {}
        kf.smoother(m)
        rows["ll_" + tag] = float(kf.results["log_like"])
        rows["score_" + tag] = np.asarray(kf.results["gradient_internal"], dtype=np.float64)
        rows["obs_" + tag] = np.asarray(m.obs).reshape(-1)
        rows["theta_" + tag] = np.asarray(theta)
        print("kalman %-12s ll=%.10f score=%s" % (tag, rows["ll_" + tag], rows["score_" + tag]))
    np.savez_compressed(os.path.join(HERE, "kalman_lgss.npz"), **rows)


def main():
    ref = ref_harness.load_reference()
    base = {"resampling_method": "systematic", "generate_initial_state": True,
            "estimate_gradient": True, "estimate_hessian": True}
    th = (0.2, 0.5, 1.0, 0.5)
    f500 = "linear_gaussian_model_T500_midSNR.csv"

    _run(ref, "lgss_N256_T50_L5", _lgss_model(ref, 50, th, f500), "lgss",
         dict(base, no_particles=256, fixed_lag=5))
    _run(ref, "lgss_N1000_T500_L10", _lgss_model(ref, 500, th, f500), "lgss",
         dict(base, no_particles=1000, fixed_lag=10), keep_full=False)
    _run(ref, "lgss_N500_T250_L10", _lgss_model(ref, 250, th, f500), "lgss",
         dict(base, no_particles=500, fixed_lag=10), keep_full=False)
    _run(ref, "lgss_N2000_T120_L10", _lgss_model(ref, 120, th, f500), "lgss",
         dict(base, no_particles=2000, fixed_lag=10), keep_full=False)
    _run(ref, "lgss_fixedinit_N128_T40_L3", _lgss_model(ref, 40, th, f500), "lgss",
         dict(base, no_particles=128, fixed_lag=3, generate_initial_state=False, initial_state=0.3))
    _run(ref, "lgss_ragged_N77_T33_L40", _lgss_model(ref, 33, th, f500), "lgss",
         dict(base, no_particles=77, fixed_lag=40))  # lag longer than the series
    _run(ref, "lgss_L1_N300_T60", _lgss_model(ref, 60, th, f500), "lgss",
         dict(base, no_particles=300, fixed_lag=1))
    _run(ref, "lgss_goodsnr_N512_T100_L10",
         _lgss_model(ref, 100, (0.2, 0.5, 1.0, 0.1), "linear_gaussian_model_T1000_goodSNR.csv"),
         "lgss", dict(base, no_particles=512, fixed_lag=10))
    _run(ref, "lgss_est2_N200_T50_L5",
         _lgss_model(ref, 50, th, f500, estimate=("phi", "sigma_v")), "lgss",
         dict(base, no_particles=200, fixed_lag=5))
    _run(ref, "lgss_stratified_N200_T30_L4", _lgss_model(ref, 30, th, f500), "lgss",
         dict(base, no_particles=200, fixed_lag=4, resampling_method="stratified"))
    _run(ref, "lgss_multinomial_N200_T30_L4", _lgss_model(ref, 30, th, f500), "lgss",
         dict(base, no_particles=200, fixed_lag=4, resampling_method="multinomial"))
    _run(ref, "sv_N256_T100_L5", _sv_model(ref, False, 100, (0.5, 0.95, 0.25)), "sv",
         dict(base, no_particles=256, fixed_lag=5))
    _run(ref, "svl_N256_T100_L5", _sv_model(ref, True, 100, (0.5, 0.95, 0.25, -0.3)),
         "sv_leverage", dict(base, no_particles=256, fixed_lag=5))
    _run(ref, "sv_N1500_T200_L10", _sv_model(ref, False, 200, (0.5, 0.95, 0.25)), "sv",
         dict(base, no_particles=1500, fixed_lag=10), keep_full=False)
    _run(ref, "svl_N1500_T200_L10", _sv_model(ref, True, 200, (0.5, 0.95, 0.25, -0.3)),
         "sv_leverage", dict(base, no_particles=1500, fixed_lag=10), keep_full=False)
    _kalman(ref)


if __name__ == "__main__":
    main()
