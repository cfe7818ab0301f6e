"""CPU tier: host-side logic of the drop-in estimator (results dict, Hessian forms, prior folding,
model mirror) against the reference's golden vectors, with the oracle-backed TEST engine standing
in for the GPU (the CUDA path itself is checked by the `gpu` tests)."""
import numpy as np
import pytest

import golden_util as gu
from oracle import pf_oracle as orc
from oracle_engine import OracleEngine
from qnmh_sysid2018_b200 import ParticleMethodsB200, model_kind_of
from qnmh_sysid2018_b200.models import MODEL_CLASSES
from qnmh_sysid2018_b200.particle_methods import segal_weinstein_cython, segal_weinstein_numpy


def _model_for(g):
    model = MODEL_CLASSES[g["model_kind"]]()
    for k, v in zip(g["param_names"], g["theta"]):
        model.params[k] = float(v)
    model.set_observations(g["obs"])
    model.fix_true_params()
    model.create_inference_model(tuple(g["params_to_estimate"]))
    return model


@pytest.mark.parametrize("name", gu.golden_names())
def test_estimator_results_match_reference(name):
    g = gu.load(name)
    model = _model_for(g)
    np.testing.assert_allclose(list(model.log_prior_gradient().values()), g["log_prior_gradient"], rtol=1e-13)
    np.testing.assert_allclose(list(model.log_prior_hessian().values()), g["log_prior_hessian"], rtol=1e-13)
    np.testing.assert_array_equal(model.params_to_estimate_idx, g["params_to_estimate_idx"])
    pf = ParticleMethodsB200({"no_particles": g["no_particles"], "fixed_lag": g["fixed_lag"],
                              "generate_initial_state": g["generate_initial_state"],
                              "initial_state": g["initial_state"], "estimate_gradient": True,
                              "estimate_hessian": True, "rng": "injected", "store_history": True,
                              "resampling_method": g["resampling_method"]},
                             engine=OracleEngine(g["model_kind"]))
    pf.draws = gu.draws_for(g)
    pf.smoother(model)
    r = pf.results
    assert abs(r["log_like"] - g["log_like"]) <= 1e-12 * abs(g["log_like"])
    np.testing.assert_allclose(r["gradient_internal"], g["gradient_internal"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(r["hessian_internal"], g["hessian_internal"], rtol=1e-11, atol=1e-11)
    np.testing.assert_allclose(r["log_joint_gradient_estimate"], g["log_joint_gradient_estimate"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(r["log_joint_hessian_estimate"], g["log_joint_hessian_estimate"], rtol=1e-11, atol=1e-11)
    np.testing.assert_array_equal(r["state_trajectory"].reshape(-1), g["state_trajectory"])
    assert r["state_trajectory"].shape == (1, g["T"] + 1)
    assert r["smo_state_est"].shape == (g["T"] + 1, 1)
    assert set(r["gradient"].keys()) == set(g["params_to_estimate"])
    assert pf.ancestors.shape == (g["no_particles"], g["T"] + 1)


def test_filter_only_and_settings_surface():
    g = gu.load("lgss_N256_T50_L5")
    pf = ParticleMethodsB200({"no_particles": 256, "generate_initial_state": True, "rng": "injected"},
                             engine=OracleEngine("lgss"))
    for key in ("no_particles", "resampling_method", "fixed_lag", "initial_state",
                "generate_initial_state", "estimate_gradient", "verbose"):
        assert key in pf.settings  # the reference's keys (standard.py:31-38)
    pf.draws = gu.draws_for(g)
    pf.filter(_model_for(g))
    assert abs(pf.results["log_like"] - g["log_like"]) <= 1e-12 * abs(g["log_like"])
    assert "smo_state_est" not in pf.results
    with pytest.raises(NameError):  # fixed_lag = 0 cannot smooth (standard.py:161)
        pf.smoother(_model_for(g))
    pf.settings["resampling_method"] = "residual"
    with pytest.raises(ValueError):
        pf.filter(_model_for(g))


def test_hessian_forms():
    rng = np.random.default_rng(0)
    G = rng.standard_normal((4, 51))
    H = segal_weinstein_numpy(G, 51)
    np.testing.assert_allclose(H, (1 - 1 / 51.0) * G @ G.T, rtol=1e-13)
    s = G.sum(axis=1)
    np.testing.assert_allclose(segal_weinstein_cython(G, 50), G @ G.T - np.outer(s, s) / 50, rtol=1e-13)
    np.testing.assert_allclose(orc.hessian_cython_form(G, 50), segal_weinstein_cython(G, 50), rtol=1e-13)


def test_model_mirror_transforms_and_dispatch():
    m = MODEL_CLASSES["sv_leverage"]()
    m.params.update(mu=0.5, phi=0.95, sigma_v=0.25, rho=-0.3)
    m.fix_true_params()
    m.create_inference_model(("mu", "phi", "sigma_v", "rho"))
    m.transform_params_to_free()
    free = m.get_free_params()
    np.testing.assert_allclose(free, [0.5, np.arctanh(0.95), np.log(0.25), np.arctanh(-0.3)])
    m.store_free_params(free + 0.01)
    np.testing.assert_allclose(m.params["phi"], np.tanh(free[1] + 0.01))
    assert m.check_parameters()
    assert model_kind_of(m) == "sv_leverage"

    class LinearGaussianModel(object):  # a reference-side class is recognised by its name
        pass
    assert model_kind_of(LinearGaussianModel()) == "lgss"
    with pytest.raises(TypeError):
        model_kind_of(object())
    m.generate_data(np.random.RandomState(1), no_obs=30)
    assert m.obs.shape == (31, 1) and m.obs[0, 0] == 0.0


@pytest.mark.reference
def test_dropin_in_reference_metropolis_hastings():
    """The unmodified reference MH driver (parameter/mcmc/metropolis_hastings.py) runs a qMH chain
    on the drop-in estimator: same calls (`smoother(model)`), same `results` keys consumed."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import ref_harness
    ref = ref_harness.load_reference()
    m = ref["LinearGaussianModel"]()
    m.params.update(mu=0.2, phi=0.5, sigma_v=1.0, sigma_e=0.5)
    m.no_obs = 60
    m.initial_state = 0.0
    m.import_data(file_name=os.path.join(ref["data_dir"], "linear_gaussian_model_T500_midSNR.csv"))
    m.fix_true_params()
    m.create_inference_model(params_to_estimate=("mu", "phi", "sigma_v"))
    pf = ParticleMethodsB200({"no_particles": 200, "fixed_lag": 5, "generate_initial_state": True,
                              "estimate_gradient": True, "estimate_hessian": True},
                             engine=OracleEngine("lgss"))
    settings = {"no_iters": 12, "no_burnin_iters": 2, "step_size": 0.5,
                "base_hessian": np.eye(3) * 0.05 ** 2, "initial_params": (0.2, 0.5, 1.0),
                "verbose": False, "verbose_wait_enter": False, "trust_region_size": None,
                "hessian_estimate": None, "hessian_correction": "replace",
                "hessian_correction_verbose": False, "no_iters_between_progress_reports": 100,
                "qn_memory_length": 20, "qn_initial_hessian": "scaled_gradient",
                "qn_strategy": "bfgs", "qn_bfgs_curvature_cond": "damped",
                "qn_initial_hessian_scaling": 0.01, "qn_initial_hessian_fixed": np.eye(3) * 0.01 ** 2,
                "qn_only_accepted_info": True, "qn_accept_all_initial": True}
    np.random.seed(1)
    mh = ref["MetropolisHastings"](m, "qmh", settings)
    mh.run(pf)
    assert np.all(np.isfinite(mh.log_like[:12]))
    assert np.all(np.isfinite(mh.params[:12]))
    assert pf.eval_id >= 12  # one estimator call per MH iteration (metropolis_hastings.py:413-416)
