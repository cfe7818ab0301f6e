"""Host logic of the lock-step chain pool (MH-driver batching) and of the result writer
(8 f4) on the CPU tier: the compute engine is the oracle-backed stand-in of tests/oracle_engine.py."""
import os

import numpy as np
import pytest

from oracle_engine import OracleEngine
from qnmh_sysid2018_b200 import LinearGaussianModel, ParticleMethodsB200
from qnmh_sysid2018_b200.lockstep import LockstepEstimatorPool
from qnmh_sysid2018_b200 import results_io

SETTINGS = {"no_particles": 64, "fixed_lag": 4, "generate_initial_state": True, "estimate_gradient": True,
            "estimate_hessian": True}


def _model(seed=3, T=25):
    m = LinearGaussianModel()
    m.params.update(mu=0.2, phi=0.5, sigma_v=1.0, sigma_e=0.5)
    m.initial_state = 0.0
    m.generate_data(np.random.RandomState(seed), no_obs=T)
    m.fix_true_params()
    m.create_inference_model(("mu", "phi", "sigma_v"))
    return m


def _toy_chain(k, n_calls, use_filter_every=0):
    """A caller with the MH driver's access pattern: perturb the parameters, call the estimator, read
    `results`.  Chains have different lengths and call patterns on purpose."""
    def run(estimator):
        model = _model()
        rng = np.random.default_rng(100 + k)
        trace = []
        for i in range(n_calls):
            model.store_params(np.array([0.2, 0.5, 1.0]) + 0.05 * rng.standard_normal(3))
            if use_filter_every and i % use_filter_every == 0:
                estimator.filter(model)
                trace.append((estimator.results["log_like"], None))
            else:
                estimator.smoother(model)
                trace.append((estimator.results["log_like"], estimator.results["gradient_internal"].copy()))
        return trace
    return run


def test_lockstep_pool_equals_sequential_estimators():
    lengths = [5, 8, 6, 3]
    pool = LockstepEstimatorPool(4, SETTINGS, engine_factory=OracleEngine)
    batched = pool.run([_toy_chain(k, n, use_filter_every=(3 if k == 1 else 0)) for k, n in enumerate(lengths)])
    assert pool.evaluations == sum(lengths)
    assert pool.batches == max(lengths)  # every launch serves all chains that are still alive
    for k, n in enumerate(lengths):
        pf = ParticleMethodsB200(SETTINGS, engine=OracleEngine("lgss"))
        pf.eval_id = k << 32
        seq = _toy_chain(k, n, use_filter_every=(3 if k == 1 else 0))(pf)
        for (ll_b, g_b), (ll_s, g_s) in zip(batched[k], seq):
            assert ll_b == ll_s
            if g_s is not None:
                np.testing.assert_array_equal(g_b, g_s)


def test_lockstep_pool_propagates_a_failing_chain():
    pool = LockstepEstimatorPool(2, SETTINGS, engine_factory=OracleEngine)

    def bad(estimator):
        estimator.smoother(_model())
        raise KeyError("chain 1 gives up")

    with pytest.raises(KeyError):
        pool.run([_toy_chain(0, 4), bad])
    assert pool.evaluations == 5  # the surviving chain finished its four calls


def test_result_files_have_the_reference_layout(tmp_path):
    """file_system.py:40-75 / output.py:147-222: four gzip JSON documents under <path>/<sim_name>/,
    arrays as nested lists, burn-in removed."""
    model = _model()
    n, burn = 30, 10
    chain = {"params": np.arange(n * 3, dtype=float).reshape(n, 3), "accepted": np.ones((n, 1)),
             "prop_params": np.zeros((n, 3)), "time_per_iteration": 0.01}
    settings = {"no_iters": n, "no_burnin_iters": burn, "step_size": 0.5, "base_hessian": np.eye(3)}
    files = results_io.store_results_to_file(chain, model, settings, "qmh (B200)", str(tmp_path), "example2_test",
                                             sim_desc="unit test")
    assert [os.path.basename(f) for f in files] == ["mcmc_output.json.gz", "data.json.gz", "settings.json.gz",
                                                    "description.txt.gz"]
    out = results_io.read_json(files[0])
    assert np.asarray(out["params"]).shape == (n - burn, 3) and out["params"][0] == [30.0, 31.0, 32.0]
    assert out["states"] is None and out["simulation_name"] == "example2_test"
    data = results_io.read_json(files[1])
    assert np.asarray(data["observations"]).shape == (model.no_obs + 1, 1)
    st = results_io.read_json(files[2])
    assert st["sampler_name"] == "qmh (B200)" and st["base_hessian"][1] == [0.0, 1.0, 0.0]
    assert set(results_io.read_json(files[3])) == {"description", "time"}
