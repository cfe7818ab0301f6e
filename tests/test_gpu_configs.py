"""GPU parity tests at BASELINE.json's own configurations and kernel instantiations (round-2 gaps):

  (a) C2 geometry: N = 2^20, default launch geometry (148 cooperating CTAs, ITEMS = 7), both the
      production flavour (Philox, replayed through the oracle) and the general flavour (injected noise)
  (b) C3 instantiation: batch of N = 4096 filters, one CTA each, ITEMS = 9, shared-memory ancestor ring
  (c) C4 shape: SV and SV-leverage at N = 2^16
  (d) filter-only launches (do_smoother = 0; the bpf_* half of the boundary, alg_type 'mh0')
  (e) the binding INTEGRATION.md documents (examples/flps_binding.py -> pfb200_run)
  (f) the model-consistent leverage index (PFB200_FLAG_SVL_OBS_MODEL)
  (g) fp32 against the fp64 oracle on injected noise: the documented fp32 bound
  (h) the distributed filter's peer paths on ONE GPU (two rank handles in this process)

All through the C ABI; the oracle is the checker.  Large-N cases use the step-locked comparison of
tests/stepwise.py (ancestors bit exact away from verified CDF ties, which are counted and printed)."""
import importlib.util
import os

import numpy as np
import pytest

import golden_util as gu
import stepwise as sw
from oracle import pf_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
THETA = {"lgss": np.array([0.2, 0.5, 1.0, 0.5]), "sv": np.array([0.5, 0.95, 0.25]),
         "sv_leverage": np.array([0.5, 0.95, 0.25, -0.3])}


def _engine(kind, dtype="float64", **opts):
    from qnmh_sysid2018_b200.engine import Engine
    eng = Engine(kind, dtype, 0)
    for k, v in opts.items():
        eng.set_option(k, v)
    return eng


def _obs(kind, T, seed):
    rng = np.random.default_rng(seed)
    th = THETA[kind]
    x, obs = th[0], np.zeros(T + 1)
    for t in range(1, T + 1):
        x = th[0] + th[1] * (x - th[0]) + th[2] * rng.standard_normal()
        obs[t] = x + th[3] * rng.standard_normal() if kind == "lgss" else np.exp(0.5 * x) * rng.standard_normal()
    return obs


def _first(d):
    return {k: v[0] for k, v in d.items()}


# ---------------------------------------------------------------------------------------------- (a)
@pytest.mark.parametrize("flavour", ["production_philox", "general_injected"])
def test_c2_geometry_matches_oracle(flavour):
    """The instantiation bench.py times -- pf_persistent_kernel<LGSS, double, 512, 7, 1, GENERAL=0> on 148
    cooperating CTAs -- and its general twin, at N = 2^20."""
    N, T, L, seed, ev = 1 << 20, 24, 10, 87655678, 4
    obs = _obs("lgss", T, 1)
    eng = _engine("lgss")
    if flavour == "production_philox":
        draws = sw.philox_draws(eng, seed, ev, N, T)
        out = eng.run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=[ev],
                      store_history=True)
    else:
        draws = orc.random_draws(np.random.default_rng(3), N, T)
        out = eng.run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", draws=draws, store_history=True)
    assert eng.info("ctas_per_filter") == 148 and eng.info("items") == 7 and eng.info("cooperative") == 1
    ties = sw.check_history(orc.MODEL_LGSS, THETA["lgss"], obs, _first(eng.fetch_history()),
                            _first(eng.fetch_steps()), _first(out), draws, L)
    print("C2 %s: %d verified CDF-tie mismatches in %d particle-steps" % (flavour, ties, N * T))
    # ring mode (what the bench runs) gives the same numbers as history mode
    if flavour == "production_philox":
        ring = eng.run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=[ev])
        for k in ("log_like", "filt", "smo", "grad_terms", "traj_index"):
            np.testing.assert_array_equal(ring[k], out[k])


# ---------------------------------------------------------------------------------------------- (b)
def test_c3_instantiation_matches_oracle():
    """Batch of one-CTA filters with ITEMS = 9 and the uint16 shared-memory ancestor ring, Philox noise
    replayed through the oracle (free running: at N = 4096 a CDF tie has probability ~1e-9 per step)."""
    B, N, T, L, seed = 12, 4096, 50, 10, 99
    obs = _obs("lgss", T, 2)
    rng = np.random.default_rng(5)
    thetas = THETA["lgss"] + 0.03 * rng.standard_normal((B, 4))
    evals = np.arange(100, 100 + B)
    eng = _engine("lgss", cluster=-1)  # the persistent kernel's batched instantiation (a batch this small would
                                       # otherwise go to the cluster kernel: tests/test_gpu_cluster.py)
    out = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=evals, store_states=True)
    assert eng.info("items") == 9 and eng.info("anc_smem") == 1 and eng.info("ctas_per_filter") == 1
    for b in range(B):
        draws = sw.philox_draws(eng, seed, int(evals[b]), N, T)
        ref = orc.smoother(orc.MODEL_LGSS, thetas[b], obs, N, L, draws)
        assert abs(out["log_like"][b] - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
        np.testing.assert_allclose(out["filt"][b], ref["filt_state_est"].reshape(-1), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["smo"][b], ref["smo_state_est"].reshape(-1), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(out["grad_terms"][b], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
        assert int(out["traj_index"][b]) == ref["trajectory_index"]
        # PFB200_FLAG_STORE_STATES: the reference's row of the particle matrix without the full history
        np.testing.assert_array_equal(out["traj"][b], ref["state_trajectory"].reshape(-1))
    # the plain ring mode (the benchmarked launch) gives the same numbers
    ring = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=evals)
    assert eng.info("items") == 9 and eng.info("anc_smem") == 1
    for k in ("log_like", "filt", "smo", "grad_terms", "traj_index"):
        np.testing.assert_array_equal(ring[k], out[k])
    assert np.all(np.isnan(ring["traj"]))


# ---------------------------------------------------------------------------------------------- (c)
@pytest.mark.parametrize("kind", ["sv", "sv_leverage"])
def test_c4_shape_matches_oracle(kind):
    N, T, L, seed, ev = 1 << 16, 30, 10, 12, 7
    obs = _obs(kind, T, 4)
    eng = _engine(kind)
    draws = sw.philox_draws(eng, seed, ev, N, T)
    out = eng.run(obs, THETA[kind], N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=[ev], store_history=True)
    ties = sw.check_history(gu.MODEL_IDS[kind], THETA[kind], obs, _first(eng.fetch_history()),
                            _first(eng.fetch_steps()), _first(out), draws, L, exact_particles=(kind == "sv"),
                            lw_rtol=1e-12)
    print("C4 %s: G = %d CTAs, %d verified tie mismatches" % (kind, eng.info("ctas_per_filter"), ties))
    ring = eng.run(obs, THETA[kind], N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=[ev])
    for k in ("log_like", "filt", "smo", "grad_terms", "traj_index"):
        np.testing.assert_array_equal(ring[k], out[k])


# ---------------------------------------------------------------------------------------------- (d)
@pytest.mark.parametrize("name", ["lgss_N1000_T500_L10", "sv_N256_T100_L5", "lgss_N2000_T120_L10"])
def test_filter_only_launch_matches_golden_and_oracle(name):
    """do_smoother = 0 (bpf_*): log-likelihood, filtered means and the trajectory draw of the reference's
    golden run; no smoother output."""
    g = gu.load(name)
    eng = _engine(g["model_kind"])
    draws = gu.draws_for(g)
    out = eng.run(g["obs"], g["theta"], g["no_particles"], g["fixed_lag"], False, g["generate_initial_state"],
                  g["initial_state"], g["resampling_method"], draws=draws, store_history=True)
    ref = orc.particle_filter(g["model_id"], g["theta"], g["obs"], g["no_particles"], draws,
                              initial_state_value=g["initial_state"],
                              generate_initial_state=g["generate_initial_state"], resampling=g["resampling_method"])
    hist = eng.fetch_history()
    np.testing.assert_array_equal(hist["ancestors"][0], ref["ancestors"])
    assert gu.sha(hist["ancestors"][0].astype(np.int32)) == str(g["ancestors_sha256"])
    assert abs(out["log_like"][0] - g["log_like"]) <= 1e-10 * abs(g["log_like"])
    np.testing.assert_allclose(out["filt"][0], g["filt_state_est"], rtol=1e-10, atol=1e-12)
    assert int(out["traj_index"][0]) == ref["trajectory_index"]
    np.testing.assert_allclose(out["traj"][0], g["state_trajectory"], rtol=1e-12, atol=1e-13)
    assert not out["smo"].any() and not out["grad_terms"].any()
    # ring mode, Philox: a filter-only launch equals the filter part of a filter + smoother launch
    a = eng.run(g["obs"], g["theta"], 20000, 10, False, True, 0.0, "systematic", seed=3, eval_ids=[1])
    b = eng.run(g["obs"], g["theta"], 20000, 10, True, True, 0.0, "systematic", seed=3, eval_ids=[1])
    for k in ("log_like", "filt", "traj_index"):
        np.testing.assert_array_equal(a[k], b[k])


def test_dropin_filter_method_matches_reference_golden():
    """`ParticleMethodsB200.filter(model)` (what alg_type='mh0' calls, metropolis_hastings.py:328-329)."""
    from qnmh_sysid2018_b200 import ParticleMethodsB200
    from qnmh_sysid2018_b200.models import MODEL_CLASSES
    g = gu.load("lgss_N500_T250_L10")
    model = MODEL_CLASSES["lgss"]()
    for k, v in zip(g["param_names"], g["theta"]):
        model.params[k] = float(v)
    model.set_observations(g["obs"])
    model.fix_true_params()
    model.create_inference_model(tuple(g["params_to_estimate"]))
    pf = ParticleMethodsB200({"no_particles": g["no_particles"], "generate_initial_state": True, "rng": "injected"})
    pf.draws = gu.draws_for(g)
    pf.filter(model)
    r = pf.results
    assert abs(r["log_like"] - g["log_like"]) <= 1e-10 * abs(g["log_like"])
    np.testing.assert_allclose(r["filt_state_est"].reshape(-1), g["filt_state_est"], rtol=1e-10, atol=1e-12)
    # default settings keep the particle rows ('store_states': 'auto'): the reference's state_trajectory
    np.testing.assert_array_equal(r["state_trajectory"].reshape(-1), g["state_trajectory"])
    assert r["state_trajectory"].shape == (1, g["T"] + 1)
    assert "smo_state_est" not in r
    # fixed_lag = 0 without gradients smooths with zero hops (standard.py:146-157): smo == filt
    pf0 = ParticleMethodsB200({"no_particles": g["no_particles"], "generate_initial_state": True,
                               "rng": "injected", "fixed_lag": 0})
    pf0.draws = gu.draws_for(g)
    pf0.smoother(model)
    np.testing.assert_allclose(pf0.results["smo_state_est"][1:], pf0.results["filt_state_est"][1:], rtol=1e-12, atol=1e-13)
    assert not pf0.results["smo_gradient_est"].any()
    pf0.settings["estimate_gradient"] = True
    with pytest.raises(NameError):
        pf0.smoother(model)


# ---------------------------------------------------------------------------------------------- (e)
def test_integration_binding_runs_pfb200_run():
    """Executes examples/flps_binding.py -- the snippet of INTEGRATION.md section 2 -- and checks both
    calls (`flps_lgss` -> pfb200_run(do_smoother=1), `bpf_lgss` -> do_smoother=0) against the oracle."""
    spec = importlib.util.spec_from_file_location("flps_binding", os.path.join(ROOT, "examples", "flps_binding.py"))
    fb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fb)
    N, T, L, seed = 1000, 120, 10, 87655678
    obs = _obs("lgss", T, 9)
    mu, phi, sv, se = THETA["lgss"]
    filt, smo, ll, grad, traj = fb.flps_lgss(obs, mu, phi, sv, se, no_particles=N, fixed_lag=L, seed=seed)  # eval_id 0
    filt_b, ll_b, traj_b = fb.bpf_lgss(obs, mu, phi, sv, se, no_particles=N, seed=seed)                     # eval_id 1
    eng = _engine("lgss")
    ref = orc.smoother(orc.MODEL_LGSS, THETA["lgss"], obs, N, L, sw.philox_draws(eng, seed, 0, N, T))
    assert abs(ll - ref["log_like"]) <= 1e-10 * abs(ref["log_like"])
    np.testing.assert_allclose(filt, ref["filt_state_est"].reshape(-1), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(smo, ref["smo_state_est"].reshape(-1), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(grad.reshape(4, T + 1), ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
    np.testing.assert_array_equal(traj, ref["state_trajectory"].reshape(-1))
    ref_b = orc.particle_filter(orc.MODEL_LGSS, THETA["lgss"], obs, N, sw.philox_draws(eng, seed, 1, N, T))
    assert abs(ll_b - ref_b["log_like"]) <= 1e-10 * abs(ref_b["log_like"])
    np.testing.assert_allclose(filt_b, ref_b["filt_state_est"].reshape(-1), rtol=1e-10, atol=1e-12)
    np.testing.assert_array_equal(traj_b, ref_b["state_trajectory"].reshape(-1))


# ---------------------------------------------------------------------------------------------- (f)
@pytest.mark.parametrize("N,G", [(3000, 1), (40000, 0)])
def test_svl_model_consistent_obs_index(N, G):
    """PFB200_FLAG_SVL_OBS_MODEL: the leverage term reads obs[t-1] (Cython twin / model definition,
    cython_sv_leverage_helper.pyx:254) instead of obs[t] (NumPy twin)."""
    T, L = 40, 6
    obs = _obs("sv_leverage", T, 6)
    draws = orc.random_draws(np.random.default_rng(N), N, T)
    eng = _engine("sv_leverage", ctas_per_filter=G) if G else _engine("sv_leverage")
    out = eng.run(obs, THETA["sv_leverage"], N, L, True, True, 0.0, "systematic", draws=draws, store_history=True,
                  svl_obs_model=True)
    sw.check_history(orc.MODEL_SVL, THETA["sv_leverage"], obs, _first(eng.fetch_history()), _first(eng.fetch_steps()),
                     _first(out), draws, L, exact_particles=False, obs_index="model", lw_rtol=1e-12)
    other = eng.run(obs, THETA["sv_leverage"], N, L, True, True, 0.0, "systematic", draws=draws)
    assert abs(other["log_like"][0] - out["log_like"][0]) > 1e-6  # the switch does something
    # and through the drop-in estimator's setting
    from qnmh_sysid2018_b200 import ParticleMethodsB200
    from qnmh_sysid2018_b200.models import MODEL_CLASSES
    model = MODEL_CLASSES["sv_leverage"]()
    for k, v in zip(model.params.keys(), THETA["sv_leverage"]):
        model.params[k] = float(v)
    model.set_observations(obs)
    model.fix_true_params()
    model.create_inference_model(("mu", "phi", "sigma_v", "rho"))
    pf = ParticleMethodsB200({"no_particles": N, "fixed_lag": L, "generate_initial_state": True, "rng": "injected",
                              "svl_obs_index": "model", "estimate_gradient": True})
    pf.draws = draws
    pf.smoother(model)
    assert pf.results["log_like"] == float(out["log_like"][0])


# ---------------------------------------------------------------------------------------------- (g)
# The documented fp32 bound.  fp32 = storage and per-particle arithmetic (rings,
# normals, propagate, log-weight, exp of the shifted weight); CDF, reductions and outputs stay fp64.
FP32_BOUND = {
    "particle_rel": 4e-7,          # step-locked: x_t against the fp64 propagate of the same parent and noise
    "logw_abs": 1e-6,              # step-locked: log-weight, relative to 1 + |logw|
    "anc_rate_per_N": 2.0 ** -28,  # step-locked: ancestor mismatch rate per particle-step, times 1/N
    "ll_term_abs": 5e-7,           # step-locked: log-likelihood term against fp64 on the stored fp32 log-weights
    "filt_abs": 2e-7,              # step-locked: filtered mean, likewise
    "smo_rel": 1e-7,               # smoother on the device genealogy: smoothed means / score terms, fp64 replay
    "ll_rel_free": 5e-4,           # free running against the fp64 oracle on the same noise (the two filters
                                   # decouple at the first ancestor that differs: a Monte Carlo difference)
}


@pytest.mark.parametrize("N", [1 << 14, 1 << 18])
def test_fp32_documented_bound_on_injected_noise(N):
    """fp32 against the fp64 oracle on the SAME injected noise, step-locked (every time step from the
    generation the device stored), plus the free-running log-likelihood / score difference (printed)."""
    T, L = 30, 10
    th = THETA["lgss"]
    obs = _obs("lgss", T, 8)
    draws = orc.random_draws(np.random.default_rng(N + 1), N, T)
    eng = _engine("lgss", "float32")
    out = _first(eng.run(obs, th, N, L, True, True, 0.0, "systematic", draws=draws, store_history=True))
    hist, steps = _first(eng.fetch_history()), _first(eng.fetch_steps())
    x, lw, anc = hist["particles"], hist["log_weights"], hist["ancestors"]
    w, m, S = sw.normalised_weights(lw)
    mism, worst = 0, dict(x=0.0, lw=0.0)
    for t in range(1, T + 1):
        a_ref = orc.systematic(w[t - 1], draws.u[t])
        d = anc[t] != a_ref
        mism += int(d.sum())
        if d.any():  # a differing ancestor is a near-tie: the position sits within 2^-20 of the CDF steps skipped
            cdf = np.cumsum(w[t - 1])
            for n in np.flatnonzero(d):
                lo, hi = sorted((int(anc[t][n]), int(a_ref[n])))
                assert np.abs(cdf[lo:hi] - (n + draws.u[t]) / N).max() <= 2.0 ** -20
        x_ref = orc.generate_state(orc.MODEL_LGSS, th, obs, x[t - 1][anc[t]], t, draws.z[t])
        worst["x"] = max(worst["x"], float((np.abs(x[t] - x_ref) / (1.0 + np.abs(x_ref))).max()))
        lw_ref = orc.evaluate_obs(orc.MODEL_LGSS, th, obs, x[t], t)
        worst["lw"] = max(worst["lw"], float((np.abs(lw[t] - lw_ref) / (1.0 + np.abs(lw_ref))).max()))
    rate = mism / float(N * T)
    ll_ref = m + np.log(S) - np.log(N)
    ll_ref[0] = 0.0
    ll_err = float(np.abs(steps["ll_terms"] - ll_ref).max())
    filt_ref = (w * x).sum(axis=1)
    filt_ref[0] = 0.0
    filt_err = float(np.abs(out["filt"] - filt_ref).max())
    sm = orc.fixed_lag_smoother(orc.MODEL_LGSS, th, obs, dict(particles=x, weights=w, ancestors=anc,
                                                               filt_state_est=filt_ref.reshape(-1, 1)), L)
    smo_err = float(np.abs(out["smo"] - sm["smo_state_est"].reshape(-1)).max())
    g_scale = np.abs(sm["smo_gradient_est"]).max()
    grad_err = float(np.abs(out["grad_terms"] - sm["smo_gradient_est"]).max() / g_scale)
    ref = orc.smoother(orc.MODEL_LGSS, th, obs, N, L, draws)
    ll_rel = abs(out["log_like"] - ref["log_like"]) / abs(ref["log_like"])
    s32, s64 = out["grad_terms"].sum(axis=1), ref["smo_gradient_est"].sum(axis=1)
    print("fp32 N=%d step-locked: particle rel %.2e, logw %.2e, ancestor mismatch rate %.3e (= N * %.3e), ll term %.2e, "
          "filt %.2e, smo %.2e, score terms %.2e | free running: ll rel %.2e, score %s vs fp64 %s, smo abs %.2e"
          % (N, worst["x"], worst["lw"], rate, rate / N, ll_err, filt_err, smo_err, grad_err, ll_rel,
             np.round(s32, 3), np.round(s64, 3), np.abs(out["smo"] - ref["smo_state_est"].reshape(-1)).max()))
    assert worst["x"] <= FP32_BOUND["particle_rel"] and worst["lw"] <= FP32_BOUND["logw_abs"]
    assert rate <= N * FP32_BOUND["anc_rate_per_N"]
    assert ll_err <= FP32_BOUND["ll_term_abs"] and filt_err <= FP32_BOUND["filt_abs"]
    assert smo_err <= FP32_BOUND["smo_rel"] and grad_err <= FP32_BOUND["smo_rel"]
    assert ll_rel <= FP32_BOUND["ll_rel_free"]
    assert np.all(np.isfinite(s32))


# ---------------------------------------------------------------------------------------------- (h)
@pytest.mark.parametrize("N,T,L,gl,resampling", [(1 << 20, 12, 4, 74, "systematic"), (100001, 20, 3, 16, "systematic"),
                                                 (200000, 16, 5, 32, "stratified"), (60000, 12, 4, 8, "multinomial")])
def test_sharded_filter_two_ranks_on_one_gpu(N, T, L, gl, resampling):
    """The distributed filter's code paths (peer stores of children, halo mirrors, system-scope
    all-gathers, peer lineage walks) with both ranks on THIS GPU: 2 x gl CTAs must reproduce the
    single-handle run with 2 * gl CTAs bit for bit (same CDF tiling, Philox keyed by global index)."""
    from qnmh_sysid2018_b200.distributed import LocalShardedParticleFilter
    obs = _obs("lgss", T, N % 97)
    sh = LocalShardedParticleFilter("lgss", "float64", devices=(0, 0), ctas_per_rank=gl)
    d = sh.run(obs, THETA["lgss"], N, L, seed=11, eval_id=3, resampling=resampling)
    one_eng = _engine("lgss", ctas_per_filter=2 * gl)
    one = one_eng.run(obs, THETA["lgss"], N, L, True, True, 0.0, resampling, seed=11, eval_ids=[3])
    assert one_eng.info("ctas_per_filter") == 2 * gl
    assert d["log_like_per_rank"][0] == d["log_like_per_rank"][1] == one["log_like"][0]
    np.testing.assert_allclose(d["filt"][0], one["filt"][0], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(d["smo"][0], one["smo"][0], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(d["grad_terms"][0], one["grad_terms"][0], rtol=1e-10, atol=1e-11)
    assert int(d["traj_index"][0]) == int(one["traj_index"][0])  # trajectory draw across ranks
    # a second evaluation re-uses the connection (update_params only) and still matches a fresh run
    th2 = THETA["lgss"] * np.array([1.1, 0.9, 1.05, 1.0])
    d2 = sh.run(obs, th2, N, L, seed=11, eval_id=4, resampling=resampling)
    one2 = one_eng.run(obs, th2, N, L, True, True, 0.0, resampling, seed=11, eval_ids=[4])
    assert d2["log_like_per_rank"][0] == one2["log_like"][0]
    np.testing.assert_allclose(d2["grad_terms"][0], one2["grad_terms"][0], rtol=1e-10, atol=1e-11)
    sh.close()


# ---------------------------------------------------------------------------------------------- (i)
@pytest.mark.parametrize("N,G", [(6000, 12), (4096, 1)])
def test_degenerate_generation_matches_oracle(N, G):
    """An observation that sends every log-weight to -inf (a derailed MH proposal does the same): all weights
    of that generation are NaN, the reference's search leaves every ancestor at particle 0
    (resampling.pyx:78-83) and the log-likelihood is NaN; the particles stay finite.  Device = oracle."""
    import warnings
    T, L = 12, 3
    obs = _obs("lgss", T, 8)
    obs[5] = 1e200
    draws = orc.random_draws(np.random.default_rng(N), N, T)
    eng = _engine("lgss", ctas_per_filter=G, cluster=-1)
    eng.configure(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", draws=draws, store_history=True)
    eng.execute()
    hist, out = _first(eng.fetch_history()), eng.fetch()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = orc.smoother(orc.MODEL_LGSS, THETA["lgss"], obs, N, L, draws)
    assert eng.info("ctas_per_filter") == G
    assert np.isnan(out["log_like"][0]) and np.isnan(ref["log_like"])
    assert np.all(ref["ancestors"][6] == 0)
    np.testing.assert_array_equal(hist["ancestors"][1:], ref["ancestors"][1:])
    np.testing.assert_array_equal(hist["particles"], ref["particles"])
    assert np.all(np.isfinite(hist["particles"]))


def test_degenerate_generations_do_not_serialise_the_step():
    """With NaN weights the children ranges used to hand all N children to CTA 0 (11x the step time at C4's
    shape); every CTA now takes its own index range."""
    N, T, L = 1 << 16, 60, 10
    obs = _obs("sv", T, 3)
    ms = {}
    for name, bad in (("regular", False), ("derailed", True)):
        o = obs.copy()
        if bad:
            o[3::4] = 1e200  # every 4th generation degenerates
        eng = _engine("sv", cluster=-1)
        eng.configure(o, THETA["sv"], N, L, True, True, 0.0, "systematic", seed=5)
        for _ in range(3):
            eng.execute()
            eng.synchronize()
        ms[name] = eng.last_kernel_ms()
        assert eng.info("ctas_per_filter") >= 64
        ll = eng.fetch_log_like()[0]
        assert np.isnan(ll) == bad
    print("kernel ms: %s" % ms)
    assert ms["derailed"] < 2.5 * ms["regular"], ms
