"""GPU parity tests: the CUDA path (through the C ABI, libpfb200.so) against the CPU oracle on the
same seeded inputs and against the reference's golden vectors.

Bars: ancestors (int32) bit exact; particles bit exact for the LGSS model with injected noise (the
fp64 kernels pin the reference's operation order); log-likelihood, filtered/smoothed means, score
terms: relative 1e-10 (the north star asks for 1e-5 in fp64)."""
import numpy as np
import pytest

import golden_util as gu
from oracle import pf_oracle as orc

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _engine(kind, dtype="float64", **opts):
    from qnmh_sysid2018_b200.engine import Engine
    eng = Engine(kind, dtype, 0)
    for k, v in opts.items():
        eng.set_option(k, v)
    return eng


def _synthetic_obs(kind, theta, T, rng):
    x = theta[0]
    obs = np.zeros(T + 1)
    for t in range(1, T + 1):
        x = theta[0] + theta[1] * (x - theta[0]) + theta[2] * rng.standard_normal()
        obs[t] = x + theta[3] * rng.standard_normal() if kind == "lgss" else np.exp(0.5 * x) * rng.standard_normal()
    return obs


def _compare(eng, kind, theta, obs, N, L, draws, resampling="systematic", gen_init=True, x0=0.0,
             exact_particles=True):
    out = eng.run(obs, theta, N, L, True, gen_init, x0, resampling, draws=draws, store_history=True)
    hist = eng.fetch_history()
    steps = eng.fetch_steps()
    ref = orc.smoother(gu.MODEL_IDS[kind], theta, obs, N, L, draws, initial_state_value=x0,
                       generate_initial_state=gen_init, resampling=resampling)
    bad = np.argwhere(hist["ancestors"][0] != ref["ancestors"])
    assert bad.size == 0, "first ancestor mismatch at (t, n) = %s of %d" % (bad[0], len(bad))
    if exact_particles:
        np.testing.assert_array_equal(hist["particles"][0], ref["particles"])
    else:
        np.testing.assert_allclose(hist["particles"][0], ref["particles"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(steps["ll_terms"][0], ref["log_like_terms"], rtol=RTOL, atol=1e-11)
    assert abs(out["log_like"][0] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
    np.testing.assert_allclose(out["filt"][0], ref["filt_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(out["smo"][0], ref["smo_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(out["grad_terms"][0], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
    assert int(out["traj_index"][0]) == ref["trajectory_index"]
    if exact_particles:
        np.testing.assert_array_equal(out["traj"][0], ref["state_trajectory"].reshape(-1))
    else:
        np.testing.assert_allclose(out["traj"][0], ref["state_trajectory"].reshape(-1), rtol=1e-12, atol=1e-13)
    return out, ref


@pytest.mark.parametrize("name", gu.golden_names())
def test_cuda_matches_reference_golden(name):
    """The CUDA path replays the reference's own random stream and must reproduce its outputs."""
    g = gu.load(name)
    eng = _engine(g["model_kind"])
    draws = gu.draws_for(g)
    out = eng.run(g["obs"], g["theta"], g["no_particles"], g["fixed_lag"], True,
                  g["generate_initial_state"], g["initial_state"], g["resampling_method"],
                  draws=draws, store_history=True)
    hist = eng.fetch_history()
    assert gu.sha(hist["ancestors"][0].astype(np.int32)) == str(g["ancestors_sha256"])
    if g["model_kind"] == "lgss":
        assert gu.sha(hist["particles"][0]) == str(g["particles_sha256"])
    assert abs(out["log_like"][0] - g["log_like"]) <= RTOL * abs(g["log_like"])
    np.testing.assert_allclose(out["filt"][0], g["filt_state_est"], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(out["smo"][0], g["smo_state_est"], rtol=RTOL, atol=1e-12)
    if g["model_kind"] == "lgss":
        np.testing.assert_array_equal(out["traj"][0], g["state_trajectory"])
    else:
        np.testing.assert_allclose(out["traj"][0], g["state_trajectory"], rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("name", ["lgss_N256_T50_L5", "lgss_N1000_T500_L10", "lgss_est2_N200_T50_L5",
                                  "sv_N256_T100_L5", "svl_N256_T100_L5"])
def test_dropin_estimator_matches_reference_results(name):
    """`ParticleMethodsB200.smoother(model)` against the reference's `results` dict (score, Hessian
    with the prior folded in -- Q7/Q8 -- and the selected sub-blocks)."""
    from qnmh_sysid2018_b200 import ParticleMethodsB200
    from qnmh_sysid2018_b200.models import MODEL_CLASSES
    g = gu.load(name)
    model = MODEL_CLASSES[g["model_kind"]]()
    for k, v in zip(g["param_names"], g["theta"]):
        model.params[k] = float(v)
    model.set_observations(g["obs"])
    model.fix_true_params()
    model.create_inference_model(tuple(g["params_to_estimate"]))
    np.testing.assert_allclose(list(model.log_prior_gradient().values()), g["log_prior_gradient"], rtol=1e-13)
    np.testing.assert_allclose(list(model.log_prior_hessian().values()), g["log_prior_hessian"], rtol=1e-13)
    pf = ParticleMethodsB200({"no_particles": g["no_particles"], "fixed_lag": g["fixed_lag"],
                              "generate_initial_state": g["generate_initial_state"],
                              "initial_state": g["initial_state"], "estimate_gradient": True,
                              "estimate_hessian": True, "rng": "injected", "store_history": True})
    pf.draws = gu.draws_for(g)
    pf.smoother(model)
    r = pf.results
    assert abs(r["log_like"] - g["log_like"]) <= RTOL * abs(g["log_like"])
    np.testing.assert_allclose(r["gradient_internal"], g["gradient_internal"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["hessian_internal"], g["hessian_internal"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(r["log_joint_gradient_estimate"], g["log_joint_gradient_estimate"], rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(r["log_joint_hessian_estimate"], g["log_joint_hessian_estimate"], rtol=1e-9, atol=1e-9)
    assert r["state_trajectory"].shape == (1, g["T"] + 1)
    assert r["filt_state_est"].shape == (g["T"] + 1, 1) and r["smo_state_est"].shape == (g["T"] + 1, 1)
    assert pf.particles.shape == (g["no_particles"], g["T"] + 1)


# (N, T, L, ctas_per_filter): single CTA, multi-chunk single CTA, several cooperating CTAs,
# ragged tails, N not a multiple of anything, N < block size, fixed lag > T
GEOMETRIES = [(1, 20, 3, 0), (2, 20, 3, 0), (37, 25, 4, 1), (1000, 40, 10, 1), (3584, 30, 5, 1),
              (3585, 30, 5, 1), (5001, 30, 6, 1), (9000, 25, 5, 1), (5001, 30, 6, 3), (20000, 30, 10, 0),
              (65536, 20, 10, 0), (70001, 15, 4, 9), (300, 12, 30, 2), (100000, 12, 3, 5),
              (100000, 10, 3, 1), (70001, 12, 4, 2)]  # the last two stream > 8 sub-chunks per CTA


@pytest.mark.parametrize("N,T,L,G", GEOMETRIES)
def test_lgss_geometries_match_oracle(N, T, L, G):
    rng = np.random.default_rng(N * 1000 + T)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    obs = _synthetic_obs("lgss", theta, T, rng)
    draws = orc.random_draws(rng, N, T)
    eng = _engine("lgss", ctas_per_filter=G) if G else _engine("lgss")
    _compare(eng, "lgss", theta, obs, N, L, draws)


@pytest.mark.parametrize("kind,theta", [("sv", [0.5, 0.95, 0.25]), ("sv_leverage", [0.5, 0.95, 0.25, -0.3])])
@pytest.mark.parametrize("N,G", [(777, 1), (6000, 1), (12000, 4)])
def test_sv_models_match_oracle(kind, theta, N, G):
    rng = np.random.default_rng(N + len(theta))
    theta = np.array(theta)
    obs = _synthetic_obs(kind, theta, 40, rng)
    draws = orc.random_draws(rng, N, 40)
    eng = _engine(kind, ctas_per_filter=G)
    _compare(eng, kind, theta, obs, N, 7, draws, exact_particles=(kind == "sv"))


def test_stratified_resampling_matches_oracle():
    rng = np.random.default_rng(11)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    obs = _synthetic_obs("lgss", theta, 30, rng)
    for N, G in ((500, 1), (8000, 3)):
        draws = orc.random_draws(rng, N, 30, resampling="stratified")
        _compare(_engine("lgss", ctas_per_filter=G), "lgss", theta, obs, N, 5, draws, resampling="stratified")


def test_multinomial_resampling_matches_oracle():
    """resampling.pyx:33-42: unsorted per-particle uniforms, ancestors in any CTA's tile (pull-style
    search of the global CDF): one CTA, several CTAs, streaming tiles, ring and history storage."""
    rng = np.random.default_rng(12)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    obs = _synthetic_obs("lgss", theta, 25, rng)
    for N, G in ((1, 1), (500, 1), (8000, 3), (70001, 2)):
        draws = orc.random_draws(rng, N, 25, resampling="multinomial")
        _compare(_engine("lgss", ctas_per_filter=G), "lgss", theta, obs, N, 5, draws, resampling="multinomial")


def test_multinomial_philox_path_replayed_through_oracle():
    rng = np.random.default_rng(29)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    N, T, L, seed, ev = 5003, 20, 4, 77, 3
    obs = _synthetic_obs("lgss", theta, T, rng)
    eng = _engine("lgss")
    z = np.zeros((T + 1, N))
    u = np.zeros((T + 1, N))
    for t in range(1, T + 1):
        z[t] = eng.philox_normals(seed, ev, t, N)
        u[t] = eng.philox_uniforms(seed, ev, t, N)
    draws = orc.Draws(eng.philox_normals(seed, ev, 0, N), u, z, eng.philox_uniforms(seed, ev, -1)[0])
    out = eng.run(obs, theta, N, L, True, True, 0.0, "multinomial", seed=seed, eval_ids=[ev], store_history=True)
    hist = eng.fetch_history()
    ref = orc.smoother(orc.MODEL_LGSS, theta, obs, N, L, draws, resampling="multinomial")
    np.testing.assert_array_equal(hist["ancestors"][0], ref["ancestors"])
    assert abs(out["log_like"][0] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
    ring = eng.run(obs, theta, N, L, True, True, 0.0, "multinomial", seed=seed, eval_ids=[ev])
    np.testing.assert_array_equal(ring["smo"], out["smo"])


@pytest.mark.parametrize("kind,theta,resampling", [("sv", [0.5, 0.95, 0.25], "multinomial"),
                                                   ("sv_leverage", [0.5, 0.95, 0.25, -0.3], "stratified"),
                                                   ("sv_leverage", [0.5, 0.95, 0.25, -0.3], "multinomial")])
def test_sv_models_with_other_resamplers_match_oracle(kind, theta, resampling):
    """The stochastic-volatility models through the general kernel flavour's other resamplers."""
    rng = np.random.default_rng(len(resampling) + len(theta))
    theta = np.array(theta)
    obs = _synthetic_obs(kind, theta, 30, rng)
    for N, G in ((900, 1), (7000, 3)):
        draws = orc.random_draws(rng, N, 30, resampling=resampling)
        _compare(_engine(kind, ctas_per_filter=G), kind, theta, obs, N, 6, draws, resampling=resampling,
                 exact_particles=(kind == "sv"))


def test_fixed_initial_state_and_degenerate_weights():
    """Edge cases: fixed initial state (all particles equal -> CDF ties are uniform steps), and an
    outlier observation that puts almost all weight on a handful of particles."""
    rng = np.random.default_rng(5)
    theta = np.array([0.2, 0.5, 1.0, 0.05])
    obs = _synthetic_obs("lgss", theta, 25, rng)
    obs[7] += 9.0
    obs[15] -= 12.0
    for N, G in ((4096, 1), (30000, 8)):
        draws = orc.random_draws(rng, N, 25)
        _compare(_engine("lgss", ctas_per_filter=G), "lgss", theta, obs, N, 4, draws, gen_init=False, x0=0.3)


def test_batched_filters_match_oracle_each():
    """B independent evaluations side by side (config C3 shape, reduced): per-filter theta, shared
    observations, per-filter injected randomness; ring-buffer mode (no history)."""
    rng = np.random.default_rng(99)
    B, N, T, L = 9, 2048, 30, 5
    theta0 = np.array([0.2, 0.5, 1.0, 0.5])
    obs = _synthetic_obs("lgss", theta0, T, rng)
    thetas = theta0 + 0.05 * rng.standard_normal((B, 4))
    draws = [orc.random_draws(rng, N, T) for _ in range(B)]
    eng = _engine("lgss", max_groups=4)  # fewer groups than filters: groups loop over filters
    out = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", draws=draws)
    for b in range(B):
        ref = orc.smoother(orc.MODEL_LGSS, thetas[b], obs, N, L, draws[b])
        assert abs(out["log_like"][b] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
        np.testing.assert_allclose(out["smo"][b], ref["smo_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(out["grad_terms"][b], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
        assert int(out["traj_index"][b]) == ref["trajectory_index"]


def test_batched_filters_with_several_ctas_each():
    """Batch + groups of several CTAs: a group runs several filters one after the other, so the
    exchange counters and the smoother-job counters keep counting across evaluations of one launch."""
    rng = np.random.default_rng(101)
    B, N, T, L = 5, 9000, 25, 6
    theta0 = np.array([0.2, 0.5, 1.0, 0.5])
    obs = _synthetic_obs("lgss", theta0, T, rng)
    thetas = theta0 + 0.05 * rng.standard_normal((B, 4))
    draws = [orc.random_draws(rng, N, T) for _ in range(B)]
    eng = _engine("lgss", ctas_per_filter=3, max_groups=2)
    out = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", draws=draws)
    assert eng.info("ctas_per_filter") == 3 and eng.info("n_groups") == 2
    for b in range(B):
        ref = orc.smoother(orc.MODEL_LGSS, thetas[b], obs, N, L, draws[b])
        assert abs(out["log_like"][b] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
        np.testing.assert_allclose(out["smo"][b], ref["smo_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(out["grad_terms"][b], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
        assert int(out["traj_index"][b]) == ref["trajectory_index"]
    # production flavour (Philox): the batch equals the same evaluations run one at a time
    evals = [11, 12, 13, 14, 15]
    batch = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", seed=5, eval_ids=evals)
    single = _engine("lgss", ctas_per_filter=3)
    for b in range(B):
        one = single.run(obs, thetas[b], N, L, True, True, 0.0, "systematic", seed=5, eval_ids=[evals[b]])
        for k in ("log_like", "filt", "smo", "grad_terms", "traj_index"):
            np.testing.assert_array_equal(batch[k][b], one[k][0])


def test_ring_mode_equals_history_mode():
    """The (L+2)-deep rings must give the same numbers as keeping all T+1 generations."""
    rng = np.random.default_rng(17)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    N, T, L = 50000, 40, 10
    obs = _synthetic_obs("lgss", theta, T, rng)
    a = _engine("lgss").run(obs, theta, N, L, True, True, 0.0, "systematic", seed=7, store_history=True)
    b = _engine("lgss").run(obs, theta, N, L, True, True, 0.0, "systematic", seed=7)
    for k in ("log_like", "filt", "smo", "grad_terms", "traj_index"):
        np.testing.assert_array_equal(a[k], b[k])


def test_philox_path_replayed_through_oracle():
    """Philox mode: dump the normals/uniforms the device uses and replay them through the oracle."""
    rng = np.random.default_rng(23)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    N, T, L, seed, ev = 6001, 30, 6, 1234, 5
    obs = _synthetic_obs("lgss", theta, T, rng)
    eng = _engine("lgss")
    z = np.zeros((T + 1, N))
    u = np.zeros(T + 1)
    for t in range(1, T + 1):
        z[t] = eng.philox_normals(seed, ev, t, N)
        u[t] = eng.philox_uniforms(seed, ev, t)[0]
    draws = orc.Draws(eng.philox_normals(seed, ev, 0, N), u, z, eng.philox_uniforms(seed, ev, -1)[0])
    assert abs(z[1:].mean()) < 0.02 and abs(z[1:].std() - 1.0) < 0.02
    out = eng.run(obs, theta, N, L, True, True, 0.0, "systematic", seed=seed, eval_ids=[ev], store_history=True)
    hist = eng.fetch_history()
    ref = orc.smoother(orc.MODEL_LGSS, theta, obs, N, L, draws)
    np.testing.assert_array_equal(hist["ancestors"][0], ref["ancestors"])
    assert abs(out["log_like"][0] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
    np.testing.assert_allclose(out["grad_terms"][0], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("N,G", [(3000, 1), (20000, 5)])
def test_fully_adapted_filter_matches_oracle_and_kalman(N, G):
    """The fully adapted (optimal proposal) LGSS filter -- not in the reference (Q13) -- against its
    NumPy restatement on injected randomness (ancestors / particles bit exact) and against the
    exact Kalman log-likelihood (its estimator has far lower variance than the bootstrap filter)."""
    from oracle import kalman_oracle as ko
    rng = np.random.default_rng(N)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    T, L = 40, 5
    obs = _synthetic_obs("lgss", theta, T, rng)
    draws = orc.random_draws(rng, N, T)
    eng = _engine("lgss_fully_adapted", ctas_per_filter=G)
    out = eng.run(obs, theta, N, L, True, True, 0.0, "systematic", draws=draws, store_history=True)
    hist = eng.fetch_history()
    ref = orc.fully_adapted_filter(theta, obs, N, draws)
    ref.update(orc.fixed_lag_smoother(orc.MODEL_LGSS, theta, obs, ref, L))
    np.testing.assert_array_equal(hist["ancestors"][0], ref["ancestors"])
    np.testing.assert_allclose(hist["particles"][0], ref["particles"], rtol=1e-13, atol=1e-14)
    assert abs(out["log_like"][0] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
    np.testing.assert_allclose(out["filt"][0], ref["filt_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(out["smo"][0], ref["smo_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(out["grad_terms"][0], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
    m0, p0 = ko.stationary_prior(theta)
    kf = ko.kalman_filter(obs, theta, initial_state=m0, initial_cov=p0)
    big = _engine("lgss_fully_adapted").run(obs, theta, 1 << 16, L, True, True, 0.0, "systematic", seed=5)
    assert abs(big["log_like"][0] - kf["log_like"]) < 0.05
    assert np.abs(big["filt"][0][1:] - kf["filt_state_est"][1:]).max() < 0.02


def test_device_normals_match_numpy_box_muller():
    """The device's Philox4x32-10 stream and its custom fp64 log / sqrt / sincospi against a NumPy
    restatement: same bits in, normals equal to a few ulp; moments and tails as expected."""
    import philox_ref
    eng = _engine("lgss")
    for seed, ev, t in ((1234, 5, 1), (87655678, 0, 0), (2 ** 40 + 17, 2 ** 33 + 1, 999)):
        n = 200001
        z = eng.philox_normals(seed, ev, t, n)
        ref = philox_ref.normals(seed, ev, t, n)
        np.testing.assert_allclose(z, ref, rtol=0, atol=5e-15)
        assert abs(eng.philox_uniforms(seed, ev, max(t, 1))[0] - philox_ref.uniform(seed, ev, 1, max(t, 1))) == 0.0
    z = eng.philox_normals(7, 7, 3, 2000000)
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1) < 2e-3
    assert abs(np.mean(z ** 4) - 3.0) < 0.03 and np.abs(z).max() > 4.5


def test_filtered_means_and_loglik_against_exact_kalman():
    """At N = 2^18 the PF's filtered means and log-likelihood must sit on the exact Kalman values
    (stationary initial law on both sides): |filt - KF| is a few Monte Carlo standard errors."""
    from oracle import kalman_oracle as ko
    rng = np.random.default_rng(8)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    obs = _synthetic_obs("lgss", theta, 200, rng)
    m0, p0 = ko.stationary_prior(theta)
    # the PF draws x_0 from the stationary law and then propagates: start the KF one step earlier
    kf = ko.kalman_filter(obs, theta, initial_state=m0, initial_cov=p0)
    eng = _engine("lgss")
    out = eng.run(obs, theta, 1 << 18, 10, True, True, 0.0, "systematic", seed=21)
    err = np.abs(out["filt"][0][1:] - kf["filt_state_est"][1:])
    assert err.max() < 6.0 * np.sqrt(kf["filt_state_cov"][1:].max() / (1 << 18)) * 3, err.max()
    assert abs(out["log_like"][0] - kf["log_like"]) < 0.15, (out["log_like"][0], kf["log_like"])
    for dtype in ("float32",):
        o32 = _engine("lgss", dtype).run(obs, theta, 1 << 18, 10, True, True, 0.0, "systematic", seed=21)
        assert abs(o32["log_like"][0] - kf["log_like"]) < 0.3  # documented fp32 bound: ~1e-3 relative
        assert np.abs(o32["filt"][0][1:] - kf["filt_state_est"][1:]).max() < 0.02
        # fp32 smoother (fp32 rings and normals, fp64 accumulation): same estimates up to Monte Carlo
        # error of two independent N = 2^18 runs
        assert np.abs(o32["smo"][0] - out["smo"][0]).max() < 0.08
        s64, s32 = out["grad_terms"][0].sum(axis=-1), o32["grad_terms"][0].sum(axis=-1)
        assert np.all(np.isfinite(s32)) and np.abs(s32 - s64).max() < 1.5, (s32, s64)


def test_score_and_smoothed_means_against_kalman_rts():
    """Statistical check of the fixed-lag smoother against the exact answers: the Kalman RTS
    smoother's means and its exact score terms (oracle/kalman_oracle.py, pinned to the reference's
    kalman_methods/standard.py:86-178).  The reference stores the term of the pair (x_{i-1}, x_i) at index i, the
    particle smoother the pair (x_i, x_{i+1}) at index i; the first / last terms (the reference's smo[0] = 0
    convention, the tail of the fixed lag) are left out.  Tolerances: Monte Carlo error at N = 2^18 plus the
    fixed-lag bias (phi^L), calibrated with the NumPy oracle at N = 40 000 (0.07 per term, 0.06 on the sums)."""
    from oracle import kalman_oracle as ko
    theta, T, N, L = np.array([0.2, 0.5, 1.0, 0.5]), 120, 1 << 18, 10
    obs = _synthetic_obs("lgss", theta, T, np.random.default_rng(11))
    m0, p0 = ko.stationary_prior(theta)
    ks = ko.kalman_smoother(obs, theta, initial_state=m0, initial_cov=p0)
    out = _engine("lgss").run(obs, theta, N, L, True, True, 0.0, "systematic", seed=5)
    pf = out["grad_terms"][0][:3, 2:T - 1]
    kf = ks["gradient_part"][:3, 3:T]
    assert np.abs(pf - kf).max() < 0.08, np.abs(pf - kf).max(axis=1)
    assert np.abs(pf.sum(axis=1) - kf.sum(axis=1)).max() < 0.15, (pf.sum(axis=1), kf.sum(axis=1))
    assert np.abs(out["smo"][0][2:T - 1] - ks["smo_state_est"][2:T - 1]).max() < 0.03


def test_philox_loglik_is_consistent_with_kalman():
    """Statistical check without injected noise: the PF log-likelihood on the reference
    data at N=2^16 is within Monte Carlo error of the reference Kalman log-likelihood."""
    with np.load(gu.GOLDEN_DIR + "/kalman_lgss.npz") as k:
        obs, theta, ll_kf = k["obs_T500_mid"], k["theta_T500_mid"], float(k["ll_T500_mid"])
    eng = _engine("lgss")
    lls = []
    for ev in range(4):
        out = eng.run(obs, theta, 1 << 16, 10, True, True, 0.0, "systematic", seed=3, eval_ids=[ev])
        lls.append(out["log_like"][0])
    # the reference KF starts from x0 = 0 with cov 1e-5, the PF from the stationary law: the
    # difference is bounded by the first-step term; PF noise at N = 65536 is ~0.05
    assert abs(np.mean(lls) - ll_kf) < 1.0, (lls, ll_kf)
    assert np.std(lls) < 0.2
