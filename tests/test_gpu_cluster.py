"""GPU parity tests of the CLUSTER kernel (csrc/pf_cluster.cuh: one filter per thread-block cluster,
parents / ancestors in distributed shared memory, one all-gather per time step) against the CPU
oracle and against the persistent kernel, through the C ABI.

Same bars as test_gpu_parity.py: ancestors bit exact, LGSS particles bit exact, ll / means 1e-10,
score terms 1e-9.  The per-CTA weight shift of this kernel rescales tile totals by exp(m_g - m): an
ancestor can differ from the oracle only where a position is within a few ulp of a CDF step."""
import numpy as np
import pytest

import golden_util as gu
from oracle import pf_oracle as orc
from test_gpu_parity import RTOL, _compare, _engine, _synthetic_obs

pytestmark = pytest.mark.gpu

THETA = {"lgss": np.array([0.2, 0.5, 1.0, 0.5]), "sv": np.array([0.5, 0.95, 0.25]),
         "sv_leverage": np.array([0.5, 0.95, 0.25, -0.3])}

# (N, T, L, cluster): one CTA, every cluster size, ragged last CTA, empty trailing CTAs, N < 32,
# fixed lag > T, fixed lag 1, the largest filter the kernel takes (16 x 1024)
CASES = [(1, 20, 3, 1), (37, 25, 4, 1), (37, 25, 4, 2), (1000, 40, 10, 1), (1000, 40, 10, 2), (1000, 60, 10, 4),
         (1000, 60, 10, 8), (1000, 60, 10, 16), (2000, 40, 10, 16), (520, 30, 5, 16), (300, 12, 30, 4),
         (777, 30, 1, 8), (4096, 30, 10, 16), (4097, 25, 6, 8), (16384, 20, 10, 16), (9001, 20, 4, 16)]


@pytest.mark.parametrize("N,T,L,C", CASES)
def test_cluster_lgss_matches_oracle(N, T, L, C):
    rng = np.random.default_rng(N * 1000 + T + C)
    obs = _synthetic_obs("lgss", THETA["lgss"], T, rng)
    draws = orc.random_draws(rng, N, T)
    eng = _engine("lgss", cluster=C)
    _compare(eng, "lgss", THETA["lgss"], obs, N, L, draws)
    assert eng.info("cluster") == C and eng.info("ctas_per_filter") == C


@pytest.mark.parametrize("kind", ["sv", "sv_leverage"])
@pytest.mark.parametrize("N,C", [(777, 8), (1500, 16), (6000, 16)])
def test_cluster_sv_models_match_oracle(kind, N, C):
    rng = np.random.default_rng(N + C)
    T, L = 40, 10
    obs = _synthetic_obs(kind, THETA[kind], T, rng)
    draws = orc.random_draws(rng, N, T)
    eng = _engine(kind, cluster=C)
    _compare(eng, kind, THETA[kind], obs, N, L, draws, exact_particles=False)
    assert eng.info("cluster") == C


@pytest.mark.parametrize("name", ["lgss_N1000_T500_L10", "lgss_fixedinit_N128_T40_L3", "lgss_ragged_N77_T33_L40",
                                  "lgss_L1_N300_T60", "sv_N1500_T200_L10", "svl_N1500_T200_L10"])
def test_cluster_matches_reference_golden(name):
    """The reference's own golden vectors through the cluster kernel (automatic cluster size)."""
    g = gu.load(name)
    if g["resampling_method"] != "systematic":
        pytest.skip("the cluster kernel is systematic-only")
    eng = _engine(g["model_kind"])
    out = eng.run(g["obs"], g["theta"], g["no_particles"], g["fixed_lag"], True, g["generate_initial_state"],
                  g["initial_state"], g["resampling_method"], draws=gu.draws_for(g), store_history=True)
    assert eng.info("cluster") >= 1, "automatic geometry should pick the cluster kernel for a single small filter"
    hist = eng.fetch_history()
    assert gu.sha(hist["ancestors"][0].astype(np.int32)) == str(g["ancestors_sha256"])
    if g["model_kind"] == "lgss":
        assert gu.sha(hist["particles"][0]) == str(g["particles_sha256"])
    assert abs(out["log_like"][0] - g["log_like"]) <= RTOL * abs(g["log_like"])
    np.testing.assert_allclose(out["smo"][0], g["smo_state_est"], rtol=RTOL, atol=1e-12)


def test_cluster_equals_persistent_kernel_on_philox_noise():
    """Same Philox counters in both kernels: ancestors and particles identical, sums to rounding; ring mode
    (no history) on the cluster side, i.e. the production configuration of an MH iteration."""
    N, T, L = 2000, 300, 10
    obs = _synthetic_obs("lgss", THETA["lgss"], T, np.random.default_rng(3))
    a = _engine("lgss", cluster=-1)
    ra = a.run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=11, store_history=True)
    ha = a.fetch_history()
    assert a.info("cluster") == 0
    for C in (4, 16):
        b = _engine("lgss", cluster=C)
        rb = b.run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=11, store_history=True)
        hb = b.fetch_history()
        np.testing.assert_array_equal(ha["ancestors"][0], hb["ancestors"][0])
        np.testing.assert_array_equal(ha["particles"][0], hb["particles"][0])
        c = _engine("lgss", cluster=C)
        rc = c.run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=11)  # rings of L + 6 generations
        for r in (rb, rc):
            assert abs(r["log_like"][0] - ra["log_like"][0]) <= 1e-12 * abs(ra["log_like"][0])
            np.testing.assert_allclose(r["filt"][0], ra["filt"][0], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(r["smo"][0], ra["smo"][0], rtol=1e-11, atol=1e-13)
            np.testing.assert_allclose(r["grad_terms"][0], ra["grad_terms"][0], rtol=1e-9, atol=1e-10)
            assert int(r["traj_index"][0]) == int(ra["traj_index"][0])


def test_cluster_batch_runs_several_evaluations_per_cluster():
    """More filters than resident clusters (max_groups = 3): a cluster runs evaluations back to back, the
    barriers / progress counters / sequence numbers carry over; each filter equals its own oracle run."""
    B, N, T, L = 10, 600, 30, 5
    rng = np.random.default_rng(8)
    obs = _synthetic_obs("lgss", THETA["lgss"], T, rng)
    thetas = THETA["lgss"] + 0.03 * rng.standard_normal((B, 4))
    draws = [orc.random_draws(np.random.default_rng(100 + b), N, T) for b in range(B)]
    eng = _engine("lgss", cluster=4, max_groups=3)
    out = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", draws=draws)
    assert eng.info("cluster") == 4 and eng.info("n_groups") == 3
    for b in range(B):
        ref = orc.smoother(orc.MODEL_LGSS, thetas[b], obs, N, L, draws[b])
        assert abs(out["log_like"][b] - ref["log_like"]) <= RTOL * abs(ref["log_like"])
        np.testing.assert_allclose(out["smo"][b], ref["smo_state_est"].reshape(-1), rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(out["grad_terms"][b], ref["smo_gradient_est"], rtol=1e-9, atol=1e-10)
        assert int(out["traj_index"][b]) == ref["trajectory_index"]


def test_cluster_philox_batch_equals_single_runs():
    """Philox path, B filters on clusters of 8 with fewer groups than filters == B single-filter launches."""
    B, N, T, L = 7, 1000, 60, 10
    rng = np.random.default_rng(9)
    obs = _synthetic_obs("lgss", THETA["lgss"], T, rng)
    thetas = THETA["lgss"] + 0.03 * rng.standard_normal((B, 4))
    evals = np.arange(40, 40 + B)
    eng = _engine("lgss", cluster=8, max_groups=2)
    out = eng.run(obs, thetas, N, L, True, True, 0.0, "systematic", seed=5, eval_ids=evals)
    for b in range(B):
        one = _engine("lgss", cluster=8)
        r = one.run(obs, thetas[b], N, L, True, True, 0.0, "systematic", seed=5, eval_ids=evals[b:b + 1])
        assert r["log_like"][0] == out["log_like"][b]
        np.testing.assert_array_equal(r["grad_terms"][0], out["grad_terms"][b])
        np.testing.assert_array_equal(r["smo"][0], out["smo"][b])


def test_cluster_filter_only_and_fp32():
    N, T, L = 1000, 100, 10
    obs = _synthetic_obs("lgss", THETA["lgss"], T, np.random.default_rng(4))
    ref = _engine("lgss", cluster=-1).run(obs, THETA["lgss"], N, L, False, True, 0.0, "systematic", seed=2)
    out = _engine("lgss", cluster=8).run(obs, THETA["lgss"], N, L, False, True, 0.0, "systematic", seed=2)
    assert abs(out["log_like"][0] - ref["log_like"][0]) <= 1e-12 * abs(ref["log_like"][0])
    np.testing.assert_allclose(out["filt"][0], ref["filt"][0], rtol=1e-11, atol=1e-13)
    f64 = _engine("lgss", cluster=8).run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=2)
    f32 = _engine("lgss", "float32", cluster=8).run(obs, THETA["lgss"], N, L, True, True, 0.0, "systematic", seed=2)
    # fp32 storage: a different particle system after the first rounding; statistical agreement only
    assert abs(f32["log_like"][0] - f64["log_like"][0]) < 0.15 * np.sqrt(T / 100.0) + 1.0
    assert np.all(np.isfinite(f32["grad_terms"][0]))


def test_cluster_rejects_what_it_cannot_run():
    from qnmh_sysid2018_b200._lib import Pfb200Error
    obs = _synthetic_obs("lgss", THETA["lgss"], 10, np.random.default_rng(1))
    with pytest.raises(Pfb200Error):
        _engine("lgss", cluster=2).run(obs, THETA["lgss"], 5000, 3, True, True, 0.0, "systematic", seed=1)  # 2 x 1024 < N
    with pytest.raises(Pfb200Error):
        _engine("lgss", cluster=3)  # not a power of two
    # stratified resampling stays on the persistent kernel, silently in automatic mode
    eng = _engine("lgss")
    eng.run(obs, THETA["lgss"], 500, 3, True, True, 0.0, "stratified", seed=1)
    assert eng.info("cluster") == 0
