"""Step-locked ("teacher-forced") comparison of a device history with the oracle.

A free-running comparison of two particle filters stops being bitwise meaningful at the first
CDF tie: one ancestor that differs changes a particle, hence every later weight and CDF.  At
N = 2^20 the device's blocked fp64 scan and the reference's sequential `np.cumsum` differ by
rounding (~1e-13), so a position (n+u)/N sits between the two CDFs about once every few time steps.
Here every time step is checked on its own instead: generation t-1 AS THE DEVICE STORED IT goes
through the oracle's resample -> propagate -> weight step (oracle/pf_oracle.py, i.e.
standard.py:69-101 + resampling.pyx) and is compared with the device's generation t:

  * ancestors equal, except where the position lies within the rounding bound of the CDF of a
    step between the two answers (a tie; counted, each one verified);
  * particles bit exact (LGSS, SV) or to 1e-12 (SV-leverage: device exp vs glibc) wherever the
    ancestor agrees; log-weights to 1e-13;
  * log-likelihood terms, filtered means recomputed from the device's log-weights.

The fixed-lag smoother is then checked on the device's own genealogy: `fixed_lag_smoother` of the
oracle runs on the device history and must give the device's smoothed means and score terms."""
import numpy as np

from oracle import pf_oracle as orc


def normalised_weights(logw):
    m = logw.max(axis=1, keepdims=True)
    s = np.exp(logw - m)
    return s / s.sum(axis=1, keepdims=True), m.ravel(), s.sum(axis=1)


def check_history(model_id, theta, obs, hist, steps, out, draws, L, resampling="systematic",
                  exact_particles=True, obs_index="numpy", gen_init=True, x0=0.0,
                  lw_rtol=1e-13, stat_rtol=1e-10, grad_rtol=1e-9, max_ties=64, fully_adapted=False):
    """hist/steps/out: Engine.fetch_history()/fetch_steps()/fetch() of ONE filter (index 0 taken by the
    caller).  Returns the number of verified tie mismatches."""
    x, lw, anc = hist["particles"], hist["log_weights"], hist["ancestors"]
    T1, N = x.shape
    theta = np.asarray(theta, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    w, m, S = normalised_weights(lw)
    tie_tol = 4.0 * N * np.finfo(np.float64).eps
    resample = orc.RESAMPLERS[resampling]

    x_init = orc.initial_state(model_id, theta, draws.z0) if gen_init else np.full(N, float(x0))
    np.testing.assert_array_equal(x[0], x_init)

    ties = 0
    for t in range(1, T1):
        u = draws.u[t]
        a_ref = resample(w[t - 1], u)
        bad = np.flatnonzero(anc[t] != a_ref)
        if bad.size:
            cdf = np.cumsum(w[t - 1])
            if resampling == "multinomial":
                pos = u[bad]
            else:
                pos = ((u[bad] if np.ndim(u) else u) + bad) / N
            for n, p, ad, ar in zip(bad, pos, anc[t][bad], a_ref[bad]):
                lo, hi = (ad, ar) if ad < ar else (ar, ad)
                gap = np.abs(cdf[lo:hi] - p).max()
                assert gap <= tie_tol, ("ancestor mismatch at (t=%d, n=%d): device %d, oracle %d, CDF is %.3e away "
                                        "from the position (tie bound %.1e)" % (t, n, ad, ar, gap, tie_tol))
            ties += bad.size
        same = anc[t] == a_ref
        x_ref = orc.generate_state(model_id, theta, obs, x[t - 1][anc[t]], t, draws.z[t], obs_index)
        if exact_particles:
            np.testing.assert_array_equal(x[t][same], x_ref[same], err_msg="particles at t=%d" % t)
        else:
            np.testing.assert_allclose(x[t][same], x_ref[same], rtol=1e-12, atol=1e-13)
        if not fully_adapted:
            lw_ref = orc.evaluate_obs(model_id, theta, obs, x[t], t)
            np.testing.assert_allclose(lw[t], lw_ref, rtol=lw_rtol, atol=1e-13)
    assert ties <= max_ties, "%d tie mismatches" % ties

    if not fully_adapted:
        ll_ref = m + np.log(S) - np.log(N)
        ll_ref[0] = 0.0
        np.testing.assert_allclose(steps["ll_terms"], ll_ref, rtol=stat_rtol, atol=1e-11)
        assert abs(out["log_like"] - ll_ref.sum()) <= stat_rtol * abs(ll_ref.sum())
        filt_ref = (w * x).sum(axis=1)
        filt_ref[0] = 0.0
        np.testing.assert_allclose(out["filt"], filt_ref, rtol=stat_rtol, atol=1e-12)
        if L is not None:
            fo = dict(particles=x, weights=w, ancestors=anc, filt_state_est=filt_ref.reshape(-1, 1))
            sm = orc.fixed_lag_smoother(model_id, theta, obs, fo, L)
            np.testing.assert_allclose(out["smo"], sm["smo_state_est"].reshape(-1), rtol=stat_rtol, atol=1e-12)
            np.testing.assert_allclose(out["grad_terms"], sm["smo_gradient_est"], rtol=grad_rtol, atol=1e-10)
        # trajectory draw (standard.py:104-106) on the device's final weights
        cdf = np.cumsum(w[-1])
        cdf = cdf / cdf[-1]
        k = int(np.searchsorted(cdf, draws.u_final, side="right"))
        if min(abs(cdf[max(k - 1, 0)] - draws.u_final), abs(cdf[min(k, N - 1)] - draws.u_final)) > tie_tol:
            assert int(out["traj_index"]) == int(anc[-1][min(k, N - 1)])
    return ties


def philox_draws(eng, seed, eval_id, N, T, resampling="systematic"):
    """The numbers the device's Philox path uses, dumped through the C ABI (pfb200_philox_*)."""
    per_particle = resampling != "systematic"
    z = np.zeros((T + 1, N))
    u = np.zeros((T + 1, N)) if per_particle else np.zeros(T + 1)
    for t in range(1, T + 1):
        z[t] = eng.philox_normals(seed, eval_id, t, N)
        u[t] = eng.philox_uniforms(seed, eval_id, t, N) if per_particle else eng.philox_uniforms(seed, eval_id, t)[0]
    return orc.Draws(eng.philox_normals(seed, eval_id, 0, N), u, z, eng.philox_uniforms(seed, eval_id, -1)[0])
