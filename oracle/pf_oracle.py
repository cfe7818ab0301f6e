"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A NumPy restatement of the sequential Monte Carlo state estimator of compops/qnmh-sysid2018
(`python/state/particle_methods/standard.py`, `resampling.pyx`, the per-particle model callbacks
in `python/models/*.py` and `python/state/base_state_inference.py`).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may import
it, and only as the checker / the CPU baseline.  The CUDA product path never calls into here.

Parity status: PINNED.  `tests/golden/make_golden.py` runs the unmodified reference (imported from
/root/reference in the build container) on seeded inputs and stores its outputs in
`tests/golden/*.npz`; `tests/test_oracle_golden.py` checks this restatement against every one of
them (log-likelihood to 1e-12, ancestors / state trajectory exactly, score / Hessian / smoothed
means to 1e-12).

Differences to the reference, none of which change results:
  * storage is time-major `[t][n]` (the reference is particle-major `(N, T+1)`);
  * the O(N T^2) re-indexing of `ancestors_resamp` (standard.py:80) is skipped unless
    `faithful_ancestry=True` -- only its last column is ever read (standard.py:105) and that column
    equals the last ancestor vector;
  * all random numbers are INJECTED (`Draws`), in the order the reference consumes them from
    `np.random` (SURVEY.md section 8 a-R): `draws_from_legacy_seed` reproduces that stream.
  * systematic/stratified resampling clamp the ancestor index to N-1 where the reference would
    raise IndexError (resampling.pyx:78-83 has no guard when cumsum[-1] < position).
Third-party arithmetic restated: `scipy.stats.norm.logpdf(x, loc, scale)` ==
`-((x-loc)/scale)**2/2 - log(sqrt(2 pi)) - log(scale)` (scipy 1.18 `_norm_logpdf`; checked against
scipy itself in tests/test_oracle_golden.py).
"""
from collections import OrderedDict

import numpy as np

MODEL_LGSS = 0  # models/linear_gaussian_model.py
MODEL_SV = 1  # models/stochastic_volatility_model.py
MODEL_SVL = 2  # models/stochastic_volatility_model_leverage.py
MODEL_NAMES = {MODEL_LGSS: "lgss", MODEL_SV: "sv", MODEL_SVL: "sv_leverage"}
PARAM_NAMES = {MODEL_LGSS: ("mu", "phi", "sigma_v", "sigma_e"),
               MODEL_SV: ("mu", "phi", "sigma_v"),
               MODEL_SVL: ("mu", "phi", "sigma_v", "rho")}
NORM_LOGC = float(np.log(np.sqrt(2 * np.pi)))  # scipy _norm_pdf_logC


class Draws(object):
    """Injected randomness for one filter evaluation.

    z0      (N,)       standard normals of the initial state (unused if the state is fixed)
    u       (T+1,) or (T+1, N)  resampling uniforms; row t is consumed at time step t (row 0 unused)
    z       (T+1, N)   propagation normals; row t is consumed at time step t (row 0 unused)
    u_final scalar     uniform consumed by the final trajectory draw (np.random.choice)
    """

    def __init__(self, z0, u, z, u_final):
        self.z0 = np.ascontiguousarray(z0, dtype=np.float64)
        self.u = np.ascontiguousarray(u, dtype=np.float64)
        self.z = np.ascontiguousarray(z, dtype=np.float64)
        self.u_final = float(u_final)


def draws_from_legacy_seed(seed, no_particles, no_obs, generate_initial_state=True,
                           resampling="systematic"):
    """The exact numbers `ParticleMethods.filter` draws after `np.random.seed(seed)`.

    Order (SURVEY.md 8 a-R): normal(size=(1,N)) [linear_gaussian_model.py:90] if the initial state
    is generated; per step one uniform() [resampling.pyx:70] (N uniforms for multinomial /
    stratified, resampling.pyx:38,49) then randn(1,N) [linear_gaussian_model.py:106]; finally the
    random_sample() inside np.random.choice [standard.py:104].  `no_obs` is model.no_obs (= T).
    """
    rs = np.random.RandomState(seed)
    N, T = int(no_particles), int(no_obs)
    z0 = rs.normal(size=(1, N))[0] if generate_initial_state else np.zeros(N)
    per_particle_u = resampling in ("multinomial", "stratified")
    u = np.zeros((T + 1, N)) if per_particle_u else np.zeros(T + 1)
    z = np.zeros((T + 1, N))
    for t in range(1, T + 1):
        u[t] = rs.uniform(size=N) if per_particle_u else rs.uniform()
        z[t] = rs.randn(1, N)[0]
    u_final = rs.random_sample()
    return Draws(z0, u, z, u_final)


def random_draws(rng, no_particles, no_obs, resampling="systematic"):
    """Fresh draws from a numpy Generator (for seeded parity tests that do not need the legacy stream)."""
    N, T = int(no_particles), int(no_obs)
    per_particle_u = resampling in ("multinomial", "stratified")
    u = rng.random((T + 1, N)) if per_particle_u else rng.random(T + 1)
    return Draws(rng.standard_normal(N), u, rng.standard_normal((T + 1, N)), rng.random())


# ----------------------------------------------------------------------------------------------
# resampling (resampling.pyx)
# ----------------------------------------------------------------------------------------------
def systematic(weights, u):
    """resampling.pyx:65-85.  The two-pointer walk returns, for every position, the first j with
    position < cumsum[j], i.e. searchsorted(cumsum, positions, side='right')."""
    N = len(weights)
    positions = (np.arange(N) + u) / N
    cumulative_sum = np.cumsum(weights)
    idx = np.searchsorted(cumulative_sum, positions, side="right")
    return np.minimum(idx, N - 1).astype(np.int32)


def stratified(weights, u):
    """resampling.pyx:44-63 (u is a vector of N uniforms)."""
    N = len(weights)
    positions = (u + np.arange(N)) / N
    cumulative_sum = np.cumsum(weights)
    idx = np.searchsorted(cumulative_sum, positions, side="right")
    return np.minimum(idx, N - 1).astype(np.int32)


def multinomial(weights, u):
    """resampling.pyx:33-42 (searchsorted default side='left', last cumsum forced to 1)."""
    cumulative_sum = np.cumsum(weights)
    cumulative_sum[-1] = 1.0
    return np.searchsorted(cumulative_sum, u).astype(np.int32)


def systematic_two_pointer(weights, u):
    """Literal transcription of the while loop (resampling.pyx:76-83); small N only.  Used by the
    tests to show that the searchsorted form above is the same function."""
    N = len(weights)
    positions = (np.arange(N) + u) / N
    cumulative_sum = np.cumsum(weights)
    out = np.zeros(N, dtype=np.int32)
    i = j = 0
    while i < N:
        if j >= N:  # the reference would raise IndexError here
            out[i:] = N - 1
            break
        if positions[i] < cumulative_sum[j]:
            out[i] = j
            i += 1
        else:
            j += 1
    return out


RESAMPLERS = {"systematic": systematic, "stratified": stratified, "multinomial": multinomial}


# ----------------------------------------------------------------------------------------------
# per-particle model arithmetic (models/*.py); operation order follows the reference line by line
# ----------------------------------------------------------------------------------------------
def _norm_logpdf(x, loc, scale):
    z = (x - loc) / scale
    return (-z ** 2 / 2.0 - NORM_LOGC) - np.log(scale)


def initial_state(model_id, theta, z0):
    """generate_initial_state: linear_gaussian_model.py:87-90, stochastic_volatility_model.py:84-87,
    stochastic_volatility_model_leverage.py:87-90 (identical in the three models)."""
    mu, phi, sigma_v = theta[0], theta[1], theta[2]
    noise_stdev = sigma_v
    noise_stdev = noise_stdev / np.sqrt(1.0 - phi ** 2)
    return mu + noise_stdev * z0


def generate_state(model_id, theta, obs, cur_state, time_step, z, obs_index="numpy"):
    """generate_state: linear_gaussian_model.py:103-107, stochastic_volatility_model.py:100-104,
    stochastic_volatility_model_leverage.py:103-108."""
    mu, phi, sigma_v = theta[0], theta[1], theta[2]
    mean = mu + phi * (cur_state - mu)
    if model_id == MODEL_SVL:
        rho = theta[3]
        # Q10: the NumPy reference reads obs[time_step]; the model-consistent index is time_step-1
        y = obs[time_step] if obs_index == "numpy" else obs[time_step - 1]
        mean = mean + sigma_v * rho * np.exp(-0.5 * cur_state) * y
        stdev = np.sqrt(1.0 - rho ** 2) * sigma_v
        return mean + stdev * z
    return mean + sigma_v * z


def evaluate_obs(model_id, theta, obs, cur_state, time_step):
    """evaluate_obs: linear_gaussian_model.py:153-154, stochastic_volatility_model.py:148-150,
    stochastic_volatility_model_leverage.py:154-156."""
    y = obs[time_step]
    if model_id == MODEL_LGSS:
        return _norm_logpdf(y, cur_state, theta[3])
    volatility = np.exp(0.5 * cur_state)
    return _norm_logpdf(y, 0.0, volatility)


def log_joint_gradient(model_id, theta, obs, next_state, cur_state, time_index):
    """log_joint_gradient: linear_gaussian_model.py:223-244, stochastic_volatility_model.py:214-230,
    stochastic_volatility_model_leverage.py:227-259.  Returns a (P, N) array in parameter order."""
    mu, phi, sigma_v = theta[0], theta[1], theta[2]
    y = obs[time_index]
    q = next_state - mu
    q = q - phi * (cur_state - mu)
    if model_id == MODEL_SVL:
        rho = theta[3]
        q = q - sigma_v * rho * np.exp(-0.5 * cur_state) * y
        rho_term = 1.0 - rho ** 2
        q_matrix = sigma_v ** (-2) * rho_term ** (-1)
    else:
        q_matrix = sigma_v ** (-2)
    g_mu = q_matrix * q * (1.0 - phi)
    g_phi = q_matrix * q
    g_phi = g_phi * (cur_state - mu)
    g_phi = g_phi * (1.0 - phi ** 2)
    g_sigmav = q_matrix * q ** 2 - 1.0
    if model_id == MODEL_LGSS:
        r_matrix = theta[3] ** (-2)
        obs_quad = y - cur_state
        g_sigmae = r_matrix * obs_quad ** 2 - 1.0
        return np.stack([g_mu, g_phi, g_sigmav, g_sigmae])
    if model_id == MODEL_SV:
        return np.stack([g_mu, g_phi, g_sigmav])
    g_sigmav = g_sigmav + q * rho * sigma_v * y * np.exp(-0.5 * cur_state) * q_matrix * sigma_v
    g_rho = rho_term ** (-1) * rho
    g_rho = g_rho + q * sigma_v * y * np.exp(-0.5 * cur_state) * q_matrix
    g_rho = g_rho - (-rho * q ** 2 * q_matrix / rho_term)
    g_rho = g_rho * (1.0 - rho ** 2)
    return np.stack([g_mu, g_phi, g_sigmav, g_rho])


# ----------------------------------------------------------------------------------------------
# bootstrap particle filter (standard.py:42-119)
# ----------------------------------------------------------------------------------------------
def particle_filter(model_id, theta, obs, no_particles, draws, initial_state_value=0.0,
                    generate_initial_state=True, resampling="systematic", obs_index="numpy",
                    faithful_ancestry=False):
    """Returns dict(particles, weights, ancestors [all (T+1, N)], log_like, log_like_terms,
    filt_state_est (T+1,1), state_trajectory (1,T+1), trajectory_index)."""
    theta = np.asarray(theta, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    T1 = obs.shape[0]  # = model.no_obs + 1
    N = int(no_particles)
    resample = RESAMPLERS[resampling]

    x = np.zeros((T1, N))
    w = np.zeros((T1, N))
    anc = np.zeros((T1, N), dtype=np.int32)
    filt = np.zeros((T1, 1))
    ll = np.zeros(T1)
    anc_resamp = np.zeros((N, T1)) if faithful_ancestry else None

    if generate_initial_state:
        x[0] = initial_state(model_id, theta, draws.z0)
    else:
        x[0] = initial_state_value
    w[0] = 1.0 / N

    for t in range(1, T1):
        a = resample(w[t - 1], draws.u[t])
        if faithful_ancestry:  # standard.py:80-81, O(N t) per step
            anc_resamp[:, 0:(t - 1)] = anc_resamp[a, 0:(t - 1)]
            anc_resamp[:, t] = a
        anc[t] = a
        x[t] = generate_state(model_id, theta, obs, x[t - 1, a], t, draws.z[t], obs_index)
        logw = evaluate_obs(model_id, theta, obs, x[t], t)
        max_weight = np.max(logw)
        shifted = np.exp(logw - max_weight)
        norm_factor = np.sum(shifted)
        w[t] = shifted / norm_factor
        ll[t] = max_weight
        ll[t] += np.log(norm_factor)
        ll[t] -= np.log(N)
        filt[t] = np.sum(w[t] * x[t])

    # trajectory draw: np.random.choice(N, 1, p=w_T) == searchsorted(cdf/cdf[-1], u, 'right')
    cdf = np.cumsum(w[T1 - 1])
    cdf = cdf / cdf[-1]
    k = int(np.searchsorted(cdf, draws.u_final, side="right"))
    idx = int(anc[T1 - 1, k])  # ancestors_resamp[k, -1] == last ancestor vector (standard.py:105)
    traj = x[:, idx].reshape(1, T1).copy()  # a ROW of the storage matrix (Q9)

    return dict(particles=x, weights=w, ancestors=anc, log_like=float(np.sum(ll)),
                log_like_terms=ll, filt_state_est=filt, state_trajectory=traj,
                trajectory_index=idx)


# ----------------------------------------------------------------------------------------------
# fixed-lag smoother (standard.py:121-189)
# ----------------------------------------------------------------------------------------------
def fixed_lag_smoother(model_id, theta, obs, filter_out, fixed_lag, estimate_gradient=True):
    """Returns dict(smo_state_est (T+1,1), smo_gradient_est (P,T+1) [= G],
    log_joint_gradient_estimate (P,), log_joint_hessian_estimate (P,P))."""
    theta = np.asarray(theta, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    x, w, anc = filter_out["particles"], filter_out["weights"], filter_out["ancestors"]
    T1, N = x.shape
    P = len(PARAM_NAMES[model_id])
    L = int(fixed_lag)

    smo = np.zeros((T1, 1))
    G = np.zeros((P, T1))
    smo[0] = filter_out["filt_state_est"][0]
    smo[-1] = filter_out["filt_state_est"][-1]

    for i in range(0, T1 - 1):
        lag = int(min(i + L, T1 - 1))
        cur = np.arange(N)
        nxt = None
        for j in range(lag, i, -1):
            nxt = cur
            cur = anc[j, cur]
        smo[i] = np.nansum(x[i, cur] * w[lag])
        if estimate_gradient:
            if nxt is None:
                raise NameError("next_ancestor is undefined for fixed_lag=0 (standard.py:161)")
            sub = log_joint_gradient(model_id, theta, obs, x[i + 1, nxt], x[i, cur], i)
            for p in range(P):
                G[p, i] = np.nansum(sub[p] * w[lag])

    score = np.zeros(P)
    hess = np.zeros((P, P))
    if estimate_gradient:
        score = np.sum(G, axis=1)
        part1 = np.dot(G, G.transpose())
        part2 = np.matmul(G, G.transpose())
        hess = part1 - part2 / T1  # Q7: (1 - 1/(T+1)) G G^T, as written at standard.py:175-178
    return dict(smo_state_est=smo, smo_gradient_est=G, log_joint_gradient_estimate=score,
                log_joint_hessian_estimate=hess)


def hessian_cython_form(G, no_obs):
    """The Segal-Weinstein form of the Cython wrapper (cython_lgss.py:62-70):
    G G^T - s s^T / model.no_obs with s = sum_t G[:, t]."""
    s = np.sum(G, axis=1)
    return np.dot(G, G.transpose()) - np.outer(s, s) / no_obs


# ----------------------------------------------------------------------------------------------
# prior folding (base_state_inference.py:38-72)
# ----------------------------------------------------------------------------------------------
def estimate_gradient_and_hessian(score, hess, param_names, params_to_estimate,
                                  params_to_estimate_idx, log_prior_gradient, log_prior_hessian):
    """Restates `_estimate_gradient_and_hessian`.  `log_prior_gradient/hessian` are ordered dicts
    over ALL model parameters.  Returns dict(gradient_internal, gradient, hessian_internal,
    log_joint_gradient_estimate, log_joint_hessian_estimate) with the in-place mutations of the
    reference applied (Q8)."""
    score = np.array(score, dtype=np.float64, copy=True)
    hess = np.array(hess, dtype=np.float64, copy=True)
    gradient = OrderedDict()
    gradient_internal = []
    for i, param in enumerate(param_names):
        if param in params_to_estimate:
            gradient[param] = float(score[i])
            gradient_internal.append(float(score[i]))
    for i, p1 in enumerate(log_prior_gradient.keys()):
        score[i] += log_prior_gradient[p1]
        for j, p2 in enumerate(log_prior_gradient.keys()):
            hess[i, j] -= log_prior_hessian[p1]
            hess[i, j] -= log_prior_hessian[p2]
    idx = np.asarray(params_to_estimate_idx, dtype=int)
    return dict(gradient_internal=np.array(gradient_internal), gradient=gradient,
                hessian_internal=np.array(hess[np.ix_(idx, idx)]),
                log_joint_gradient_estimate=score, log_joint_hessian_estimate=hess)


# ----------------------------------------------------------------------------------------------
# fully adapted auxiliary particle filter for the LGSS model.  NOT in the reference (only
# the bootstrap filter exists, standard.py:43); the north star names it.  Parity status of this
# function: UNPINNED by reference vectors -- it is pinned instead to the exact Kalman log-likelihood
# (tests) and shares resampling / smoother code with the pinned bootstrap path above.
#   resampling weights of generation t:  v_t[n] = p(y_{t+1} | x_t[n]) = N(y_{t+1}; mu+phi(x-mu), sv^2+se^2)
#   proposal:  x_{t+1} ~ p(x_{t+1} | x_t, y_{t+1}) = N(A pred + B y_{t+1}, s^2),
#              A = se^2/(sv^2+se^2), B = sv^2/(sv^2+se^2), s^2 = sv^2 se^2/(sv^2+se^2)
#   log_like = sum_t log( mean_n v_{t-1}[n] );  filtered mean = plain average (equal weights)
# ----------------------------------------------------------------------------------------------
def fully_adapted_filter(theta, obs, no_particles, draws, initial_state_value=0.0,
                         generate_initial_state=True):
    theta = np.asarray(theta, dtype=np.float64)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    mu, phi, sv, se = theta
    T1, N = obs.shape[0], int(no_particles)
    tot = sv * sv + se * se
    A, Bc, s_prop, s_pred = se * se / tot, sv * sv / tot, np.sqrt(sv * sv * se * se / tot), np.sqrt(tot)
    x = np.zeros((T1, N))
    w = np.zeros((T1, N))
    anc = np.zeros((T1, N), dtype=np.int32)
    filt = np.zeros((T1, 1))
    ll = np.zeros(T1)
    x[0] = initial_state(MODEL_LGSS, theta, draws.z0) if generate_initial_state else initial_state_value
    for t in range(0, T1):
        if t > 0:
            a = systematic(w[t - 1], draws.u[t])
            anc[t] = a
            pred = mu + phi * (x[t - 1, a] - mu)
            x[t] = (A * pred + Bc * obs[t]) + s_prop * draws.z[t]
            filt[t] = np.sum(x[t]) / N
        if t < T1 - 1:  # look-ahead weights
            logv = _norm_logpdf(obs[t + 1], mu + phi * (x[t] - mu), s_pred)
            m = np.max(logv)
            sh = np.exp(logv - m)
            S = np.sum(sh)
            w[t] = sh / S
            ll[t + 1] = m + np.log(S) - np.log(N)
        else:
            w[t] = 1.0 / N
    cdf = np.cumsum(w[T1 - 1])
    cdf = cdf / cdf[-1]
    k = int(np.searchsorted(cdf, draws.u_final, side="right"))
    idx = int(anc[T1 - 1, min(k, N - 1)])
    return dict(particles=x, weights=w, ancestors=anc, log_like=float(np.sum(ll)), log_like_terms=ll,
                filt_state_est=filt, state_trajectory=x[:, idx].reshape(1, T1).copy(), trajectory_index=idx)


def smoother(model_id, theta, obs, no_particles, fixed_lag, draws, **filter_kwargs):
    """filter + fixed-lag smoother in one call (what `ParticleMethods.smoother` does)."""
    estimate_gradient = filter_kwargs.pop("estimate_gradient", True)
    out = particle_filter(model_id, theta, obs, no_particles, draws, **filter_kwargs)
    out.update(fixed_lag_smoother(model_id, theta, obs, out, fixed_lag, estimate_gradient))
    return out
