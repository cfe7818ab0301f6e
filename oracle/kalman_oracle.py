"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Kalman filter for the linear Gaussian model: the exact
log-likelihood and filtered means the particle filter estimates.

Restates `KalmanMethods.filter` (python/state/kalman_methods/standard.py:37-84): scalar recursion
with the prediction x_{t|t-1} = mu + phi (x_{t-1|t-1} - mu), P_{t|t-1} = phi^2 P_{t-1|t-1} + sigma_v^2,
gain K = P/(P + sigma_e^2) and the log-likelihood as the sum of Gaussian predictive log-densities.
Pinned to the reference's own outputs in tests/golden/kalman_lgss.npz (tests/test_oracle_golden.py).
`initial_cov` = stationary variance makes it the exact counterpart of a particle filter that draws
its initial state from the stationary law (`generate_initial_state=True`)."""
import numpy as np


def kalman_filter(obs, theta, initial_state=0.0, initial_cov=1e-5):
    mu, phi, sigma_v, sigma_e = (float(v) for v in theta)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    T1 = len(obs)
    sv2, se2 = sigma_v ** 2, sigma_e ** 2
    filt = np.zeros(T1)
    cov = np.zeros(T1)
    filt[0], cov[0] = initial_state, initial_cov
    ll = 0.0
    for t in range(1, T1):
        pred = mu + phi * (filt[t - 1] - mu)
        pcov = phi * cov[t - 1] * phi + sv2
        s = pcov + se2
        k = pcov / s
        filt[t] = pred + k * (obs[t] - pred)
        cov[t] = pcov - k * pcov
        z = (obs[t] - pred) / np.sqrt(s)
        ll += -z ** 2 / 2.0 - np.log(np.sqrt(2 * np.pi)) - np.log(np.sqrt(s))
    return dict(log_like=float(ll), filt_state_est=filt, filt_state_cov=cov)


def stationary_prior(theta):
    mu, phi, sigma_v = float(theta[0]), float(theta[1]), float(theta[2])
    return mu, sigma_v ** 2 / (1.0 - phi ** 2)


def kalman_smoother(obs, theta, initial_state=0.0, initial_cov=1e-5, estimate_gradient=True):
    """RTS smoother + exact score terms + Segal-Weinstein Hessian estimate: restates
    `KalmanMethods.smoother` (python/state/kalman_methods/standard.py:86-178).

    Returns dict(log_like, filt_*, pred_*, smo_state_est (T+1,), smo_state_cov (T+1,),
    gradient_part (4, T), log_joint_gradient_estimate (4,), log_joint_hessian_estimate (4, 4)).
    Index conventions are the reference's: the backward pass runs i = T-1 .. 1 (smo[0] stays 0,
    standard.py:119-126), the two-step covariance is seeded at T-1 (:130-139), the score terms run
    i = 1 .. T-1 with (x_{i-1}, x_i) pairs (:143-162), the sigma_e entry is 0 (:162), and the Hessian
    is G G^T - s s^T / T (:166-170).  Pinned to tests/golden/kalman_rts_lgss.npz."""
    mu, phi, sigma_v, sigma_e = (float(v) for v in theta)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    T = len(obs) - 1
    sv2 = sigma_v ** 2
    kf = kalman_filter_full(obs, theta, initial_state, initial_cov)
    pred, pcov, filt, fcov, gain = (kf[k] for k in ("pred_state_est", "pred_state_cov", "filt_state_est",
                                                    "filt_state_cov", "kalman_gain"))
    smo_gain = np.zeros(T + 1)
    twostep = np.zeros(T + 1)
    smo = np.zeros(T + 1)
    scov = np.zeros(T + 1)
    smo[-1], scov[-1] = filt[-1], fcov[-1]
    for i in range(T - 1, 0, -1):
        smo_gain[i] = fcov[i] * phi
        smo_gain[i] /= pcov[i + 1]
        smo[i] = filt[i] + smo_gain[i] * (smo[i + 1] - pred[i + 1])
        scov[i] = fcov[i]
        scov[i] += smo_gain[i] ** 2 * (scov[i + 1] - pcov[i + 1])
    out = dict(kf, smo_state_est=smo, smo_state_cov=scov)
    if not estimate_gradient:
        return out
    two = (1 - gain[-1]) * phi
    two *= fcov[-1]
    twostep[T - 1] = two
    for i in range(T - 1, 0, -1):
        term1 = fcov[i] * smo_gain[i - 1]
        term2 = smo_gain[i - 1] ** 2
        twostep[i] = term1 + term2 * (twostep[i + 1] - phi * fcov[i])
    G = np.zeros((4, T))
    for i in range(1, T):
        nxt, cur = smo[i], smo[i - 1]
        eta = nxt * nxt + scov[i]
        eta1 = cur ** 2 + scov[i - 1]
        psi = cur * nxt + twostep[i]
        quad = nxt - mu - phi * (cur - mu)
        isv2 = 1.0 / sv2
        G[0, i] = isv2 * quad * (1.0 - phi)
        t1 = isv2 * (1.0 - phi ** 2)
        t2 = psi - phi * eta1
        t2 -= cur * mu * (1.0 - 2.0 * phi)
        t2 += -nxt * mu + mu ** 2 * (1.0 - phi)
        G[1, i] = t1 * t2
        t1 = eta - 2 * phi * psi + phi ** 2 * eta1
        t2 = -2.0 * (nxt - phi * smo[i - 1])
        t2 *= (1.0 - phi) * mu
        t3 = mu ** 2 * (1.0 - phi) ** 2
        G[2, i] = isv2 * (t1 + t2 + t3) - 1.0
        G[3, i] = 0.0
    score = np.sum(G, axis=1)
    hess = np.dot(G, G.T) - np.outer(score, score) / T
    out.update(gradient_part=G, log_joint_gradient_estimate=score, log_joint_hessian_estimate=hess)
    return out


def kalman_filter_full(obs, theta, initial_state=0.0, initial_cov=1e-5):
    """`kalman_filter` with the intermediate arrays the smoother needs (standard.py:45-83)."""
    mu, phi, sigma_v, sigma_e = (float(v) for v in theta)
    obs = np.asarray(obs, dtype=np.float64).reshape(-1)
    T1 = len(obs)
    sv2, se2 = sigma_v ** 2, sigma_e ** 2
    pred, pcov, filt, fcov, gain = (np.zeros(T1) for _ in range(5))
    filt[0], fcov[0] = initial_state, initial_cov
    ll = 0.0
    for i in range(1, T1):
        pred[i] = mu
        pred[i] += phi * (filt[i - 1] - mu)
        pcov[i] = phi * fcov[i - 1] * phi
        pcov[i] += sv2
        s = pcov[i] + se2
        gain[i] = pcov[i] / s
        innov = (obs[i] - pred[i])
        innov *= gain[i]
        filt[i] = pred[i] + innov
        fcov[i] = pcov[i] - gain[i] * pcov[i]
        z = (obs[i] - pred[i]) / np.sqrt(s)
        ll += -z ** 2 / 2.0 - np.log(np.sqrt(2 * np.pi)) - np.log(np.sqrt(s))
    return dict(log_like=float(ll), pred_state_est=pred, pred_state_cov=pcov, filt_state_est=filt,
                filt_state_cov=fcov, kalman_gain=gain)
