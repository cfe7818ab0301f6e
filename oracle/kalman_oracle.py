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
