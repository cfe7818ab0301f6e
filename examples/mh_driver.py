"""A compact particle Metropolis-Hastings driver (example harness, not part of the hot path).

The reference's driver (python/parameter/mcmc/metropolis_hastings.py + quasi_newton/) is the CALLER of
the state estimator and out of scope for this repository; it is used unmodified in
tests/test_host_logic.py.  This file is an independent, much smaller driver written from the
algorithm description of the paper (Dahlin, Wills & Ninness, "Constructing Metropolis-Hastings
proposals using damped BFGS updates"), so that complete chains can run where the reference is not
installed (the GPU box).  It talks to the estimator through the same surface the reference uses:
`estimator.smoother(model)` / `.filter(model)` and `results['log_like']`,
`results['gradient_internal']`.

  alg 'mh0'  random-walk proposal            theta' ~ N(theta, eps^2 Sigma)
  alg 'mh1'  first-order (MALA) proposal     theta' ~ N(theta + eps^2/2 Sigma g(theta), eps^2 Sigma)
  alg 'qmh'  quasi-Newton proposal: Sigma(theta) is the damped-BFGS estimate of the inverse negative
             Hessian built from the last `memory` (theta, gradient) pairs of the chain; falls back
             to the fixed Sigma while the memory is short or the estimate is not positive definite.

All moves are made in the model's free (unconstrained) parameterisation; the target is the
posterior times the Jacobian of the reparameterisation, as in the reference.
"""
import time

import numpy as np


def damped_bfgs_inverse_hessian(thetas, grads, init_scale=0.01):
    """Inverse negative Hessian estimate of log pi from (theta_k, grad_k) pairs with Powell damping
    (keeps every update positive definite).  Returns None if there is too little information."""
    if len(thetas) < 3:
        return None
    d = len(thetas[0])
    order = np.argsort([np.sum(t) for t in thetas])  # any consistent ordering of the memory works
    th = np.asarray(thetas)[order]
    gr = -np.asarray(grads)[order]  # gradients of the NEGATIVE log target
    B = np.eye(d) / init_scale  # Hessian (not inverse) estimate
    used = 0
    for k in range(1, len(th)):
        s = th[k] - th[k - 1]
        y = gr[k] - gr[k - 1]
        if np.dot(s, s) < 1e-16:
            continue
        Bs = B @ s
        sBs = float(s @ Bs)
        sy = float(s @ y)
        if sy < 0.2 * sBs:  # Powell's damping
            phi = 0.8 * sBs / (sBs - sy)
            y = phi * y + (1.0 - phi) * Bs
            sy = float(s @ y)
        if sy <= 1e-12 or sBs <= 1e-12:
            continue
        B = B - np.outer(Bs, Bs) / sBs + np.outer(y, y) / sy
        used += 1
    if used < 2:
        return None
    try:
        H = np.linalg.inv(B)
        H = 0.5 * (H + H.T)
        if np.all(np.linalg.eigvalsh(H) > 1e-10):
            return H
    except np.linalg.LinAlgError:
        pass
    return None


def _log_mvn(x, mean, cov):
    d = x - mean
    sign, logdet = np.linalg.slogdet(cov)
    return -0.5 * (len(x) * np.log(2 * np.pi) + logdet + d @ np.linalg.solve(cov, d))


class ParticleMH(object):
    def __init__(self, model, estimator, alg="qmh", step_size=0.5, base_cov=None, memory=20, seed=1):
        self.model, self.est, self.alg = model, estimator, alg
        self.eps = float(step_size)
        self.d = len(model.get_free_params())
        self.base_cov = np.eye(self.d) * 0.05 ** 2 if base_cov is None else np.asarray(base_cov)
        self.memory = int(memory)
        self.max_drift = 0.01  # qn_initial_hessian_scaling of the reference's examples
        self.rng = np.random.default_rng(seed)
        self.use_grad = alg in ("mh1", "qmh")
        self.est_seconds, self.est_calls, self.est_device_ms = 0.0, 0, 0.0  # time spent in the state estimator
        if self.use_grad:
            estimator.settings['estimate_gradient'] = True  # metropolis_hastings.py:244-245

    def _evaluate(self, free_params):
        m = self.model
        m.store_free_params(free_params)
        if not m.check_parameters():
            return None
        t0 = time.perf_counter()
        (self.est.smoother if self.use_grad else self.est.filter)(m)
        self.est_seconds += time.perf_counter() - t0
        self.est_calls += 1
        self.est_device_ms += getattr(self.est, 'last_device_ms', None) or 0.0
        res = self.est.results
        _, log_prior = m.log_prior()
        target = float(res['log_like']) + log_prior + m.log_jacobian()
        grad = None
        if self.use_grad:
            prior_grad = m.log_prior_gradient()
            grad = np.array(res['gradient_internal'], dtype=float)
            grad = grad + np.array([prior_grad[k] for k in m.params_to_estimate])
        if not np.isfinite(target) or (grad is not None and not np.all(np.isfinite(grad))):
            return None
        return dict(free=np.array(free_params, dtype=float), params=m.get_params().copy(),
                    target=target, log_like=float(res['log_like']), grad=grad)

    def _proposal(self, state, hist):
        cov = self.base_cov
        if self.alg == "qmh":
            H = damped_bfgs_inverse_hessian([h['free'] for h in hist], [h['grad'] for h in hist])
            # trust the quasi-Newton estimate only while it stays within a plausible scale (a proposal
            # standard deviation of 1 in free-parameter space); otherwise keep the fixed covariance
            if H is not None and np.max(np.linalg.eigvalsh(H)) <= 1.0:
                cov = H
        cov = self.eps ** 2 * cov
        if not np.all(np.isfinite(cov)) or np.linalg.cond(cov) > 1e10:
            cov = self.eps ** 2 * self.base_cov  # a degenerate quasi-Newton estimate: fixed covariance
        if self.use_grad:
            # 'scaled_gradient' safeguard (quasi_newton/main.py:108-110): the drift is never longer than
            # `max_drift`; a state-dependent proposal, evaluated the same way for the reverse move
            drift = 0.5 * cov @ state['grad']
            norm = float(np.linalg.norm(drift))
            if norm > self.max_drift:
                cov = cov * max(self.max_drift / norm, 1e-4)
                drift = 0.5 * cov @ state['grad']
            return state['free'] + drift, cov
        return state['free'], cov

    def run(self, no_iters, initial_free_params, burn_in=0, verbose=False):
        cur = self._evaluate(initial_free_params)
        if cur is None:
            raise ValueError("initial parameters are not valid")
        d = self.d
        chain = np.zeros((no_iters, d))
        free_chain = np.zeros((no_iters, d))
        ll = np.zeros(no_iters)
        accepted = np.zeros(no_iters, dtype=bool)
        hist = [cur]
        chain[0], free_chain[0], ll[0] = cur['params'], cur['free'], cur['log_like']
        t0 = time.perf_counter()
        for i in range(1, no_iters):
            mean_f, cov_f = self._proposal(cur, hist)
            prop_free = self.rng.multivariate_normal(mean_f, cov_f)
            prop = self._evaluate(prop_free)
            if prop is not None:
                mean_r, cov_r = self._proposal(prop, hist)
                log_a = (prop['target'] - cur['target'] + _log_mvn(cur['free'], mean_r, cov_r)
                         - _log_mvn(prop['free'], mean_f, cov_f))
                if np.log(self.rng.random()) < log_a:
                    cur = prop
                    accepted[i] = True
                    if self.use_grad:
                        hist.append(cur)
                        hist = hist[-self.memory:]
            chain[i], free_chain[i], ll[i] = cur['params'], cur['free'], cur['log_like']
            if verbose and i % verbose == 0:
                print("iter %6d  accept rate %.2f  params %s  ll %.2f" % (
                    i, accepted[:i + 1].mean(), np.round(chain[burn_in:i + 1].mean(axis=0), 3), ll[i]), flush=True)
        wall = time.perf_counter() - t0
        return dict(params=chain, free_params=free_chain, log_like=ll, accepted=accepted,
                    accept_rate=float(accepted[1:].mean()), seconds=wall,
                    time_per_iteration=wall / max(1, no_iters - 1),
                    estimator_calls=self.est_calls, estimator_seconds=self.est_seconds,
                    estimator_device_ms_per_call=self.est_device_ms / max(1, self.est_calls),
                    posterior_mean=chain[burn_in:].mean(axis=0), posterior_std=chain[burn_in:].std(axis=0))
