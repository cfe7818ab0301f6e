"""`ParticleMethodsB200`: drop-in for the reference's state estimators
`state.particle_methods.standard.ParticleMethods` (standard.py:25-189) and its Cython twins
(`cython_lgss.py:24`, `cython_sv.py:24`, `cython_sv_leverage.py:24`).

Same surface -- `__init__(new_settings)`, `.name`, `.settings`, `.results`, `.filter(model)`,
`.smoother(model)` -- so `MetropolisHastings.run(state_estimator)`
(parameter/mcmc/metropolis_hastings.py:227,326-329,413-416) can use it unchanged.  The particle
work (resample / propagate / weight / fixed-lag score terms) runs in the sm_100a kernels behind
libpfb200.so; what stays here is the P x P host arithmetic of the reference: score = sum_t G[:,t],
the Hessian estimate built from G (standard.py:172-178 or cython_lgss.py:62-70) and the prior
folding of `_estimate_gradient_and_hessian` (base_state_inference.py:38-72).

Semantics follow the NumPy twin (SURVEY.md section 8 a-Q): lag = min(i+L, T) (Q4), the sigma_e
score term is computed (Q6), `results` is an instance dict (Q12).  Switches for the documented
divergences: settings['compat'] = 'numpy' | 'cython' selects the Hessian formula (Q7);
settings['svl_obs_index'] = 'numpy' | 'model' the observation the leverage term reads (Q10).

`results['state_trajectory']` is the reference's row of the particle matrix, `particles[a_T[k], :]`
(standard.py:104-106, Q9).  It needs every generation of the particles: settings['store_states'] =
'auto' (default) keeps them on the device when (T+1) * N * 8 bytes fit `STATE_HISTORY_BUDGET`
(2 GiB; every configuration of the reference's scripts is a few MB), True always, False never; when
they are not kept the field is NaN and a warning says so -- never a different estimator.

Random streams: a run draws from the Philox stream (settings['seed'] + settings['seed_offset'],
eval_id) where `eval_id` counts the evaluations of THIS estimator instance (0, 1, 2, ...).  Two
instances with the same seed replay the same noise; give replicates / parallel chains distinct
`seed_offset`s, as the reference's scripts do with `np.random.seed(87655678 + seed_offset)`
(helper_linear_gaussian.py:33).
"""
import warnings
from collections import OrderedDict

import numpy as np

STATE_HISTORY_BUDGET = 2 << 30  # bytes of particle history 'store_states': 'auto' accepts

from .models import model_kind_of


def segal_weinstein_numpy(G, no_obs_plus_one):
    """standard.py:174-178 as written: part1 = part2 = G G^T, so H = (1 - 1/(T+1)) G G^T (Q7)."""
    part1 = np.dot(G, G.transpose())
    part2 = np.matmul(G, G.transpose())
    return part1 - part2 / no_obs_plus_one


def segal_weinstein_cython(G, no_obs):
    """cython_lgss.py:62-70: G G^T - s s^T / model.no_obs with s = sum_t G[:, t]."""
    s = np.sum(G, axis=1)
    return np.dot(G, G.transpose()) - np.outer(s, s) / no_obs


def fold_in_log_prior(results, model, settings):
    """base_state_inference.py:38-72, including its in-place updates (Q8): `gradient_internal`
    is taken before the prior gradient is added to `log_joint_gradient_estimate`; every Hessian
    entry (i, j) loses h[p_i] + h[p_j]; `hessian_internal` is indexed with
    `model.params_to_estimate_idx` as the reference builds it."""
    grad = results['log_joint_gradient_estimate']
    hess = results['log_joint_hessian_estimate']
    gradient = OrderedDict()
    internal = []
    for i, param in enumerate(model.params.keys()):
        if param in model.params_to_estimate:
            gradient[param] = grad[i]
            internal.append(grad[i])
    internal = np.array(internal)
    if settings['estimate_gradient']:
        prior_grad = model.log_prior_gradient()
        prior_hess = model.log_prior_hessian()
        keys = list(prior_grad.keys())
        for i, p1 in enumerate(keys):
            grad[i] += prior_grad[p1]
            for j, p2 in enumerate(keys):
                hess[i, j] -= prior_hess[p1]
                hess[i, j] -= prior_hess[p2]
        idx = np.asarray(model.params_to_estimate_idx, dtype=int)
        results['gradient_internal'] = internal
        results['gradient'] = gradient
        results['hessian_internal'] = np.array(hess[np.ix_(idx, idx)])


class ParticleMethodsB200(object):
    """Bootstrap particle filter and fixed-lag particle smoother on a B200."""

    def __init__(self, new_settings=None, engine=None):
        self.name = "Particle methods (B200 CUDA implementation)"
        self.settings = {'no_particles': 100,
                         'resampling_method': 'systematic',
                         'fixed_lag': 0,
                         'initial_state': 0.0,
                         'generate_initial_state': False,
                         'estimate_gradient': False,
                         'estimate_hessian': False,
                         'verbose': False,
                         # keys beyond the reference's
                         'dtype': 'float64',      # 'float64' | 'float32' storage/arithmetic
                         'device': -1,            # CUDA device ordinal, -1 = current
                         'rng': 'philox',         # 'philox' | 'injected' (self.draws)
                         'seed': 87655678,        # helper_linear_gaussian.py:33
                         'seed_offset': 0,        # added to the seed (replicates / parallel chains)
                         'store_history': False,  # keep (T+1, N) particles/weights/ancestors
                         'store_states': 'auto',  # keep (T+1, N) particles only (state_trajectory rows)
                         'compat': 'numpy',       # Hessian formula: 'numpy' | 'cython'
                         'svl_obs_index': 'numpy',  # leverage term reads obs[t] | 'model': obs[t-1]
                         # 'bootstrap' (the reference's filter) | 'fully_adapted' (optimal proposal,
                         # LGSS only; not in the reference -- north star / SURVEY Q13)
                         'filter_type': 'bootstrap'
                         }
        if new_settings:
            self.settings.update(new_settings)
        self.results = {}
        self.draws = None          # injected randomness for the next call (rng='injected')
        self.eval_id = 0           # Philox stream id; advanced on every evaluation
        self._engine = engine
        self._engine_key = None
        self._cfg_key = None
        self.last_device_ms = None

    def __repr__(self):
        return self.name

    # ------------------------------------------------------------------------------------------
    def _get_engine(self, kind):
        key = (kind, self.settings['dtype'], self.settings['device'])
        if self._engine is None or (self._engine_key is not None and self._engine_key != key):
            from .engine import Engine  # raises if libpfb200.so or the GPU is missing
            self._engine = Engine(kind, self.settings['dtype'], self.settings['device'])
            self._engine_key = key
            self._cfg_key = None
        return self._engine

    def _evaluate(self, model, do_smoother):
        s = self.settings
        kind = model_kind_of(model)
        if s['filter_type'] == 'fully_adapted':
            if kind != 'lgss':
                raise ValueError("the fully adapted particle filter is only available for the linear Gaussian model")
            kind = 'lgss_fully_adapted'
        elif s['filter_type'] != 'bootstrap':
            raise ValueError("filter_type must be 'bootstrap' or 'fully_adapted'")
        if s['resampling_method'] not in ('systematic', 'stratified', 'multinomial'):
            raise ValueError("Unknown resampling method selected...")
        engine = self._get_engine(kind)
        obs = np.asarray(model.obs, dtype=np.float64).reshape(-1)
        theta = np.asarray(model.get_all_params(), dtype=np.float64)
        if s['verbose']:
            print("")
            print("Particle filter running with model parameters:")
            print(["%.3f" % v for v in theta])
        injected = s['rng'] == 'injected'
        if injected and self.draws is None:
            raise ValueError("settings['rng']='injected' needs self.draws")
        states = s['store_states']
        if states == 'auto':
            states = len(obs) * int(s['no_particles']) * 8 <= STATE_HISTORY_BUDGET
        states = bool(states) and not s['store_history']
        self._have_rows = bool(s['store_history']) or states
        cfg_key = (kind, obs.tobytes(), s['no_particles'], s['fixed_lag'], do_smoother,
                   s['generate_initial_state'], float(s['initial_state']), s['resampling_method'],
                   s['store_history'], states, s['svl_obs_index'], injected)
        eval_id = self.eval_id
        self.eval_id += 1
        seed = int(s['seed']) + int(s['seed_offset'])
        if injected or cfg_key != self._cfg_key:
            engine.configure(obs, theta, s['no_particles'], s['fixed_lag'], do_smoother,
                             s['generate_initial_state'], s['initial_state'], s['resampling_method'],
                             seed=seed, eval_ids=[eval_id], draws=self.draws if injected else None,
                             store_history=s['store_history'],
                             svl_obs_model=(s['svl_obs_index'] == 'model'), store_states=states)
            self._cfg_key = None if injected else cfg_key
        else:  # same data and geometry: only the parameters and the random stream change
            engine.update_params(theta, seed=seed, eval_ids=[eval_id])
        engine.execute()
        out = engine.fetch()
        try:
            self.last_device_ms = engine.last_execute_ms()
        except Exception:
            self.last_device_ms = None
        return engine, out

    def _store_filter_results(self, model, engine, out, b=0):
        T1 = model.no_obs + 1
        filt = out['filt'][b].reshape(T1, 1).copy()
        if self.settings['store_history'] and engine is not None:
            hist = engine.fetch_history()
            # reference layout is particle-major (N, T+1) (standard.py:54-57)
            self.particles = hist['particles'][0].T.copy()
            lw = hist['log_weights'][0]
            w = np.exp(lw - lw.max(axis=1, keepdims=True))
            self.weights = (w / w.sum(axis=1, keepdims=True)).T.copy()
            self.ancestors = hist['ancestors'][0].T.astype(np.float64)
        if self._have_rows:
            traj = out['traj'][b].reshape(1, T1).copy()  # particles[a_T[k], :] (Q9)
        else:
            # without all generations of the particles the reference's row-of-the-storage-matrix
            # trajectory (Q9) does not exist: NaN, never a different estimator in its place
            traj = np.full((1, T1), np.nan)
            if not getattr(self, '_warned_rows', False):
                warnings.warn("state_trajectory is NaN: the particle history ((T+1) x N x 8 bytes) exceeds "
                              "STATE_HISTORY_BUDGET; set settings['store_states']=True to keep it anyway")
                self._warned_rows = True
        self.results.update({'filt_state_est': filt,
                             'log_like': float(out['log_like'][b]),
                             'state_trajectory': traj,
                             'trajectory_index': int(out['traj_index'][b])})
        if self.settings['verbose']:
            print("Log-likelihood estimate is: " + str(self.results['log_like']))

    # ------------------------------------------------------------------------------------------
    def filter(self, model):
        """Bootstrap particle filter (standard.py:42-119)."""
        self.name = "Bootstrap particle filter (B200)"
        engine, out = self._evaluate(model, do_smoother=False)
        self._store_filter_results(model, engine, out)

    def smoother(self, model):
        """Fixed-lag particle smoother (standard.py:121-189)."""
        self.name = "Bootstrap particle filter and fixed-lag particle smoother (B200)."
        no_obs = model.no_obs + 1
        no_params = model.no_params
        if self.settings['fixed_lag'] < 1 and self.settings['estimate_gradient']:
            # the reference only trips over fixed_lag=0 inside its gradient branch (standard.py:160-163);
            # without gradients it smooths with zero hops (smo == filt), and so does the device
            raise NameError("name 'next_ancestor' is not defined (fixed_lag must be >= 1 when "
                            "estimate_gradient is set, standard.py:161)")
        engine, out = self._evaluate(model, do_smoother=True)
        self._store_filter_results(model, engine, out)
        self._store_smoother_results(model, out)

    def _store_smoother_results(self, model, out, b=0):
        """Host side of `smoother` for row b of a (possibly batched) device result."""
        no_obs = model.no_obs + 1
        no_params = model.no_params
        G = out['grad_terms'][b]  # (P, T+1), the reference's smo_gradient_est
        grad_est = np.zeros(no_params)
        hess_est = np.zeros((no_params, no_params))
        if self.settings['estimate_gradient']:
            grad_est = np.sum(G, axis=1)
            if self.settings['compat'] == 'cython':
                hess_est = segal_weinstein_cython(G, model.no_obs)
            else:
                hess_est = segal_weinstein_numpy(G, no_obs)
            if not np.all(np.isfinite(hess_est)):
                print("Numerical problems in Segal-Weinstein estimator, returning identity.")
                hess_est = np.eye(no_params)
        self.results.update({'smo_state_est': out['smo'][b].reshape(no_obs, 1).copy(),
                             'smo_gradient_est': G.copy(),
                             'log_joint_gradient_estimate': grad_est,
                             'log_joint_hessian_estimate': hess_est})
        if self.settings['estimate_gradient']:
            fold_in_log_prior(self.results, model, self.settings)
