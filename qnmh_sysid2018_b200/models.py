"""Host-side mirror of the reference's model surface for the state-estimation hot path.

The particle filter kernels carry the per-particle arithmetic of the three reference models
(generate_initial_state / generate_state / evaluate_obs / log_joint_gradient); what stays on the
host is the small scalar surface the estimator and the Metropolis-Hastings driver read
(python/models/base_model.py:29-311, python/helpers/model_params.py:24-134,
python/helpers/inference_model.py:23-56, python/helpers/distributions/{normal,gamma}.py):
`params`, `obs`, `no_obs`, `no_params`, `get_all_params()`, `params_to_estimate(_idx)`,
`log_prior_gradient()`, `log_prior_hessian()`, the free-parameter transforms and `generate_data`.
The class names match the reference so that `ParticleMethodsB200` dispatches on either family.

The reparameterisation of every model is table driven: `identity`, `tanh` (phi, rho) or `log`
(sigma_v, sigma_e) -- linear_gaussian_model.py:246-273, stochastic_volatility_model.py:232-256,
stochastic_volatility_model_leverage.py:261-288.
"""
import copy
import math
from collections import OrderedDict

import numpy as np


# prior families: name -> (logpdf, d/dx logpdf, d2/dx2 logpdf)   (helpers/distributions)
def _normal(x, mean, stdev):
    return (-0.5 * math.log(2 * math.pi * stdev ** 2) - 0.5 / stdev ** 2 * (x - mean) ** 2,
            -(mean - x) / stdev ** 2, -1.0 / stdev ** 2)


def _gamma(x, shape, rate):
    return (shape * math.log(rate) - math.lgamma(shape) + (shape - 1.0) * math.log(x) - rate * x,
            (shape - 1.0) / x - rate, -(shape - 1.0) / (x ** 2))


PRIORS = {"normal": _normal, "gamma": _gamma}

_TO_FREE = {"identity": lambda v: v, "tanh": math.atanh, "log": math.log}
_FROM_FREE = {"identity": lambda v: v, "tanh": math.tanh, "log": math.exp}
# d theta / d free, and the second factor the reference uses in log_prior_hessian
_JAC = {"identity": lambda v: 1.0, "tanh": lambda v: 1.0 - v ** 2, "log": lambda v: v}
_JAC2 = {"identity": lambda v: 0.0, "tanh": lambda v: 1.0 - 2.0 * v ** 2, "log": lambda v: v}
_LOGJAC = {"identity": lambda v: 0.0, "tanh": lambda v: math.log(1.0 - v ** 2), "log": math.log}


class StateSpaceModel(object):
    model_kind = None  # 'lgss' | 'sv' | 'sv_leverage' -> device functor
    PARAM_NAMES = ()
    TRANSFORMS = {}
    DEFAULT_PARAMS = ()
    DEFAULT_PRIORS = {}

    def __init__(self):
        self.params = OrderedDict(zip(self.PARAM_NAMES, self.DEFAULT_PARAMS))
        self.free_params = OrderedDict((k, 0.0) for k in self.PARAM_NAMES)
        self.no_params = len(self.params)
        self.params_prior = dict(self.DEFAULT_PRIORS)
        self.initial_state = 0.0
        self.no_obs = 0
        self.states = None
        self.obs = None
        self.params_to_estimate = ()
        self.params_to_estimate_idx = np.zeros(0, dtype=int)
        self.no_params_to_estimate = 0
        self.true_params = None
        self.transform_params_to_free()

    # ---- parameters (helpers/model_params.py)
    def get_all_params(self):
        return np.array([self.params[k] for k in self.params])

    def get_params(self):
        return np.array([self.params[k] for k in self._estimated()])

    def get_free_params(self):
        return np.array([self.free_params[k] for k in self._estimated()])

    def _estimated(self):
        pte = self.params_to_estimate
        return (pte,) if isinstance(pte, str) else tuple(pte)

    def fix_true_params(self):
        self.true_params = copy.deepcopy(self.params)

    def store_params(self, new_params):
        self.params = copy.deepcopy(self.true_params)
        for k, v in zip(self._estimated(), np.atleast_1d(new_params)):
            self.params[k] = float(v)
        self.transform_params_to_free()

    def store_free_params(self, new_params):
        self.params = copy.deepcopy(self.true_params)
        self.transform_params_to_free()
        for k, v in zip(self._estimated(), np.atleast_1d(new_params)):
            self.free_params[k] = float(v)
        self.transform_params_from_free()

    def transform_params_to_free(self):
        for k in self.params:
            try:
                self.free_params[k] = _TO_FREE[self.TRANSFORMS[k]](self.params[k])
            except ValueError:
                self.free_params[k] = float("nan")

    def transform_params_from_free(self):
        for k in self.params:
            self.params[k] = _FROM_FREE[self.TRANSFORMS[k]](self.free_params[k])

    def create_inference_model(self, params_to_estimate):
        """helpers/inference_model.py:23-54 -- note that `params_to_estimate_idx` holds positions
        inside `params_to_estimate`, not inside `params` (reproduced as is)."""
        self.model_type = "Inference model"
        self.params_to_estimate = params_to_estimate
        self.no_params_to_estimate = 1 if isinstance(params_to_estimate, str) else len(params_to_estimate)
        idx = [int(params_to_estimate.index(k)) for k in self.params if k in params_to_estimate]
        self.params_to_estimate_idx = np.asarray(idx)

    def check_parameters(self):
        ok = True
        for k, kind in self.TRANSFORMS.items():
            v = self.params[k]
            if kind == "tanh" and abs(v) > 1.0:
                ok = False
            if kind == "log" and v < 0.0:
                ok = False
        return ok

    # ---- priors (base_model.py:183-235 + the per-model chain rule factors)
    def _prior_terms(self, k):
        fam, a, b = self.params_prior[k]
        return PRIORS[fam](self.params[k], a, b)

    def log_prior(self):
        prior = OrderedDict((k, self._prior_terms(k)[0]) for k in self.params)
        return prior, sum(prior[k] for k in self._estimated())

    def log_prior_gradient(self):
        out = OrderedDict()
        for k in self.params:
            out[k] = self._prior_terms(k)[1] * _JAC[self.TRANSFORMS[k]](self.params[k])
        return out

    def log_prior_hessian(self):
        out = OrderedDict()
        for k in self.params:
            _, grad, hess = self._prior_terms(k)
            kind, v = self.TRANSFORMS[k], self.params[k]
            out[k] = hess * _JAC[kind](v) + grad * _JAC2[kind](v) if kind != "identity" else hess
        return out

    def log_jacobian(self):
        try:
            return sum(_LOGJAC[self.TRANSFORMS[k]](self.params[k]) for k in self._estimated())
        except ValueError:
            return -np.inf

    # ---- data
    def set_observations(self, obs):
        obs = np.asarray(obs, dtype=np.float64).reshape(-1, 1)
        self.obs = obs
        self.no_obs = obs.shape[0] - 1

    def import_data(self, file_name):
        """CSV with a header line containing `observation` (and optionally `state`);
        helpers/data_handling.py:51-90.  Honour a preset `no_obs` like the reference."""
        with open(file_name) as fh:
            header = [s.strip() for s in fh.readline().split(",")]
            rows = np.loadtxt(fh, delimiter=",", ndmin=2)
        if "observation" not in header:
            raise ValueError("No observations in file, header must be observation.")
        obs = rows[:, header.index("observation")]
        if self.no_obs:
            obs = obs[: self.no_obs + 1]
        self.set_observations(obs)
        if "state" in header:
            self.states = rows[: self.no_obs + 1, header.index("state")].reshape(-1, 1).copy()

    def generate_data(self, rng=None, no_obs=None):
        """Synthetic data by the model recursion (helpers/data_handling.py:92-113): states[0] is the
        initial state, obs[0] the dummy 0.  `rng` is a numpy Generator / RandomState."""
        if no_obs is not None:
            self.no_obs = int(no_obs)
        rng = rng if rng is not None else np.random.RandomState(12345)
        T = self.no_obs
        x = np.zeros(T + 1)
        y = np.zeros(T + 1)
        x[0] = self.initial_state
        for t in range(1, T + 1):
            x[t] = self._simulate_state(x[t - 1], y[t - 1], rng.standard_normal())
            y[t] = self._simulate_obs(x[t], rng.standard_normal())
        self.states = x.reshape(-1, 1)
        self.obs = y.reshape(-1, 1)


class LinearGaussianModel(StateSpaceModel):
    """x[t+1] = mu + phi (x[t]-mu) + sigma_v v[t];  y[t] = x[t] + sigma_e e[t]
    (models/linear_gaussian_model.py:27-75)."""
    model_kind = "lgss"
    name = "Linear Gaussian state-space model with four parameters."
    file_prefix = "linear_gaussian_model"
    PARAM_NAMES = ("mu", "phi", "sigma_v", "sigma_e")
    TRANSFORMS = {"mu": "identity", "phi": "tanh", "sigma_v": "log", "sigma_e": "log"}
    DEFAULT_PARAMS = (0.2, 0.75, 1.0, 0.1)
    DEFAULT_PRIORS = {"mu": ("normal", 0.0, 1.0), "phi": ("normal", 0.5, 1.0),
                      "sigma_v": ("gamma", 2.0, 2.0), "sigma_e": ("gamma", 2.0, 2.0)}

    def _simulate_state(self, x, y_prev, v):
        p = self.params
        return p["mu"] + p["phi"] * (x - p["mu"]) + p["sigma_v"] * v

    def _simulate_obs(self, x, e):
        return x + self.params["sigma_e"] * e


class StochasticVolatilityModel(StateSpaceModel):
    """x[t+1] = mu + phi (x[t]-mu) + sigma_v v[t];  y[t] = exp(x[t]/2) e[t]
    (models/stochastic_volatility_model.py:27-72)."""
    model_kind = "sv"
    name = "Stochastic volatility model with three parameters."
    file_prefix = "stochastic_volatility_model"
    PARAM_NAMES = ("mu", "phi", "sigma_v")
    TRANSFORMS = {"mu": "identity", "phi": "tanh", "sigma_v": "log"}
    DEFAULT_PARAMS = (0.2, 0.75, 1.0)
    DEFAULT_PRIORS = {"mu": ("normal", 0.0, 1.0), "phi": ("normal", 0.95, 0.05),
                      "sigma_v": ("gamma", 2.0, 10.0)}

    def _simulate_state(self, x, y_prev, v):
        p = self.params
        return p["mu"] + p["phi"] * (x - p["mu"]) + p["sigma_v"] * v

    def _simulate_obs(self, x, e):
        return math.exp(0.5 * x) * e


class StochasticVolatilityModelLeverage(StochasticVolatilityModel):
    """SV model with leverage: corr(v[t], e[t]) = rho
    (models/stochastic_volatility_model_leverage.py:27-75).  Synthetic data use the
    model-consistent recursion x[t] | x[t-1], y[t-1] (python/README.md:103)."""
    model_kind = "sv_leverage"
    name = "Stochastic volatility model with leverage and four parameters."
    file_prefix = "stochastic_volatility_model_leverage"
    PARAM_NAMES = ("mu", "phi", "sigma_v", "rho")
    TRANSFORMS = {"mu": "identity", "phi": "tanh", "sigma_v": "log", "rho": "tanh"}
    DEFAULT_PARAMS = (0.2, 0.75, 1.0, -0.5)
    DEFAULT_PRIORS = {"mu": ("normal", 0.0, 1.0), "phi": ("normal", 0.95, 0.05),
                      "sigma_v": ("gamma", 2.0, 10.0), "rho": ("normal", 0.0, 1.0)}

    def _simulate_state(self, x, y_prev, v):
        p = self.params
        mean = p["mu"] + p["phi"] * (x - p["mu"]) + p["sigma_v"] * p["rho"] * math.exp(-0.5 * x) * y_prev
        return mean + math.sqrt(1.0 - p["rho"] ** 2) * p["sigma_v"] * v


MODEL_CLASSES = {"lgss": LinearGaussianModel, "sv": StochasticVolatilityModel,
                 "sv_leverage": StochasticVolatilityModelLeverage}
# dispatch for model objects of the reference itself (by class name) or of this module
KIND_BY_CLASS_NAME = {"LinearGaussianModel": "lgss", "StochasticVolatilityModel": "sv",
                      "StochasticVolatilityModelLeverage": "sv_leverage"}


def model_kind_of(model):
    kind = getattr(model, "model_kind", None)
    if kind in MODEL_CLASSES:
        return kind
    for cls in type(model).__mro__:
        if cls.__name__ in KIND_BY_CLASS_NAME:
            return KIND_BY_CLASS_NAME[cls.__name__]
    raise TypeError("no device functor for model class %r: only the linear Gaussian and the two "
                    "stochastic volatility models run on the GPU (there is no generic/CPU fallback)"
                    % type(model).__name__)
