"""Loading the golden fixtures (tests/golden/*.npz) and replaying them through the oracle."""
import glob
import hashlib
import os

import numpy as np

from oracle import pf_oracle as orc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_IDS = {"lgss": orc.MODEL_LGSS, "sv": orc.MODEL_SV, "sv_leverage": orc.MODEL_SVL}


def golden_names():
    names = [os.path.basename(p)[:-4] for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))]
    return [n for n in names if n not in ("kalman_lgss", "kalman_rts_lgss", "mh_chains_ref")]


def load(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        g = {k: z[k] for k in z.files}
    for k in ("model_kind", "resampling_method"):
        g[k] = str(g[k])
    for k in ("no_particles", "fixed_lag", "seed"):
        g[k] = int(g[k])
    g["generate_initial_state"] = bool(g["generate_initial_state"])
    g["initial_state"] = float(g["initial_state"])
    g["log_like"] = float(g["log_like"])
    g["model_id"] = MODEL_IDS[g["model_kind"]]
    g["param_names"] = [str(s) for s in g["param_names"]]
    g["params_to_estimate"] = [str(s) for s in g["params_to_estimate"]]
    g["T"] = len(g["obs"]) - 1
    return g


def draws_for(g):
    return orc.draws_from_legacy_seed(g["seed"], g["no_particles"], g["T"],
                                      g["generate_initial_state"], g["resampling_method"])


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def prior_dicts(g):
    from collections import OrderedDict
    pg = OrderedDict(zip(g["param_names"], g["log_prior_gradient"]))
    ph = OrderedDict(zip(g["param_names"], g["log_prior_hessian"]))
    return pg, ph
