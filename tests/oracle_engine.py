"""A CPU stand-in for `qnmh_sysid2018_b200.engine.Engine`, backed by the oracle.  TEST ONLY: it lets
the CPU tier exercise the host-side logic of the drop-in estimator and of the batch sharding
(results dict, prior folding, gather) without a GPU.  The product never constructs it."""
import numpy as np

from oracle import pf_oracle as orc

KINDS = {"lgss": orc.MODEL_LGSS, "sv": orc.MODEL_SV, "sv_leverage": orc.MODEL_SVL}


class OracleEngine(object):
    def __init__(self, model="lgss", dtype="float64", device=-1):
        self.model, self.model_id = model, KINDS[model]
        self.no_params = len(orc.PARAM_NAMES[self.model_id])
        self._cfg = None

    def configure(self, obs, theta, no_particles, fixed_lag=0, do_smoother=True,
                  generate_initial_state=True, initial_state=0.0, resampling="systematic", seed=0,
                  eval_ids=None, draws=None, store_history=False, svl_obs_model=False, store_states=False):
        theta = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        B = theta.shape[0]
        self._cfg = dict(obs=np.asarray(obs, dtype=np.float64).reshape(-1), theta=theta, N=int(no_particles),
                         L=int(fixed_lag), do_smoother=do_smoother, gen=generate_initial_state,
                         x0=initial_state, resampling=resampling, seed=int(seed),
                         eval_ids=np.arange(B) if eval_ids is None else np.asarray(eval_ids),
                         draws=draws, obs_index="model" if svl_obs_model else "numpy")
        return self

    def update_params(self, theta, seed=0, eval_ids=None):
        self._cfg["theta"] = np.atleast_2d(np.asarray(theta, dtype=np.float64))
        self._cfg["seed"] = int(seed)
        if eval_ids is not None:
            self._cfg["eval_ids"] = np.asarray(eval_ids)

    def execute(self):
        c = self._cfg
        T = len(c["obs"]) - 1
        outs = []
        for b in range(c["theta"].shape[0]):
            draws = c["draws"]
            if isinstance(draws, (list, tuple)):
                draws = draws[b]
            if draws is None:  # any reproducible stream keyed by (seed, eval_id) will do on the CPU
                rng = np.random.default_rng([c["seed"], int(c["eval_ids"][b])])
                draws = orc.random_draws(rng, c["N"], T, c["resampling"])
            out = orc.particle_filter(self.model_id, c["theta"][b], c["obs"], c["N"], draws,
                                      initial_state_value=c["x0"], generate_initial_state=c["gen"],
                                      resampling=c["resampling"], obs_index=c["obs_index"])
            if c["do_smoother"]:
                out.update(orc.fixed_lag_smoother(self.model_id, c["theta"][b], c["obs"], out, c["L"]))
            outs.append(out)
        self._outs = outs

    def fetch(self, out=None):
        c, outs = self._cfg, self._outs
        T1, P, B = len(c["obs"]), self.no_params, len(outs)
        res = dict(log_like=np.array([o["log_like"] for o in outs]),
                   filt=np.stack([o["filt_state_est"].reshape(-1) for o in outs]),
                   smo=np.zeros((B, T1)), grad_terms=np.zeros((B, P, T1)),
                   traj_index=np.array([o["trajectory_index"] for o in outs], dtype=np.int32),
                   traj=np.stack([o["state_trajectory"].reshape(-1) for o in outs]))
        if c["do_smoother"]:
            res["smo"] = np.stack([o["smo_state_est"].reshape(-1) for o in outs])
            res["grad_terms"] = np.stack([o["smo_gradient_est"] for o in outs])
        return res

    def fetch_history(self):
        o = self._outs
        return dict(particles=np.stack([x["particles"] for x in o]),
                    log_weights=np.log(np.stack([x["weights"] for x in o])),
                    ancestors=np.stack([x["ancestors"] for x in o]))

    def run(self, *args, **kwargs):
        self.configure(*args, **kwargs)
        self.execute()
        return self.fetch()

    def last_execute_ms(self):
        return 0.0
