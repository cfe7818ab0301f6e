"""MH-driver batching: K independent Metropolis-Hastings chains -- Monte Carlo
replicates (`run_script.py:18-24`), the four algorithm variants of example 2
(`scripts/example2_lgss_particles.py:61-116`), parallel chains -- share ONE persistent launch per
iteration instead of K sequential estimator calls.

The MH driver stays what it is (the reference's `MetropolisHastings`, or any other caller of the
estimator surface): it calls `state.smoother(model)` / `state.filter(model)` synchronously once per
iteration (`metropolis_hastings.py:326-329,413-416`).  The pool hands every chain a PROXY estimator
with that surface.  A proxy call posts the chain's current parameters and blocks; when every live
chain has posted, the last arrival evaluates all K parameter vectors with one
`pfb200_update_params` + `pfb200_execute` (B = K filters side by side) and hands each chain its own
`results` dict -- score / Hessian / prior folding per chain as in `ParticleMethodsB200`.  Chains run in
threads (the GPU does the work; the GIL only serialises the small host parts), so K latency-bound
N = 2000 evaluations (config C1: 3.7 ms each) become one throughput-bound batch (config C3).

Random streams: chain k's i-th evaluation uses the Philox stream (seed, eval_id = k * 2^32 + i): the
batch reproduces, bit for bit, K estimators run one after the other with those stream ids."""
import threading

import numpy as np

from .models import model_kind_of
from .particle_methods import STATE_HISTORY_BUDGET, ParticleMethodsB200


class _ProxyEstimator(ParticleMethodsB200):
    """What a chain sees: the `ParticleMethods` surface; evaluation is delegated to the pool."""

    def __init__(self, pool, slot, settings):
        super().__init__(settings)
        self._pool, self._slot = pool, slot
        self.name = "Particle methods (B200 CUDA implementation, lock-step batch of %d)" % pool.size

    def filter(self, model):
        self.name = "Bootstrap particle filter (B200, lock-step batch)"
        out, row = self._pool._evaluate(self._slot, model, False)
        self._store_filter_results(model, None, out, row)

    def smoother(self, model):
        self.name = "Bootstrap particle filter and fixed-lag particle smoother (B200, lock-step batch)."
        if self.settings['fixed_lag'] < 1 and self.settings['estimate_gradient']:
            raise NameError("name 'next_ancestor' is not defined (fixed_lag must be >= 1 when "
                            "estimate_gradient is set, standard.py:161)")
        out, row = self._pool._evaluate(self._slot, model, True)
        self._store_filter_results(model, None, out, row)
        self._store_smoother_results(model, out, row)


class LockstepEstimatorPool(object):
    """K proxy estimators over one batched engine.  `engine_factory(kind, dtype, device)` defaults to the
    CUDA `Engine`; the CPU test tier injects a stand-in to exercise the scheduling logic."""

    def __init__(self, size, settings=None, engine_factory=None):
        self.size = int(size)
        self._engine_factory = engine_factory
        self._engine = None
        self._engine_key = None
        self._cfg_key = None
        self._last_ms = None
        self._cond = threading.Condition()
        self._live = set(range(self.size))
        self._pending = {}        # slot -> (theta, kind, obs, want_smoother)
        self._generation = 0
        self._results = None
        self._error = None
        self._calls = np.zeros(self.size, dtype=np.uint64)
        self._theta = None
        self.batches = 0          # launches so far
        self.evaluations = 0      # chain evaluations served
        self.device_ms = 0.0
        self.estimators = [_ProxyEstimator(self, k, dict(settings or {})) for k in range(self.size)]

    def estimator(self, slot):
        return self.estimators[slot]

    def retire(self, slot):
        """A finished (or failed) chain leaves the barrier; the others go on without it."""
        with self._cond:
            self._live.discard(slot)
            if self._live and len(self._pending) == len(self._live) and self._error is None:
                self._run_batch_locked()

    # ------------------------------------------------------------------------------------------
    def _evaluate(self, slot, model, want_smoother):
        est = self.estimators[slot]
        theta = np.asarray(model.get_all_params(), dtype=np.float64)
        request = (theta, model_kind_of(model), np.asarray(model.obs, dtype=np.float64).reshape(-1), bool(want_smoother))
        with self._cond:
            if slot not in self._live:
                raise RuntimeError("estimator %d was retired from the pool" % slot)
            self._pending[slot] = request
            generation = self._generation
            if len(self._pending) == len(self._live):
                self._run_batch_locked()
            else:
                while self._generation == generation and self._error is None:
                    self._cond.wait()
            if self._error is not None:
                raise RuntimeError("lock-step batch failed: %s" % (self._error,))
            est.last_device_ms = self._last_ms
            return self._results, slot

    def _run_batch_locked(self):
        """Called with the lock held by the thread that completed the batch."""
        try:
            self._launch()
        except Exception as err:  # wake everybody up with the error instead of dead-locking them
            self._error = err
        self._pending = {}
        self._generation += 1
        self._cond.notify_all()

    def _launch(self):
        any_slot = next(iter(self._pending))
        s = self.estimators[any_slot].settings
        theta0, kind, obs, _ = self._pending[any_slot]
        P = len(theta0)
        for slot, (th, kd, ob, _) in self._pending.items():
            if kd != kind or len(th) != P or ob.shape != obs.shape or not np.array_equal(ob, obs):
                raise ValueError("a lock-step batch needs one model class and one observation record")
        if self._theta is None:
            self._theta = np.tile(theta0, (self.size, 1))
        do_smoother = any(r[3] for r in self._pending.values())
        for slot, (th, _, _, _) in self._pending.items():
            self._theta[slot] = th
            self._calls[slot] += 1
        # stream ids: chain k, evaluation i -> k * 2^32 + (i - 1), identical to K sequential estimators
        eval_ids = (np.arange(self.size, dtype=np.uint64) << np.uint64(32)) + np.maximum(self._calls, 1) - np.uint64(1)
        if s['filter_type'] == 'fully_adapted':
            kind = 'lgss_fully_adapted'
        states = s['store_states']
        if states == 'auto':
            states = self.size * len(obs) * int(s['no_particles']) * 8 <= STATE_HISTORY_BUDGET
        states = bool(states)
        for est in self.estimators:
            est._have_rows = states
        key = (obs.tobytes(), s['no_particles'], s['fixed_lag'], do_smoother, s['generate_initial_state'],
               float(s['initial_state']), s['resampling_method'], states, s['svl_obs_index'])
        seed = int(s['seed']) + int(s['seed_offset'])
        engine_key = (kind, s['dtype'], s['device'])
        if self._engine is None or engine_key != self._engine_key:
            factory = self._engine_factory
            if factory is None:
                from .engine import Engine  # raises if libpfb200.so or the GPU is missing
                factory = Engine
            self._engine = factory(kind, s['dtype'], s['device'])
            self._engine_key, self._cfg_key = engine_key, None
        if key != self._cfg_key:
            self._engine.configure(obs, self._theta, s['no_particles'], s['fixed_lag'], do_smoother,
                                   s['generate_initial_state'], s['initial_state'], s['resampling_method'],
                                   seed=seed, eval_ids=eval_ids, store_states=states,
                                   svl_obs_model=(s['svl_obs_index'] == 'model'))
            self._cfg_key = key
        else:
            self._engine.update_params(self._theta, seed=seed, eval_ids=eval_ids)
        self._engine.execute()
        self._results = self._engine.fetch()
        try:
            self._last_ms = self._engine.last_execute_ms()
        except Exception:
            self._last_ms = None
        self.batches += 1
        self.evaluations += len(self._pending)
        self.device_ms += self._last_ms or 0.0

    # ------------------------------------------------------------------------------------------
    def run(self, chain_functions):
        """Runs `chain_functions[k](estimator_k)` in K threads and returns their results in order.  A
        chain that raises is retired (the others finish) and its exception is re-raised at the end."""
        if len(chain_functions) != self.size:
            raise ValueError("need one chain function per estimator")
        results, errors = [None] * self.size, [None] * self.size

        def work(k):
            try:
                results[k] = chain_functions[k](self.estimators[k])
            except BaseException as err:  # noqa: B902 -- reported below
                errors[k] = err
            finally:
                self.retire(k)

        threads = [threading.Thread(target=work, args=(k,), name="chain-%d" % k) for k in range(self.size)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for err in errors:
            if err is not None:
                raise err
        return results
