"""Batched independent filter evaluations (MH proposals, Monte Carlo replicates) and their
sharding over the GPUs of one box.

The reference has no batching: `run_script.py:18-24` loops over Monte Carlo replicates and
`example2_lgss_particles.py:61-116` runs its four chains one after the other, one
`state.smoother(model)` call per MH iteration (`metropolis_hastings.py:413-416`).  Independent
evaluations share nothing, so they shard trivially: evaluation b goes to rank b mod R... here to
the contiguous slice `shard_range(B, rank, R)`; each rank runs its slice through one handle of
libpfb200 (filters side by side in one persistent launch) and the only communication is the final
gather of B x (1 + (P+2)(T+1)) doubles of results.  The Philox stream of an evaluation is keyed by
its GLOBAL index (eval_id), so results do not depend on R.
"""
import numpy as np


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of `n_items` for `rank`: the first n % R ranks get one
    extra item."""
    base, extra = divmod(int(n_items), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class BatchedParticleFilter(object):
    """B independent PF (+ fixed-lag smoother) evaluations on the GPU of this rank.

    `engine_factory(model_kind, dtype, device)` creates the compute engine; the default is the CUDA
    `Engine` (libpfb200).  Tests inject a CPU stand-in to exercise the sharding / gather logic with
    the gloo backend."""

    def __init__(self, model_kind, no_particles, fixed_lag, dtype="float64", device=-1, seed=87655678,
                 rank=0, world_size=1, engine_factory=None):
        if engine_factory is None:
            from .engine import Engine
            engine_factory = Engine
        self.engine = engine_factory(model_kind, dtype, device)
        self.no_particles, self.fixed_lag, self.seed = int(no_particles), int(fixed_lag), int(seed)
        self.rank, self.world_size = int(rank), int(world_size)

    def evaluate_local(self, obs, thetas, do_smoother=True, generate_initial_state=True,
                       initial_state=0.0, resampling="systematic", eval_id_offset=0):
        """Runs this rank's slice of the batch.  `thetas` is the FULL (B, P) batch on every rank.
        Returns (lo, hi, results-of-slice)."""
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        lo, hi = shard_range(thetas.shape[0], self.rank, self.world_size)
        if hi == lo:
            return lo, hi, None
        ids = np.arange(lo, hi, dtype=np.uint64) + np.uint64(eval_id_offset)
        out = self.engine.run(obs, thetas[lo:hi], self.no_particles, self.fixed_lag, do_smoother,
                              generate_initial_state, initial_state, resampling, seed=self.seed,
                              eval_ids=ids)
        return lo, hi, out

    def evaluate(self, obs, thetas, **kwargs):
        """Evaluates the whole batch across all ranks and returns the gathered results on every
        rank: dict(log_like (B,), filt (B,T+1), smo (B,T+1), grad_terms (B,P,T+1), traj_index (B,))."""
        lo, hi, out = self.evaluate_local(obs, thetas, **kwargs)
        if self.world_size == 1:
            return out
        import torch.distributed as dist
        pieces = [None] * self.world_size
        dist.all_gather_object(pieces, (lo, hi, out))
        keys = ("log_like", "filt", "smo", "grad_terms", "traj_index")
        merged = {}
        for k in keys:
            merged[k] = np.concatenate([pc[2][k] for pc in sorted(pieces, key=lambda t: t[0])
                                        if pc[2] is not None], axis=0)
        return merged
