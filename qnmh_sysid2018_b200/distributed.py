"""ONE particle filter sharded over the GPUs of a box (BASELINE config C5).

One process per GPU (`torch.distributed`).  Rank r holds the particles with global indices
[r*span, (r+1)*span).  Inside the persistent kernel every CTA pushes its children to the owner of
their index range with NVLink peer stores, and the two per-step exchanges (CDF tile totals, max
log-weight) are all-gathers through peer memory with system-scope release/acquire -- no NCCL call
sits on the data path.  `torch.distributed` is only the control plane: it carries the CUDA IPC
handles between the processes once per configuration and sums the per-rank partial results
((P+2)(T+1) doubles) at the end.

Random numbers are keyed by the GLOBAL particle index, and the CDF tiling depends only on the total
number of CTAs, so the result equals a single-GPU run with `ctas_per_filter = world * ctas_per_rank`
(checked by tests/test_gpu_distributed.py)."""
import numpy as np


class DistributedParticleFilter(object):
    def __init__(self, model_kind="lgss", dtype="float64", device=None, ctas_per_rank=0):
        import torch
        import torch.distributed as dist
        from .engine import Engine
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if device is None:
            device = self.rank % max(1, torch.cuda.device_count())
        self.engine = Engine(model_kind, dtype, device)
        self.engine.set_option("dist_world", self.world)
        self.engine.set_option("dist_rank", self.rank)
        if ctas_per_rank:
            self.engine.set_option("ctas_per_filter", ctas_per_rank)
        self.device = device

    def prepare(self, obs, theta, no_particles, fixed_lag, do_smoother=True,
                generate_initial_state=True, initial_state=0.0, seed=87655678, eval_id=0):
        """Configures every rank for N_global particles and connects the peer memory."""
        import torch.distributed as dist
        self.engine.configure(obs, theta, no_particles, fixed_lag, do_smoother, generate_initial_state,
                              initial_state, "systematic", seed=seed, eval_ids=[eval_id])
        blobs = [None] * self.world
        dist.all_gather_object(blobs, self.engine.dist_export())
        self.engine.dist_connect(b"".join(blobs), self.world)
        dist.barrier()  # every counter is zero and mapped before any rank launches

    def execute(self):
        self.engine.execute()

    def fetch(self):
        """Sums the per-rank partial sums; returns the same dict as `Engine.fetch` on every rank."""
        import torch
        import torch.distributed as dist
        out = self.engine.fetch()
        parts = [out["filt"].ravel(), out["smo"].ravel(), out["grad_terms"].ravel()]
        flat = torch.from_numpy(np.concatenate(parts))
        if dist.get_backend() == "nccl":
            flat = flat.cuda(self.device)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat = flat.cpu().numpy()
        o = 0
        for key in ("filt", "smo", "grad_terms"):
            n = out[key].size
            out[key] = flat[o:o + n].reshape(out[key].shape).copy()
            o += n
        return out

    def run(self, *args, **kwargs):
        self.prepare(*args, **kwargs)
        self.execute()
        return self.fetch()
