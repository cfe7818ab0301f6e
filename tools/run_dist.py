"""Distributed single-filter sweep (config C5): launch with torchrun, one process per GPU.
   python -m torch.distributed.run --nproc-per-node R --master-addr 127.0.0.1 tools/run_dist.py --log2 22 24 26"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from qnmh_sysid2018_b200.distributed import DistributedParticleFilter  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2", type=int, nargs="+", default=[22, 24])
ap.add_argument("--T", type=int, default=50)
ap.add_argument("--L", type=int, default=10)
ap.add_argument("--dtype", default="float64")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--no-smoother", action="store_true")
ap.add_argument("--ctas", type=int, default=0, help="CTAs per rank (0: automatic)")
ap.add_argument("--general", action="store_true", help="single GPU: general kernel flavour (stratified resampling)")
a = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
obs = bench.synthetic_obs("lgss", a.T)
for k in a.log2:
    N = 1 << k
    if world > 1:
        dpf = DistributedParticleFilter("lgss", a.dtype, local, ctas_per_rank=a.ctas)
        dpf.prepare(obs, bench.THETA_LGSS, N, a.L, not a.no_smoother, seed=1)
        eng = dpf.engine
    else:
        from qnmh_sysid2018_b200.engine import Engine
        eng = Engine("lgss", a.dtype, local)
        eng.configure(obs, bench.THETA_LGSS, N, a.L, not a.no_smoother, True, 0.0,
                      "stratified" if a.general else "systematic", seed=1)
    ms = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        eng.execute()
        eng.synchronize()
        ms.append(eng.last_kernel_ms())
    t = torch.tensor([min(ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ll = (dpf.fetch() if world > 1 else eng.fetch())["log_like"][0]
    if rank == 0:
        print("C5 N=2^%d T=%d %s on %d GPU(s): %.3f ms -> %.2f G particle-steps/s (%.1f us/step), ctas/rank %d, tile_cap %d, ll %.4f"
              % (k, a.T, a.dtype, world, float(t[0]), N * a.T / float(t[0]) / 1e6, float(t[0]) * 1e3 / a.T,
                 eng.info("ctas_per_rank"), eng.info("tile_cap"), ll), flush=True)
    prof = eng.debug_profile()
    if rank == 0 and prof.any():  # -DPFB_PROFILE builds: per-phase clocks of this rank's CTAs (thread 0 of each)
        names = ["A:scan(+rest)", "filt block_sum", "exchange1 wait", "prefix+bounds", "children", "block_max", "exchange2 wait",
                 "xchg_max", "A:load+exp", "X1 arrive", "X2 arrive", "B: search", "-", "B: gather + normals", "B: propagate/weight", "B: stores"]
        per_step = prof / float(a.T) / 1.965e3
        for i, nm in enumerate(names):
            print("   %-20s mean %7.2f us  max %7.2f us  min %7.2f us" % (nm, per_step[:, i].mean(), per_step[:, i].max(), per_step[:, i].min()), flush=True)
        print("   total (mean over CTAs) %.2f us/step" % per_step[:, :13].sum(axis=1).mean(), flush=True)
    del eng
    if world > 1:
        del dpf
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
