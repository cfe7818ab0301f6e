"""One filter sharded over 2 GPUs (peer-memory exchange inside the persistent kernel) must equal the
single-GPU run with the same total number of CTAs.  Needs >= 2 GPUs (skipped on the 1-GPU tier)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

CASES = [(200000, 30, 5, 32), (100001, 20, 3, 16), (1 << 20, 12, 4, 148)]


def _inputs(N, T):
    rs = np.random.RandomState(N % 1000 + T)
    theta = np.array([0.2, 0.5, 1.0, 0.5])
    x, obs = 0.0, np.zeros(T + 1)
    for t in range(1, T + 1):
        x = theta[0] + theta[1] * (x - theta[0]) + theta[2] * rs.randn()
        obs[t] = x + theta[3] * rs.randn()
    return obs, theta


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qnmh_sysid2018_b200.distributed import DistributedParticleFilter
    res = []
    for N, T, L, gl in CASES:
        obs, theta = _inputs(N, T)
        dpf = DistributedParticleFilter("lgss", "float64", rank, ctas_per_rank=gl)
        out = dpf.run(obs, theta, N, L, seed=11, eval_id=3)
        res.append({k: out[k][0].copy() for k in ("log_like", "filt", "smo", "grad_terms")})
        del dpf
    ret[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpus_equal_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29541, ret), nprocs=2, join=True)
    from qnmh_sysid2018_b200.engine import Engine
    for ci, (N, T, L, gl) in enumerate(CASES):
        obs, theta = _inputs(N, T)
        eng = Engine("lgss", "float64", 0)
        if 2 * gl <= 148:  # same tiling on one GPU -> bitwise identical CDF, ancestors, log-likelihood
            eng.set_option("ctas_per_filter", 2 * gl)
        one = eng.run(obs, theta, N, L, True, True, 0.0, "systematic", seed=11, eval_ids=[3])
        for rank in (0, 1):
            d = ret[rank][ci]
            if 2 * gl <= 148:
                assert d["log_like"] == one["log_like"][0]
                tol = 1e-12
            else:  # different tiling: ties may move single ancestors -> statistical agreement only
                assert abs(d["log_like"] - one["log_like"][0]) < 0.05
                tol = None
            if tol:
                np.testing.assert_allclose(d["filt"], one["filt"][0], rtol=tol, atol=1e-13)
                np.testing.assert_allclose(d["smo"], one["smo"][0], rtol=tol, atol=1e-13)
                np.testing.assert_allclose(d["grad_terms"], one["grad_terms"][0], rtol=1e-10, atol=1e-11)
        np.testing.assert_array_equal(ret[0][ci]["filt"], ret[1][ci]["filt"])
