"""World-size-2 CPU test (gloo) of the batched-evaluation sharding: every rank evaluates its
contiguous slice and all ranks end up with the same gathered results as a single-rank run.  The
oracle-backed test engine stands in for the GPU."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    import torch.distributed as dist
    from oracle_engine import OracleEngine
    from qnmh_sysid2018_b200.batch import BatchedParticleFilter
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(3)
    obs = np.concatenate([[0.0], rs.randn(20)])
    thetas = np.array([0.2, 0.5, 1.0, 0.5]) + 0.02 * rs.randn(5, 4)
    bpf = BatchedParticleFilter("lgss", 64, 3, rank=rank, world_size=world, engine_factory=OracleEngine)
    out = bpf.evaluate(obs, thetas)
    ret[rank] = out["log_like"]
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_balanced_partition():
    from qnmh_sysid2018_b200.batch import shard_range
    for B in (0, 1, 5, 8, 4096, 4097):
        for R in (1, 2, 3, 8):
            spans = [shard_range(B, r, R) for r in range(R)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_match_one_rank():
    sys.path.insert(0, HERE)
    from oracle_engine import OracleEngine
    from qnmh_sysid2018_b200.batch import BatchedParticleFilter
    rs = np.random.RandomState(3)
    obs = np.concatenate([[0.0], rs.randn(20)])
    thetas = np.array([0.2, 0.5, 1.0, 0.5]) + 0.02 * rs.randn(5, 4)
    single = BatchedParticleFilter("lgss", 64, 3, engine_factory=OracleEngine).evaluate(obs, thetas)["log_like"]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, ret), nprocs=2, join=True)
    np.testing.assert_array_equal(ret[0], ret[1])      # every rank holds the full gathered batch
    np.testing.assert_array_equal(ret[0], single)      # and it does not depend on the number of ranks
