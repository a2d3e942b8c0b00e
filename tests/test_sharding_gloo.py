"""world_size-2 gloo test (CPU) of the multi-GPU host logic: contiguous shards keyed by global
image index decode exactly the unsharded schedule, with no data-path collective (the all_gather
here only collects results for the check), and the bench's max-over-ranks timing reduction works.
"""

import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, results):
    sys.path.insert(0, ROOT)
    import oracle
    from chambers_b200.sharding import shard_bounds, shard_kwargs
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, stop = shard_bounds(total, rank, world)
        kw = shard_kwargs(total, rank, world)
        ok = True
        for pol, ew in [(oracle.randaugment_policy(2, 10), True), (oracle.randaugment_policy(2, 10), False),
                        (oracle.autoaugment_policy(), True)]:
            local = oracle.decode_schedule(pol, 123, 4, kw["image_index_base"], stop - start, 224, 224, ew)
            gathered = [None] * world
            dist.all_gather_object(gathered, local)
            whole = oracle.decode_schedule(pol, 123, 4, 0, total, 224, 224, ew)
            ok = ok and bool((np.concatenate(gathered) == whole).all())
        # bench.py's timing reduction: max over ranks
        t = torch.tensor([1.0 + rank], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == float(world)
        counts = torch.tensor([stop - start], dtype=torch.int64)
        dist.all_reduce(counts)
        ok = ok and int(counts) == total
        results[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [64, 37])
def test_two_rank_sharding_reproduces_the_unsharded_schedule(total):
    world = 2
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), total, results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}


def test_shard_bounds_cover_everything():
    sys.path.insert(0, ROOT)
    from chambers_b200.sharding import shard_bounds
    for total in (0, 1, 7, 256, 4096):
        for world in (1, 2, 4, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
