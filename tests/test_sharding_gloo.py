"""world_size-2 gloo test of the N>1 path: shard by track, compute locally, gather on rank 0."""

import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from track_analyser_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tracks = [np.full(100 + 10 * i, float(i), dtype=np.float32) for i in range(7)]

    def compute(local):  # stand-in for engine.analyse_batch on this rank's GPU
        return [{"sum": float(t.sum()), "rank": rank} for t in local]

    out = sharding.analyse_sharded(tracks, compute, rank=rank, world=world)
    if rank == 0:
        ret["sums"] = [o["sum"] for o in out]
        ret["ranks"] = [o["rank"] for o in out]
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        expect = [float(i) * (100 + 10 * i) for i in range(7)]
        assert list(ret["sums"]) == expect
        assert set(ret["ranks"]) == {0, 1}


def test_partition_by_rank_speed():
    """Shards in proportion to per-rank weights (the host-link rate each GPU sustains): equal lengths -> contiguous blocks
    sized by weight; ragged -> greedy earliest-finish; every track exactly once; bad weights rejected."""
    import pytest

    from track_analyser_b200 import sharding

    parts = sharding.partition([100] * 1024, 8, weights=[9.1] * 4 + [13.3] * 4)
    assert [len(p) for p in parts] == [104] * 4 + [152] * 4
    assert sorted(i for p in parts for i in p) == list(range(1024))
    assert all(p == list(range(p[0], p[-1] + 1)) for p in parts)
    lengths = [5, 9, 3, 7, 7, 2, 8, 4, 6]
    parts = sharding.partition(lengths, 2, weights=[1.0, 3.0])
    assert sorted(i for p in parts for i in p) == list(range(len(lengths)))
    load = [sum(lengths[i] for i in p) for p in parts]
    assert abs(load[0] / 1.0 - load[1] / 3.0) <= max(lengths)   # finish times within one track of each other
    assert sharding.partition([100] * 10, 3, weights=None) == sharding.partition([100] * 10, 3)
    for bad in ([1.0], [1.0, 0.0, 1.0], [1.0, float("nan"), 1.0]):
        with pytest.raises(ValueError):
            sharding.partition([1, 2, 3], 3, weights=bad)
