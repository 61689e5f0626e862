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
