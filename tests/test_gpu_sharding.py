"""The N > 1 product path for real: sharding.analyse_sharded with engine.analyse_batch on two GPUs (one process per GPU, no
data-path collective; rank 0 gathers the per-track results).  Skipped on a single-GPU box.  Run with -m gpu."""

import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2,
                                                  reason="needs two CUDA devices")]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


OUTS = ("onset_env", "lufs", "chroma", "tuning", "rolloff_bin", "magnitude")


def _tracks():
    from track_analyser_b200 import synth

    return [synth.synth_track(700 + i, 2.0 + 0.9 * i, 44_100, 2) for i in range(5)]


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    from track_analyser_b200 import engine, sharding

    plan = engine.Plan(44_100, 2048, 512, 128, device=rank)

    def compute(local):
        res = engine.analyse_batch(plan, local, OUTS)
        return [{k: r[k] for k in OUTS if k != "magnitude"} | {"mag_sum": float(np.sum(r["magnitude"], dtype=np.float64)), "rank": rank}
                for r in res]

    out = sharding.analyse_sharded(_tracks(), compute, rank=rank, world=world)
    if rank == 0:
        ret["out"] = out
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_analyse_sharded_two_gpus_equals_one_gpu():
    import torch.multiprocessing as mp

    from track_analyser_b200 import engine

    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        out = ret["out"]
    assert len(out) == 5 and {o["rank"] for o in out} == {0, 1}
    plan = engine.Plan(44_100, 2048, 512, 128, device=0)
    ref = engine.analyse_batch(plan, _tracks(), OUTS)
    for o, r in zip(out, ref):
        np.testing.assert_array_equal(o["onset_env"], r["onset_env"])
        np.testing.assert_array_equal(o["chroma"], r["chroma"])
        np.testing.assert_array_equal(o["rolloff_bin"], r["rolloff_bin"])
        assert o["tuning"] == r["tuning"] and o["lufs"] == pytest.approx(r["lufs"], abs=1e-9)
        assert o["mag_sum"] == pytest.approx(float(np.sum(r["magnitude"], dtype=np.float64)), rel=1e-12)
