"""Track-level sharding across the GPUs of one box (no data-path collective).

Tracks are independent (reference: pipeline.py:32 analyses one source per call),
so a batch is partitioned by track: contiguous blocks for equal lengths,
longest-processing-time-first for ragged batches.  One process per GPU; the only
communication is the final gather of small per-track results onto rank 0 through
``torch.distributed`` (NCCL on the GPU box, gloo in the CPU tests).
"""

from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def partition(lengths: Sequence[int], world: int, weights: Sequence[float] | None = None) -> list[list[int]]:
    """Assign track indices to ranks.  Equal lengths -> contiguous blocks; otherwise LPT greedy.

    ``weights``: relative speed of every rank (e.g. the host-link rate each GPU reaches when all of them copy at once:
    on a box where the GPUs do not share the host evenly, the end-to-end time of equal shards is that of the slowest
    link).  Shards are then sized in proportion, so that all ranks finish together."""
    lengths = [int(v) for v in lengths]
    n = len(lengths)
    if world <= 0:
        raise ValueError("world must be positive")
    if weights is None:
        w = [1.0] * world
    else:
        w = [float(v) for v in weights]
        if len(w) != world or min(w) <= 0.0 or not all(np.isfinite(w)):
            raise ValueError("weights must be one positive finite number per rank")
    if n == 0:
        return [[] for _ in range(world)]
    if len(set(lengths)) == 1:
        cum = np.concatenate([[0.0], np.cumsum(w)]) / float(sum(w))
        bounds = [int(round(n * c)) for c in cum] if weights is not None else [(n * r) // world for r in range(world + 1)]
        return [list(range(bounds[r], bounds[r + 1])) for r in range(world)]
    shards = [[] for _ in range(world)]
    load = [0.0] * world
    for i in sorted(range(n), key=lambda i: (-lengths[i], i)):
        r = min(range(world), key=lambda r: ((load[r] + lengths[i]) / w[r], r))   # finishes this track first
        shards[r].append(i)
        load[r] += lengths[i]
    return [sorted(s) for s in shards]


def analyse_sharded(tracks: Sequence[np.ndarray], compute: Callable[[list], list], *, rank: int, world: int,
                    gather: bool = True, weights: Sequence[float] | None = None):
    """Run ``compute`` on this rank's shard and (optionally) gather ordered results on rank 0.

    ``compute(list_of_tracks) -> list_of_results`` is the per-GPU frontend call
    (``engine.analyse_batch`` bound to this rank's plan).  Returns the full ordered
    result list on rank 0 (``None`` elsewhere) when ``gather`` is set, else the local list.
    ``weights``: see ``partition`` (every rank must pass the same values).
    """
    shards = partition([t.shape[-1] for t in tracks], world, weights)
    mine = shards[rank]
    local = compute([tracks[i] for i in mine]) if mine else []
    if not gather or world == 1:
        if world == 1:
            return local
        return local
    import torch.distributed as dist

    gathered = [None] * world if rank == 0 else None
    dist.gather_object(list(zip(mine, local)), gathered, dst=0)
    if rank != 0:
        return None
    out = [None] * len(tracks)
    for part in gathered:
        for idx, res in part:
            out[idx] = res
    return out
