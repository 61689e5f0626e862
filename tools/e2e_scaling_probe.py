"""Why does the end-to-end leg lose efficiency as ranks are added?  Run under torchrun on N GPUs: every rank times
(a) the bench's e2e pipeline (engine.HostPipeline, every frontend output copied back),
(b) the same bytes per chunk as two plain copies (one up, one down) with the same three-buffer rotation and no kernels,
(c) the same bytes as one long copy per direction (what bench.py's PCIe probe does),
and prints its own seconds per 128-track step for each, so imbalance between ranks is visible."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import engine, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device(f"cuda:{local}")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
barrier = (lambda: (torch.cuda.synchronize(), dist.barrier())) if world > 1 else torch.cuda.synchronize
SR, nt, chunk, seconds = 44_100, 128, 8, 180.0
x = synth.synth_track(5 + rank, seconds, SR, 2)
n = x.shape[1]
pool = [torch.from_numpy(np.roll(x, 1000 * i, axis=1).reshape(-1).copy()).pin_memory() for i in range(4)]
tracks = [pool[i % 4] for i in range(nt)]
plan = engine.Plan(SR, 2048, 512, 128, device=local)
pipe = engine.HostPipeline(plan, n, 2, chunk, engine.FRONTEND_OUTPUTS)
up_bytes, down_bytes = chunk * 2 * n * 4, pipe.d2h_bytes_per_chunk
n_chunks = nt // chunk


def timed(fn, reps=2):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    mine = (time.perf_counter() - t0) / reps
    barrier()
    return mine


def copies_only():
    nb = 3
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    ev_up = [torch.cuda.Event() for _ in range(nb)]
    ev_dn = [torch.cuda.Event() for _ in range(nb)]
    for ci in range(n_chunks):
        b = ci % nb
        ev_dn[b].synchronize()
        with torch.cuda.stream(s_up):
            d_up[b].copy_(h_up, non_blocking=True)
            ev_up[b].record(s_up)
        with torch.cuda.stream(s_dn):
            s_dn.wait_event(ev_up[b])
            h_dn[b].copy_(d_dn[b], non_blocking=True)
            ev_dn[b].record(s_dn)
    torch.cuda.synchronize()


def long_copies():
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(n_chunks):
        with torch.cuda.stream(s_up):
            d_up[0].copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_dn[0].copy_(d_dn[0], non_blocking=True)
    torch.cuda.synchronize()


a = timed(lambda: pipe.run(tracks, None))
h_up = torch.empty(up_bytes, dtype=torch.uint8, pin_memory=True)
d_up = [torch.empty(up_bytes, dtype=torch.uint8, device=dev) for _ in range(3)]
h_dn = [torch.empty(down_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(3)]
d_dn = [torch.empty(down_bytes, dtype=torch.uint8, device=dev) for _ in range(3)]
b = timed(copies_only)
c = timed(long_copies)
res = torch.tensor([a, b, c], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(res) for _ in range(world)]
if world > 1:
    dist.all_gather(allr, res)
else:
    allr = [res]
if rank == 0:
    gb = (nt // chunk) * down_bytes / 1e9
    print(f"per 128-track step and rank: {gb:.2f} GB down, {(nt // chunk) * up_bytes / 1e9:.2f} GB up; seconds per step (down GB/s)")
    for r, t in enumerate(allr):
        t = t.tolist()
        print(f"rank {r}: pipeline {t[0]:.3f} ({gb / t[0]:.1f})   plain chunk copies {t[1]:.3f} ({gb / t[1]:.1f})   back-to-back copies {t[2]:.3f} ({gb / t[2]:.1f})")
if world > 1:
    dist.destroy_process_group()
