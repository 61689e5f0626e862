"""Where pipeline.analyse_tracks spends its time (parent process): cProfile over 32 three-minute stereo tracks."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import pipeline, synth  # noqa: E402
from track_analyser_b200.utils import AudioInput  # noqa: E402

n_tracks = int(sys.argv[1]) if len(sys.argv) > 1 else 32
workers = int(sys.argv[2]) if len(sys.argv) > 2 else None
sr = 44_100
base = [synth.synth_track(1 + i, 180.0, sr, 2) for i in range(4)]
if os.environ.get("TA_PINNED"):   # like bench.py: the stereo pairs live in pinned host memory
    import torch

    pinned = [torch.from_numpy(x).pin_memory() for x in base]
    base = [t.numpy() for t in pinned]
audios = [AudioInput(samples=np.mean(x, axis=0), sample_rate=sr, stereo_samples=x) for x in base]
srcs = [audios[i % 4] for i in range(n_tracks)]
pipeline.analyse_tracks(srcs[:8], workers=workers)
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
out = pipeline.analyse_tracks(srcs, workers=workers)
pr.disable()
dt = time.perf_counter() - t0
print(f"{n_tracks} tracks in {dt:.3f} s = {dt / n_tracks * 1e3:.1f} ms per track = {n_tracks * 180.0 / dt:.0f} x real-time")
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
t0 = time.perf_counter()
pipeline.analyse_track(srcs[0])
print(f"single analyse_track: {(time.perf_counter() - t0) * 1e3:.1f} ms")
t0 = time.perf_counter()
pipeline.analyse_tracks(srcs[:8], workers=0)
print(f"workers=0, 8 tracks: {(time.perf_counter() - t0) / 8 * 1e3:.1f} ms per track (kernels + host stages in this process)")
pr = cProfile.Profile()
pr.enable()
pipeline.analyse_tracks(srcs[:8], workers=0)
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
