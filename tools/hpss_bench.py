"""Device time of the HPSS curves (K9) per 3-minute track: fused run with and without the two HPSS outputs."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from track_analyser_b200 import engine, runtime, synth

sr, n_tracks = 44_100, 32
plan = runtime.get_plan(sr)
x = synth.synth_track(synth.DEFAULT_SEED, 180.0, sr, 2)
batch = engine.upload(plan, [x] * n_tracks)


def timed(outputs, reps=5):
    bufs = engine.FrontendBuffers(batch, outputs)
    for _ in range(2):
        engine.run_device(plan, batch, bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.run_device(plan, batch, bufs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


base = timed(("magnitude",))
full = timed(("magnitude", "hpss_harmonic", "hpss_percussive"))
print(f"magnitude only {base:.2f} ms, + HPSS {full:.2f} ms  =>  HPSS {(full - base) / n_tracks:.3f} ms per 3-minute track")
