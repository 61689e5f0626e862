"""End-to-end latency of pipeline.analyse_track() for BASELINE configs[1] (one 3-minute 44.1 kHz stereo track),
host arrays in, dataclasses out, and of the CPU oracle frontend for the same track (one process)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from track_analyser_b200 import pipeline, synth, runtime
from track_analyser_b200.utils import AudioInput

sr = 44_100
x = synth.synth_track(synth.DEFAULT_SEED, 180.0, sr, 2)
audio = AudioInput(samples=np.mean(x, axis=0), sample_rate=sr, stereo_samples=x)
stages = {}
def cb(name, t=[time.perf_counter()]):
    now = time.perf_counter(); stages[name] = stages.get(name, 0.0) + now - t[0]; t[0] = now
for i in range(3):
    stages.clear(); t0 = time.perf_counter(); cb.__defaults__[0][0] = t0
    res = pipeline.analyse_track(audio, progress_callback=cb)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"analyse_track run {i}: {dt*1e3:.1f} ms  ->  {180.0/dt:.0f} x real-time;  stages (ms): " +
          ", ".join(f"{k} {v*1e3:.1f}" for k, v in stages.items()))
print("bpm", res.beat.bpm, "segments", len(res.structure.segments), "lufs", res.loudness.integrated_lufs,
      "true peak", res.loudness.true_peak_dbfs)
if "--cpu" in sys.argv:
    import bench
    t0 = time.perf_counter(); bench._oracle_frontend(x); dt = time.perf_counter() - t0
    print(f"CPU oracle frontend (one process, STFTs shared, no HPSS/true peak): {dt*1e3:.0f} ms -> {180.0/dt:.0f} x real-time")
