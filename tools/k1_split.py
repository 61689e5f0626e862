"""K1 time split: runs the fused STFT kernel with different output sets (128 x 180 s stereo) and prints ms per launch."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import engine, synth

SR = 44_100
nt = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = synth.synth_track(1, 180.0, SR, 2)
n = x.shape[1]
plan = engine.Plan(SR, 2048, 512, 128, device=0)
pcm = torch.from_numpy(x.reshape(-1)).cuda().repeat(nt)
batch = engine.DeviceBatch(plan, pcm, np.arange(nt, dtype=np.int64) * 2 * n, np.full(nt, n, dtype=np.int64), 2)
sets = {
    "band_energy only (FFT + per-bin sums)": ("band_energy",),
    "frame_max only (FFT + (c))": ("frame_max",),
    "magnitude only (FFT + (a))": ("magnitude",),
    "mel only (FFT + (b))": ("mel",),
    "magnitude+mel": ("magnitude", "mel"),
    "rolloff only (FFT + exact chain)": ("rolloff_bin",),
    "all but rolloff": ("magnitude", "mel", "ltas", "centroid", "band_energy", "frame_max"),
    "all K1 outputs": ("magnitude", "mel", "ltas", "centroid", "rolloff_bin", "band_energy", "frame_max"),
}
for name, outs in sets.items():
    bufs = engine.FrontendBuffers(batch, outs)
    for _ in range(2):
        engine.run_device(plan, batch, bufs, "stft")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        engine.run_device(plan, batch, bufs, "stft")
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:45s} {e0.elapsed_time(e1) / 5:8.3f} ms")
    del bufs
