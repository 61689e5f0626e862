"""Runs the fused frontend with every output (HPSS, true peak, MFCC included) on 32 three-minute stereo tracks: the
target for `ncu -k regex:...` sweeps over the small kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from track_analyser_b200 import engine, runtime, synth

sr = 44_100
plan = runtime.get_plan(sr)
x = synth.synth_track(synth.DEFAULT_SEED, 180.0, sr, 2)
batch = engine.upload(plan, [x] * 32)
bufs = engine.FrontendBuffers(batch, tuple(o for o in engine.ALL_OUTPUTS if o != "cqt_mag"))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    engine.run_device(plan, batch, bufs)
torch.cuda.synchronize()
print("ok", engine.launch_count())
