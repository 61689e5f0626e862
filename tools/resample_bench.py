"""Device time of the resampler (K11) for one 3-minute stereo track, 48 kHz -> 44.1 kHz and 22.05 kHz -> 44.1 kHz."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from track_analyser_b200 import resample as rs

for sr0, sr1 in ((48_000, 44_100), (22_050, 44_100), (96_000, 44_100)):
    x = torch.randn(2, 180 * sr0, device="cuda") * 0.1
    r = rs.get_resampler(sr0, sr1)
    for _ in range(2):
        y = r.run_device(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        y = r.run_device(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{sr0} -> {sr1}: {ms:.2f} ms per 3-minute stereo track ({180.0 / (ms * 1e-3):.0f} x real-time), out {tuple(y.shape)}")
