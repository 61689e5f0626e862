"""Small ragged batches through every kernel, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import engine, synth

sr = 44_100
for n_fft, hop, mels in ((2048, 512, 128), (1024, 256, 64), (4096, 256, 256), (512, 200, 40), (256, 64, 40), (2048, 441, 128)):
    plan = engine.Plan(sr, n_fft, hop, mels, device=0)
    for ch in (1, 2):
        tracks = [synth.synth_track(3 + i, d, sr, ch) for i, d in enumerate((0.51, 1.237, 0.9))]
        tracks.append(tracks[0][..., :1001])   # shorter than one FFT frame, odd length
        outs = tuple(o for o in engine.available_outputs(plan) if o not in ("kw_blocks", "lufs"))
        res = engine.analyse_batch(plan, tracks, outs)
        assert all(np.all(np.isfinite(r["mel"])) for r in res)
    res = engine.analyse_batch(plan, [synth.synth_track(9, 1.0, sr, 2)], engine.available_outputs(plan))
    assert np.isfinite(res[0]["lufs"])
print("sanitize_small ok")
