"""cProfile of pipeline.analyse_track() on one 3-minute 44.1 kHz stereo track (host-side cost breakdown)."""
import cProfile, os, pstats, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import pipeline, synth
from track_analyser_b200.utils import AudioInput

sr = 44_100
x = synth.synth_track(synth.DEFAULT_SEED, 180.0, sr, 2)
audio = AudioInput(samples=np.mean(x, axis=0), sample_rate=sr, stereo_samples=x)
for _ in range(2):
    pipeline.analyse_track(audio)
pr = cProfile.Profile()
pr.enable()
pipeline.analyse_track(audio)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(int(sys.argv[1]) if len(sys.argv) > 1 else 40)
