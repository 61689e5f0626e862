"""Seeded synthetic tracks for tests and benchmarks.

A track is music-like enough to exercise every stage: an exponentially decaying
click/kick train at a per-seed tempo in [90, 135] BPM (as tests/test_tempo.py:22-29
of the reference builds one), triad pads that change every two bars (as
tests/test_harmony.py:11-21), a section with the drums muted (as
tests/test_structure.py:15-26) and Gaussian noise decorrelated between channels.
Output is planar float32 with peak <= 0.9.
"""

from __future__ import annotations

import numpy as np

DEFAULT_SEED = 13_370  # reference: utils.py:25


def synth_track(seed: int, seconds: float, sample_rate: int = 44_100, channels: int = 2,
                noise: float = 0.02) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n = int(round(seconds * sample_rate))
    t = np.arange(n, dtype=np.float64) / sample_rate
    bpm = float(rng.uniform(90.0, 135.0))
    beat = 60.0 / bpm
    # kick / click train, muted in the middle third
    n_beats = int(seconds / beat) + 1
    starts = (np.arange(n_beats) * beat * sample_rate).astype(np.int64)
    click_len = int(0.05 * sample_rate)
    env = np.exp(-np.linspace(0.0, 6.0, click_len))
    body = env * np.sin(2 * np.pi * 60.0 * np.arange(click_len) / sample_rate) + 0.5 * env * (np.arange(click_len) < 32)
    drums = np.zeros(n + click_len)
    mute_lo, mute_hi = seconds * 0.4, seconds * 0.6
    for s in starts:
        ts = s / sample_rate
        if s < n and not (mute_lo <= ts < mute_hi):
            drums[s: s + click_len] += body
    drums = drums[:n]
    # triad pads every two bars
    roots = np.array([261.63, 349.23, 392.00, 220.00])
    seg = 8 * beat
    idx = np.minimum((t / seg).astype(np.int64), 10**9) % len(roots)
    f0 = roots[idx]
    phase = 2 * np.pi * np.cumsum(f0) / sample_rate
    pad = (np.sin(phase) + np.sin(phase * 1.25) + np.sin(phase * 1.5)) / 3.0
    mono = 0.45 * drums + 0.25 * pad
    out = np.empty((channels, n), dtype=np.float64)
    for c in range(channels):
        pan = 1.0 if channels == 1 else (0.8 + 0.4 * c)
        out[c] = mono * pan + (0.15 * np.sin(2 * np.pi * (3.0 + c) * t) * pad if channels == 2 else 0.0)
        out[c] += rng.normal(scale=noise, size=n)
    peak = np.max(np.abs(out))
    if peak > 0.9:
        out *= 0.9 / peak
    return out.astype(np.float32) if channels > 1 else out[0].astype(np.float32)
