"""Test signals: the reference's own test inputs (restated) plus edge cases."""

import math

import numpy as np


def sine(freq, sr=22_050, seconds=1.0, amp=1.0, phase=0.0):
    t = np.linspace(0, seconds, int(sr * seconds), endpoint=False)
    return (amp * np.sin(2 * np.pi * freq * t + phase)).astype(np.float32)


def minus18_sine(sr=44_100, seconds=1.0, freq=1000.0):
    """reference tests/test_loudness.py:20-30"""
    t = np.linspace(0.0, seconds, int(sr * seconds), endpoint=False)
    peak = 10 ** (-18.0 / 20.0) * math.sqrt(2.0)
    return (peak * np.sin(2.0 * np.pi * freq * t)).astype(np.float32)


def noisy_click_track(bpm=120.0, bars=64, sr=48_000, noise_level=0.02):
    """reference tests/test_tempo.py:10-36"""
    total_beats = bars * 4
    period = 60.0 / bpm
    length = int(total_beats * period * sr)
    click = np.zeros(length, dtype=np.float32)
    beat_samples = (np.arange(total_beats) * period * sr).astype(int)
    click_length = int(0.01 * sr)
    decay = np.exp(-np.linspace(0.0, 6.0, click_length))
    for idx in beat_samples:
        end = min(length, idx + click_length)
        click[idx:end] += decay[: end - idx]
    rng = np.random.default_rng(1234)
    noise = rng.normal(scale=noise_level, size=length)
    return (click + noise.astype(np.float32)).astype(np.float32), sr, beat_samples / sr


def tiny_click(sr=44_100):
    """reference scripts/make_tiny_click.py:20-53 (float values before the PCM16 write)."""
    def click(freq, amp):
        n = int(0.03 * sr)
        t = np.linspace(0.0, 0.03, n, endpoint=False)
        return (amp * np.sin(2 * np.pi * freq * t) * np.exp(-t * 50.0)).astype(np.float32)

    spb = 60.0 / 120
    reg, acc = click(1000.0, 0.6), click(1500.0, 0.9)
    total = int(np.ceil(4 * spb * sr)) + reg.shape[0]
    audio = np.zeros(total, dtype=np.float32)
    for beat in range(4):
        start = int(round(beat * spb * sr))
        w = acc if beat == 0 else reg
        audio[start:start + w.shape[0]] += w[: total - start]
    audio = np.clip(audio, -1.0, 1.0)
    return (np.round(audio * 32767.0) / 32768.0).astype(np.float32)  # PCM16 round trip like the WAV on disk


def drums_muted_track(sr=22_050, duration=32.0):
    """reference tests/test_structure.py:10-27: 110 Hz sine + 0.5 s drum hits muted between 12 s and 20 s."""
    t = np.linspace(0.0, duration, int(sr * duration), endpoint=False)
    harmonic = 0.3 * np.sin(2 * np.pi * 110.0 * t)
    drum_times = np.arange(0.0, duration, 0.5)
    active = drum_times[(drum_times < 12.0) | (drum_times >= 20.0)]
    drums = np.zeros_like(t)
    hit = int(sr * 0.05)
    env = np.linspace(1.0, 0.0, hit, dtype=np.float32)
    for tm in active:
        a = int(tm * sr)
        b = min(len(drums), a + hit)
        if b > a:
            drums[a:b] += env[: b - a]
    return (harmonic + drums).astype(np.float32), sr, np.arange(0.0, duration, 0.5)


def triad_progression(sr=22_050, duration=1.0):
    """reference tests/test_harmony.py:11-36: C, F, G, C major triads (Hann envelopes), peak-normalised."""
    def triad(root):
        t = np.linspace(0.0, duration, int(sr * duration), endpoint=False)
        chord = np.zeros_like(t)
        for k in (0, 4, 7):
            chord += np.sin(2 * np.pi * 440.0 * 2.0 ** ((root + k - 69) / 12.0) * t)  # librosa.midi_to_hz
        env = np.hanning(t.size)
        return (chord * env / np.max(env)).astype(np.float32)

    x = np.concatenate([triad(60), triad(65), triad(67), triad(60)])
    x /= np.max(np.abs(x))
    return x.astype(np.float32), sr
