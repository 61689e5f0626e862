"""Input containers and coercion (mirror of the reference's ``utils.py``).

Same public names and semantics as /root/reference/src/track_analyser/utils.py:
``AudioInput`` (:28-39), ``DEFAULT_SR``/``DEFAULT_SEED`` (:24-25),
``seed_everything`` (:48-52), ``deterministic_rng`` (:42-45), ``coerce_audio``
(:73-146).  File decoding and resampling are outside the hot path (SURVEY.md
section 8f rank 4): WAV files are decoded by ``io.load_audio``; inputs whose
sample rate differs from ``target_sr`` are rejected instead of being resampled.
"""

from __future__ import annotations

import random
from dataclasses import dataclass
from pathlib import Path
from typing import Optional

import numpy as np

DEFAULT_SR = 44_100
DEFAULT_SEED = 13_370


@dataclass(slots=True)
class AudioInput:
    samples: np.ndarray
    sample_rate: int
    path: Optional[str] = None
    stereo_samples: Optional[np.ndarray] = None

    @property
    def duration(self) -> float:
        return float(len(self.samples)) / float(self.sample_rate)


def deterministic_rng(seed: int = DEFAULT_SEED) -> np.random.Generator:
    return np.random.default_rng(seed)


def seed_everything(seed: int = DEFAULT_SEED) -> None:
    np.random.seed(seed)
    random.seed(seed)


def _need_same_rate(sr: int, target_sr: int) -> None:
    if int(sr) != int(target_sr):
        raise NotImplementedError(
            f"resampling {sr} -> {target_sr} Hz is outside the B200 frontend's scope (SURVEY 8f rank 4); "
            "pass an AudioInput at its native rate instead"
        )


def _split(samples: np.ndarray, mono: bool):
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim > 1:
        return (np.mean(samples, axis=0) if mono else samples), samples
    return samples, None


def coerce_audio(source, *, target_sr: int = DEFAULT_SR, mono: bool = True) -> AudioInput:
    if isinstance(source, AudioInput):
        _need_same_rate(source.sample_rate, target_sr)
        stereo = None if source.stereo_samples is None else np.asarray(source.stereo_samples, dtype=np.float32)
        return AudioInput(np.asarray(source.samples, dtype=np.float32), target_sr, source.path, stereo)
    if isinstance(source, (str, Path)):
        from .io import load_audio

        data, sr, _meta = load_audio(str(source), mono=False)
        _need_same_rate(sr, target_sr)
        if data.ndim > 1:
            stereo = np.asarray(data, dtype=np.float32)
            return AudioInput(np.mean(stereo, axis=0), target_sr, str(source), stereo)
        return AudioInput(np.asarray(data, dtype=np.float32), target_sr, str(source), None)
    if isinstance(source, np.ndarray):
        samples, stereo = _split(source, mono)
        return AudioInput(samples, target_sr, None, stereo)
    if isinstance(source, tuple) and len(source) == 2:
        data, sr = source
        _need_same_rate(int(sr), target_sr)
        samples, stereo = _split(np.asarray(list(data), dtype=np.float32), mono)
        return AudioInput(samples, target_sr, None, stereo)
    raise TypeError(f"Unsupported audio source type: {type(source)!r}")
