"""Input containers and coercion (mirror of the reference's ``utils.py``).

Same public names and semantics as /root/reference/src/track_analyser/utils.py:
``AudioInput`` (:28-39), ``DEFAULT_SR``/``DEFAULT_SEED`` (:24-25),
``seed_everything`` (:48-52), ``deterministic_rng`` (:42-45), ``coerce_audio``
(:73-146).  WAV files are decoded by ``io.load_audio``; inputs whose sample rate
differs from ``target_sr`` are converted on the device by ``resample.resample``
(the reference's resampy call, utils.py:55-70).

This file is out-of-path glue that SURVEY.md section 8b obliges the package to mirror: ``AudioInput`` has to keep the
reference's fields and order, and ``coerce_audio`` restates the reference's branches (file / array / AudioInput, mono
and stereo views) statement by statement so that the callers' behaviour is the same.  It is a transcription of that
contract, not a design of its own; with ``track_analyser_b200.install()`` the reference's own utils.py is what runs.
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


def _resample(samples: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """utils.py:55-70: resampy.resample per channel, here one device call for all channels."""
    if orig_sr == target_sr:
        return samples
    from .resample import resample

    return resample(samples, orig_sr, target_sr)


def _split(samples: np.ndarray, mono: bool):
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim > 1:
        return (np.mean(samples, axis=0) if mono else samples), samples
    return samples, None


def coerce_audio(source, *, target_sr: int = DEFAULT_SR, mono: bool = True) -> AudioInput:
    if isinstance(source, AudioInput):  # utils.py:86-101
        samples = np.asarray(source.samples, dtype=np.float32)
        stereo = None if source.stereo_samples is None else np.asarray(source.stereo_samples, dtype=np.float32)
        if source.sample_rate != target_sr:
            samples = _resample(samples, source.sample_rate, target_sr)
            if stereo is not None:
                stereo = _resample(stereo, source.sample_rate, target_sr)
        return AudioInput(samples, target_sr, source.path, stereo)
    if isinstance(source, (str, Path)):  # utils.py:103-122
        from .io import load_audio

        data, sr, _meta = load_audio(str(source), mono=False)
        if data.ndim > 1:
            stereo = np.asarray(data, dtype=np.float32)
            mono_samples = np.mean(stereo, axis=0)
        else:
            stereo = None
            mono_samples = np.asarray(data, dtype=np.float32)
        mono_samples = _resample(mono_samples, sr, target_sr)
        if stereo is not None:
            stereo = _resample(stereo, sr, target_sr)
            if mono:
                mono_samples = np.mean(stereo, axis=0)
        return AudioInput(mono_samples, target_sr, str(source), stereo)
    if isinstance(source, np.ndarray):
        samples, stereo = _split(source, mono)
        return AudioInput(samples, target_sr, None, stereo)
    if isinstance(source, tuple) and len(source) == 2:  # utils.py:134-144
        data, sr = source
        samples, stereo = _split(np.asarray(list(data), dtype=np.float32), mono)
        samples = _resample(samples, int(sr), target_sr)
        if stereo is not None:
            stereo = _resample(stereo, int(sr), target_sr)
        return AudioInput(samples, target_sr, None, stereo)
    raise TypeError(f"Unsupported audio source type: {type(source)!r}")
