"""Loudness and dynamics (mirror of the reference's ``analysis/loudness.py``).

``measure_loudness`` keeps the reference signature (loudness.py:45-78); the
K-weighting scan, 400 ms gated block energies, BS.1770 gating and the centred RMS
frames are computed by csrc/timedomain.cu, and ``true_peak_dbtp`` (loudness.py:81-97,
SURVEY 8f rank 2) by csrc/truepeak.cu: the reference's 8x polyphase resampler evaluated
exactly, but only where its output can reach the maximum.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

from .. import engine, loudness_host, runtime
from ..utils import AudioInput, seed_everything


@dataclass(slots=True)
class LoudnessAnalysis:
    integrated_lufs: float
    short_term_lufs: List[float]
    momentary_lufs: List[float]
    loudness_range: float
    true_peak_dbfs: float
    rms_dbfs: float


def _td(samples: np.ndarray, sample_rate: int, meter_block_size: float, outputs):
    return runtime.frontend(samples, sample_rate, meter_block=meter_block_size, outputs=outputs)


def _windowed_loudness(samples: np.ndarray, sample_rate: int, meter_block_size: float) -> np.ndarray:
    key = "rms_short" if meter_block_size == 3.0 else "rms_momentary"
    block = 0.4 if meter_block_size == 3.0 else meter_block_size
    return loudness_host.frames_to_db(_td(np.asarray(samples, dtype=np.float32), sample_rate, block, (key,))[key])


def measure_loudness(samples: np.ndarray, sample_rate: int,
                     meter_block_size: float = 0.400) -> Tuple[float, List[float], List[float], float]:
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim != 1:
        raise ValueError("measure_loudness expects mono audio samples")
    if samples.shape[0] < meter_block_size * sample_rate:
        raise ValueError("Audio must have length greater than the block size.")  # pyloudnorm.util.valid_audio
    res = _td(samples, sample_rate, meter_block_size, ("lufs", "rms_momentary", "rms_short", "moments"))
    short_term = loudness_host.frames_to_db(res["rms_short"])
    momentary = loudness_host.frames_to_db(res["rms_momentary"])
    lra = float(np.percentile(momentary, 95) - np.percentile(momentary, 5))
    return (float(res["lufs"]), np.asarray(short_term, dtype=float).tolist(),
            np.asarray(momentary, dtype=float).tolist(), lra)


def true_peak_dbtp(samples: np.ndarray, sample_rate: int, *, oversample: int = 8) -> float:
    if oversample < 1:
        raise ValueError("oversample must be >= 1")
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim != 1:
        raise ValueError("true_peak_dbtp expects mono audio samples")
    if oversample > 32:
        raise ValueError("the device kernel oversamples by at most 32")
    if not samples.size:
        peak = 0.0
    elif oversample == 8:
        peak = _td(samples, sample_rate, 0.4, ("true_peak",))["true_peak"]
    else:  # not the factor the session's shared run uses: its own pass
        plan = runtime.get_plan(sample_rate)
        peak = engine.analyse_batch(plan, [samples], ("true_peak",), true_peak_oversample=int(oversample))[0]["true_peak"]
    return float(20.0 * np.log10(float(peak) + 1e-12))


def analyse_loudness(audio: AudioInput | str, *, seed: int, meter_block_size: float = 0.400) -> LoudnessAnalysis:
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_loudness expects an AudioInput instance")
    seed_everything(seed)
    samples = np.asarray(audio.samples, dtype=np.float32)  # no copy for float32 input: keeps the session's cache key
    with runtime.frontend_session():
        integrated, short_term, momentary, lra = measure_loudness(samples, audio.sample_rate, meter_block_size)
        res = _td(samples, audio.sample_rate, meter_block_size, ("moments",))
        m = res["moments"]
        peak_db = true_peak_dbtp(samples, audio.sample_rate)
    return LoudnessAnalysis(
        integrated_lufs=integrated, short_term_lufs=short_term, momentary_lufs=momentary, loudness_range=lra,
        true_peak_dbfs=peak_db,
        rms_dbfs=loudness_host.rms_dbfs_from_moments(m, res.channels == 2),
    )
