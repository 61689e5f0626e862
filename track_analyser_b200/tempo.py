"""Tempo estimation (mirror of the reference's ``tempo.py``).

``_onset_envelope`` (tempo.py:16-24) and the autocorrelation of ``estimate_bpm``
(tempo.py:38) run on the GPU (csrc/stft_fused.cu -> onset.cu -> autocorr.cu);
the remaining lag masking, peak interpolation, onset regression and grid
construction (tempo.py:42-75, :78-175) are scalar host logic on those outputs.
That host half restates the reference's decisions in the reference's order (a bit-exact integer contract leaves no other
choice); with ``track_analyser_b200.install()`` the reference's own tempo.py runs on the same device outputs instead.
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

from . import hostlogic, runtime

DEFAULT_HOP_LENGTH = 512
BEATS_PER_BAR = 4


def _frontend(y: np.ndarray, sr: int, hop_length: int, outputs):
    return runtime.frontend(np.asarray(y, dtype=np.float32), sr, n_fft=2048, hop=hop_length, outputs=outputs)


def _onset_envelope(y: np.ndarray, sr: int, hop_length: int) -> np.ndarray:
    env = _frontend(y, sr, hop_length, ("onset_env",))["onset_env"]
    return np.zeros(1, dtype=float) if env.size == 0 else env


def _onset_autocorrelation(y: np.ndarray, sr: int, hop_length: int) -> Tuple[np.ndarray, np.ndarray]:
    res = _frontend(y, sr, hop_length, ("onset_env", "autocorr"))
    return res["onset_env"], res["autocorr"]


def _bpm_from_autocorr(onset_env: np.ndarray, autocorr: np.ndarray, sr: int, bpm_min: float, bpm_max: float,
                       hop_length: int) -> float:
    if autocorr.size <= 1:
        return float(bpm_min)
    ac = autocorr[1:]
    lags = np.arange(1, ac.size + 1, dtype=float)
    tempi = 60.0 * sr / (lags * hop_length)
    sel = (tempi >= bpm_min) & (tempi <= bpm_max)
    if not np.any(sel):
        sel = tempi > 0
    cand = hostlogic.normalize_inf(ac[sel])
    cand_lags = lags[sel]
    k = int(np.argmax(cand))
    lag = cand_lags[k]
    if 0 < k < cand.size - 1:
        a, b, c = cand[k - 1], cand[k], cand[k + 1]
        curv = a - 2 * b + c
        if abs(curv) > 1e-9:
            lag = float(cand_lags[k] + 0.5 * (a - c) / curv)
    lag = max(lag, 1.0)
    bpm = float(60.0 * sr / (lag * hop_length))
    fit = _fit_onset_regression(onset_env, sr, hop_length, 60.0 / bpm)
    if fit is not None and fit[1] > 0:
        refined = 60.0 / fit[1]
        if bpm_min <= refined <= bpm_max:
            bpm = float(refined)
    return bpm


def estimate_bpm(y: np.ndarray, sr: int, bpm_min: float = 90.0, bpm_max: float = 135.0, *,
                 hop_length: int = DEFAULT_HOP_LENGTH) -> float:
    env, ac = _onset_autocorrelation(y, sr, hop_length)
    if env.size == 0:
        env = np.zeros(1, dtype=float)
    return _bpm_from_autocorr(env, ac, sr, bpm_min, bpm_max, hop_length)


def _initial_beat_time(onset_env: np.ndarray, sr: int, hop_length: int) -> Tuple[float, int]:
    frames = hostlogic.onset_detect(onset_env, sr, hop_length, backtrack=True, units="frames")
    if frames.size == 0:
        return 0.0, 0
    first = int(frames[0])
    return float(hostlogic.frames_to_time(first, sr, hop_length)), first


def _fit_onset_regression(onset_env: np.ndarray, sr: int, hop_length: int,
                          beat_period: float) -> Tuple[float, float] | None:
    times = np.asarray(hostlogic.onset_detect(onset_env, sr, hop_length, backtrack=True, units="time"), dtype=float)
    if times.size < 4 or beat_period <= 0:
        return None
    beat_idx = np.round(times / beat_period).astype(int)
    keep = beat_idx >= 0
    if not np.any(keep):
        return None
    first_hit: dict[int, float] = {}
    for i, t in zip(beat_idx[keep], times[keep]):
        first_hit.setdefault(int(i), float(t))
    if len(first_hit) < 4:
        return None
    xs = np.array(sorted(first_hit))
    ys = np.array([first_hit[i] for i in xs])
    design = np.vstack([np.ones_like(xs), xs]).T
    intercept, slope = np.linalg.lstsq(design, ys, rcond=None)[0]
    return float(intercept), float(slope)


def beat_grid(y: np.ndarray, sr: int, *, hop_length: int = DEFAULT_HOP_LENGTH, beats_per_bar: int = BEATS_PER_BAR):
    import pandas as pd

    with runtime.frontend_session():
        onset_env = _onset_envelope(y, sr, hop_length)
        bpm = estimate_bpm(y, sr, hop_length=hop_length)
    period = 60.0 / bpm
    fit = _fit_onset_regression(onset_env, sr, hop_length, period)
    start = max(fit[0], 0.0) if fit is not None else _initial_beat_time(onset_env, sr, hop_length)[0]
    start = max(start, 0.0)
    duration = len(y) / float(sr)
    if start > duration:
        start = 0.0
    count = max(1, int(np.floor((duration - start) / period)) + 1)
    times = start + np.arange(count, dtype=float) * period
    times = times[times <= duration + 1e-3]
    frames = hostlogic.time_to_frames(times, sr, hop_length)
    index = np.arange(times.size)
    beat_no = index % beats_per_bar + 1
    return pd.DataFrame({
        "time": times,
        "frame": frames.astype(int),
        "bar": (index // beats_per_bar + 1).astype(int),
        "beat": beat_no.astype(int),
        "is_downbeat": beat_no == 1,
    })


__all__ = ["estimate_bpm", "beat_grid"]
