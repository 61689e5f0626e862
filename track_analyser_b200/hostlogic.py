"""Host-side decision logic that consumes GPU envelopes (integer outputs).

These are restatements of the small librosa 0.10.2 routines the reference's
tempo/structure code calls on the onset envelope:
``util.peak_pick``, ``onset.onset_backtrack``, ``onset.onset_detect``,
``frames_to_time`` / ``time_to_frames`` and ``util.normalize``.  They are O(T)
scalar loops over a few thousand frames and stay on the host by design (section
8a marks them out of the hot path); the parity tests feed them the oracle's and
the GPU's envelopes and require identical integers.
"""

from __future__ import annotations

import numpy as np


def tiny(x) -> float:
    x = np.asarray(x)
    dt = x.dtype if np.issubdtype(x.dtype, np.floating) else np.dtype(np.float32)
    return float(np.finfo(dt).tiny)


def normalize_inf(x: np.ndarray) -> np.ndarray:
    """librosa.util.normalize(x) for 1-d input: divide by max |x| unless it is below tiny."""
    x = np.asarray(x)
    length = float(np.max(np.abs(x).astype(float))) if x.size else 0.0
    if length < tiny(x):
        length = 1.0
    out = np.empty_like(x)
    out[:] = x / length
    return out


def frames_to_time(frames, sr: int, hop_length: int):
    return (np.asanyarray(frames) * hop_length) / float(sr)


def time_to_frames(times, sr: int, hop_length: int):
    samples = (np.asanyarray(times) * sr).astype(int)
    return np.floor(np.asanyarray(samples) // hop_length).astype(int)


def peak_pick(x: np.ndarray, pre_max: int, post_max: int, pre_avg: int, post_avg: int, delta: float,
              wait: int) -> np.ndarray:
    """Greedy local-max / above-local-mean picker (librosa 0.10.2 semantics, sequential means in x's dtype).

    The local-maximum test is evaluated for all frames at once (pure comparisons, so exactly the loop's
    `x[n] == max(window)`); the sequential-mean test and the `wait` rule then run over those candidates only.
    """
    x = np.asarray(x)
    n_total = x.shape[0]
    pre_max, post_max, pre_avg, post_avg, wait = (int(np.ceil(v)) for v in (pre_max, post_max, pre_avg, post_avg, wait))
    acc_t = x.dtype.type if np.issubdtype(x.dtype, np.floating) else np.float64

    def seq_mean(lo: int, hi: int):
        acc = acc_t(0)
        for v in x[lo:hi]:
            acc = acc_t(acc + v)
        return acc_t(acc / acc_t(hi - lo))

    peaks = []
    if n_total == 0:
        return np.asarray(peaks, dtype=int)
    first = bool(x[0] >= np.max(x[: min(post_max, n_total)]))
    first = first and bool(x[0] >= seq_mean(0, min(post_avg, n_total)) + acc_t(delta))
    n_min = 1
    if first:
        peaks.append(0)
        n_min = wait + 1
    # running maximum over [n - pre_max, n + post_max): windows truncated at the ends like the loop's slices
    pad_lo, pad_hi = pre_max, max(post_max - 1, 0)
    padded = np.concatenate([np.full(pad_lo, -np.inf, dtype=x.dtype), x, np.full(pad_hi, -np.inf, dtype=x.dtype)])
    win = np.lib.stride_tricks.sliding_window_view(padded, pad_lo + pad_hi + 1) if padded.size >= pad_lo + pad_hi + 1 else None
    local_max = win.max(axis=1) if win is not None else np.full(n_total, np.inf, dtype=x.dtype)
    cand = np.flatnonzero(x == local_max)
    cand = cand[cand >= 1]
    if cand.size == 0:
        return np.asarray(peaks, dtype=int)
    # sequential means of all candidates at once: the same left-to-right additions in x's dtype, vectorised
    # over candidates (adding nothing past the end of a truncated window)
    lo = np.maximum(0, cand - pre_avg)
    hi = np.minimum(cand + post_avg, n_total)
    acc = np.zeros(cand.size, dtype=acc_t)
    for j in range(pre_avg + post_avg):
        idx = lo + j
        ok = idx < hi
        acc = np.where(ok, (acc + x[np.minimum(idx, n_total - 1)].astype(acc_t, copy=False)).astype(acc_t, copy=False), acc)
    mean = (acc / (hi - lo).astype(acc_t)).astype(acc_t, copy=False)
    passing = cand[x[cand] >= (mean + acc_t(delta)).astype(acc_t, copy=False)]
    for n in passing:
        n = int(n)
        if n >= n_min:
            peaks.append(n)
            n_min = n + wait + 1
    return np.asarray(peaks, dtype=int)


def onset_backtrack(events: np.ndarray, energy: np.ndarray) -> np.ndarray:
    """Move each event to the closest preceding local minimum of ``energy``."""
    energy = np.asarray(energy)
    minima = np.flatnonzero((energy[1:-1] <= energy[:-2]) & (energy[1:-1] < energy[2:]))
    minima = np.concatenate(([0], 1 + minima)).astype(int)
    idx = np.searchsorted(minima, np.asarray(events), side="right") - 1
    return minima[np.maximum(idx, 0)]


def onset_detect(onset_envelope: np.ndarray, sr: int, hop_length: int = 512, backtrack: bool = False,
                 units: str = "frames") -> np.ndarray:
    env = np.asarray(onset_envelope)
    env = env - np.min(env, keepdims=True, axis=-1)
    env = env / (np.max(env, keepdims=True, axis=-1) + tiny(env))
    env = env.astype(np.asarray(onset_envelope).dtype, copy=False)
    if not env.any() or not np.all(np.isfinite(env)):
        onsets = np.array([], dtype=int)
    else:
        onsets = peak_pick(env, pre_max=0.03 * sr // hop_length, post_max=0.00 * sr // hop_length + 1,
                           pre_avg=0.10 * sr // hop_length, post_avg=0.10 * sr // hop_length + 1,
                           delta=0.07, wait=0.03 * sr // hop_length)
        if backtrack and onsets.size:
            onsets = onset_backtrack(onsets, env)
    if units == "frames":
        return onsets
    if units == "samples":
        return (np.asanyarray(onsets) * hop_length).astype(int)
    if units == "time":
        return frames_to_time(onsets, sr, hop_length)
    raise ValueError(f"Invalid unit type: {units}")
