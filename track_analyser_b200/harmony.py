"""Harmony analysis (mirror of the reference's ``harmony.py``), frontend half on the GPU.

On the section-8a path: ``_spectral_balance`` (harmony.py:253-267; a 4096/1024 STFT reduced to
three band ratios) and the ``chroma_stft`` projection (harmony.py:108,149), plus the stereo image
(harmony.py:270-282) from the time-domain moments.  Key scoring, chord hints, change points and the
seeded MIDI suggestions are small host-side decisions on (12, T) chroma, restated from
harmony.py:190-465 so that ``analyse_harmony`` / ``key_estimate`` keep the reference's signatures
and dataclasses.

``librosa.feature.chroma_cqt`` (harmony.py:107,148) -- the input of the key scores, chord
hints, change points and MIDI suggestions -- runs on the device too (``csrc/cqt.cu``: tuning estimate,
2:1 decimation chain, per-octave transforms against the sparsified wavelet basis, chroma fold).  Its octave
decimator is a STATED stage: librosa calls libsoxr (``soxr_hq``), which is not on this machine; the kernel
uses a Kaiser-windowed sinc built from soxr HQ's published band edges ("chroma_cqt"), so this
row is parity-unpinned against the real librosa like the rest of the oracle.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import runtime
from .analysis.beats import BeatAnalysis, DownbeatAnalysis
from .utils import AudioInput, deterministic_rng, seed_everything

MAJOR_PROFILE = np.array([6.35, 2.23, 3.48, 2.33, 4.38, 4.09, 2.52, 5.19, 2.39, 3.66, 2.29, 2.88])
MINOR_PROFILE = np.array([6.33, 2.68, 3.52, 5.38, 2.6, 3.53, 2.54, 4.75, 3.98, 2.69, 3.34, 3.17])
PITCH_CLASS_NAMES = ["C", "C#", "D", "Eb", "E", "F", "F#", "G", "Ab", "A", "Bb", "B"]


@dataclass(slots=True)
class SpectralBalance:
    low_band: float
    mid_band: float
    high_band: float


@dataclass(slots=True)
class StereoImage:
    correlation: float
    balance: float


@dataclass(slots=True)
class KeyEstimate:
    key: str
    confidence: float


@dataclass(slots=True)
class KeyEstimation:
    best: KeyEstimate
    second_best: KeyEstimate


@dataclass(slots=True)
class ChordHint:
    time: float
    chord: str
    confidence: float


@dataclass(slots=True)
class ChordChangePoint:
    time: float
    strength: float


@dataclass(slots=True)
class MidiSuggestion:
    name: str
    notes: object  # pandas.DataFrame with columns start, duration, pitch, velocity, channel


@dataclass(slots=True)
class HarmonyAnalysis:
    spectral_balance: SpectralBalance
    stereo_image: StereoImage
    primary_key: KeyEstimate
    secondary_key: KeyEstimate
    chord_hints: List[ChordHint]
    chord_change_points: List[ChordChangePoint]
    hook_suggestion: MidiSuggestion
    bass_suggestion: MidiSuggestion

    @property
    def key_estimate(self) -> KeyEstimate:
        return self.primary_key


@dataclass(slots=True)
class HarmonyFrontend:
    """GPU outputs of the harmony stage for one track."""

    spectral_balance: SpectralBalance
    chroma_stft: np.ndarray
    tuning: float


def chroma_stft(y: np.ndarray, sr: int) -> Tuple[np.ndarray, float]:
    """librosa.feature.chroma_stft(y=y, sr=sr) -> ((12, T) float32, tuning)."""
    res = runtime.frontend(np.asarray(y, dtype=np.float32), sr, outputs=("chroma", "tuning"))
    return res["chroma"], res["tuning"]


def _spectral_balance(audio: AudioInput) -> SpectralBalance:
    sr = audio.sample_rate
    res = runtime.frontend(np.asarray(audio.samples, dtype=np.float32), sr, n_fft=4096, hop=1024, n_mels=0,
                           outputs=("ltas",))
    sums = res["ltas"].astype(np.float64) * res.n_frames  # per-bin time sums of |X|
    total = float(np.sum(sums))
    if total <= 0:
        return SpectralBalance(0.0, 0.0, 0.0)
    freqs = np.fft.rfftfreq(n=4096, d=1.0 / sr)
    lo, mid, hi = freqs < 200, (freqs >= 200) & (freqs < 2000), freqs >= 2000
    return SpectralBalance(float(sums[lo].sum() / total), float(sums[mid].sum() / total), float(sums[hi].sum() / total))


def _stereo_image(audio: AudioInput) -> StereoImage:
    """np.corrcoef(L, R) and mean|L| - mean|R| (harmony.py:270-282) from the device's time-domain moments."""
    samples = audio.stereo_samples if audio.stereo_samples is not None else audio.samples
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim == 1 or samples.shape[0] < 2:
        return StereoImage(correlation=1.0, balance=0.0)
    if samples.shape[1] == 0:
        return StereoImage(correlation=0.0, balance=0.0)
    m = runtime.frontend(np.ascontiguousarray(samples[:2]), audio.sample_rate, outputs=("moments",))["moments"]
    n = float(m[7])
    cov = m[4] - m[0] * m[1] / n
    # one-pass moments can come out slightly negative for a (near-)constant channel: clamp, the ratio is nan then anyway
    var_l, var_r = max(m[2] - m[0] * m[0] / n, 0.0), max(m[3] - m[1] * m[1] / n, 0.0)
    denom = np.sqrt(var_l) * np.sqrt(var_r)
    corr = float(cov / denom) if denom > 0 else float("nan")  # np.corrcoef of a constant channel is nan as well
    return StereoImage(correlation=corr, balance=float(m[8] / n - m[9] / n))


def _key_names() -> List[str]:
    return [f"{p} major" for p in PITCH_CLASS_NAMES] + [f"{p} minor" for p in PITCH_CLASS_NAMES]


def _score_keys(chroma_matrices: Sequence[np.ndarray]) -> Tuple[np.ndarray, List[str]]:
    if not chroma_matrices:
        return np.array([]), []
    major = MAJOR_PROFILE / np.linalg.norm(MAJOR_PROFILE)
    minor = MINOR_PROFILE / np.linalg.norm(MINOR_PROFILE)
    total = np.zeros(24, dtype=float)
    for chroma in chroma_matrices:
        if chroma.size == 0:
            continue
        mean = np.mean(chroma, axis=1)
        nrm = np.linalg.norm(mean)
        if nrm <= 0:
            continue
        mean = mean / nrm
        total[:12] += [float(np.dot(mean, np.roll(major, s))) for s in range(12)]
        total[12:] += [float(np.dot(mean, np.roll(minor, s))) for s in range(12)]
    return total, _key_names()


def _rank_keys(scores: np.ndarray, keys: List[str]) -> KeyEstimation:
    if not scores.size:
        fallback = KeyEstimate(key="C major", confidence=0.0)
        return KeyEstimation(best=fallback, second_best=fallback)
    pos = np.maximum(scores, 0.0)
    conf = pos / (float(np.sum(pos)) or 1.0)
    first = int(np.argmax(conf))
    best = KeyEstimate(key=keys[first], confidence=float(conf[first]))
    conf[first] = -np.inf
    second = int(np.argmax(conf))
    return KeyEstimation(best=best, second_best=KeyEstimate(key=keys[second], confidence=float(max(conf[second], 0.0))))


def key_estimate(y: np.ndarray, sr: int) -> KeyEstimation:
    """Best and second-best key (harmony.py:99-129); the chroma_cqt slot is filled by ``_chroma_cqt``."""
    with runtime.frontend_session():
        return _rank_keys(*_score_keys([_chroma_cqt(y, sr), chroma_stft(y, sr)[0]]))


def key_index(estimate: KeyEstimation) -> int:
    """Integer index (0..23) of the best key: the 'key index' integer output of the north star."""
    return _key_names().index(estimate.best.key)


def harmony_frontend(audio: AudioInput) -> HarmonyFrontend:
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_harmony expects an AudioInput instance")
    chroma, tuning = chroma_stft(audio.samples, audio.sample_rate)
    return HarmonyFrontend(spectral_balance=_spectral_balance(audio), chroma_stft=chroma, tuning=tuning)


def _chroma_cqt(y: np.ndarray, sr: int) -> np.ndarray:
    """librosa.feature.chroma_cqt(y=y, sr=sr) -> (12, T) float32 (harmony.py:107,148), by the constant-Q kernels of
    ``csrc/cqt.cu``.  Raises like librosa does when the basis does not fit under the Nyquist frequency."""
    if runtime.is_precomputed():
        return runtime.frontend(np.asarray(y, dtype=np.float32), sr, outputs=("chroma_cqt",))["chroma_cqt"]
    plan = runtime.get_plan(sr)
    if not plan.cqt_ok:
        try:
            plan.cqt_frame_count(0)
        except RuntimeError as exc:   # librosa: ParameterError (wavelet basis exceeds the Nyquist frequency)
            raise ValueError(str(exc)) from exc
    return runtime.frontend(np.asarray(y, dtype=np.float32), sr, outputs=("chroma_cqt",))["chroma_cqt"]


def _beat_profiles(chroma: np.ndarray, beat_result: BeatAnalysis):
    """L2-normalised mean chroma of the 4 frames around each beat (harmony.py:295-304, 354-363)."""
    out = []
    for idx, frame in enumerate(beat_result.beat_frames):
        window = chroma[:, max(0, frame - 2): frame + 2]
        if window.size == 0:
            continue
        mean = np.mean(window, axis=1)
        nrm = np.linalg.norm(mean)
        if nrm > 0:
            out.append((idx, mean / nrm))
    return out


def _chord_templates() -> Dict[str, np.ndarray]:
    shapes = {"maj": (0, 4, 7), "min": (0, 3, 7), "dim": (0, 3, 6), "sus2": (0, 2, 7), "sus4": (0, 5, 7)}
    out: Dict[str, np.ndarray] = {}
    for root, name in enumerate(PITCH_CLASS_NAMES):
        for quality, steps in shapes.items():
            t = np.zeros(12)
            t[[(root + k) % 12 for k in steps]] = 1.0
            out[f"{name}{quality}"] = t / np.linalg.norm(t)
    return out


def _estimate_chords(chroma: np.ndarray, beat_result: BeatAnalysis, rng: np.random.Generator, profiles=None) -> List[ChordHint]:
    if not beat_result.beat_frames:
        return []
    names, mats = zip(*_chord_templates().items())
    mats = np.stack(mats)
    prof = _beat_profiles(chroma, beat_result) if profiles is None else profiles
    if not prof:
        return []
    # the reference scores one template at a time with np.dot (harmony.py:305-312); one matrix product gives the same
    # 60 numbers per beat, and is used only if it reproduces np.dot bit for bit on this BLAS (checked on two beats)
    P = np.stack([p for _, p in prof])
    S = P @ mats.T
    for r in {0, len(prof) - 1}:
        if not np.array_equal(S[r], np.array([float(np.dot(t, P[r])) for t in mats])):
            S = np.array([[float(np.dot(t, p)) for t in mats] for p in P])
            break
    noise = rng.normal(0.0, 1e-6, size=S.shape)  # row i is the reference's i-th draw (seeded tie-breaker)
    best = np.argmax(S + noise, axis=1)
    top = np.max(S + 1e-9, axis=1)
    return [ChordHint(time=float(beat_result.beat_times[idx]), chord=names[int(b)], confidence=float(S[r, b] / float(top[r])))
            for r, ((idx, _), b) in enumerate(zip(prof, best))]


def _detect_chord_changes(chroma: np.ndarray, beat_result: BeatAnalysis, chord_hints: Sequence[ChordHint],
                          profiles=None) -> List[ChordChangePoint]:
    if len(beat_result.beat_frames) < 2:
        return []
    prof = _beat_profiles(chroma, beat_result) if profiles is None else profiles
    if len(prof) < 2:
        return []
    times = [float(beat_result.beat_times[i]) for i, _ in prof]
    strengths = [float(np.clip(1.0 - float(np.clip(np.dot(a[1], b[1]), -1.0, 1.0)), 0.0, 1.0)) for a, b in zip(prof, prof[1:])]
    arr = np.asarray(strengths)
    keep = max(1, int(np.ceil(arr.size * 0.9)))
    threshold = float(np.min(arr)) if keep >= arr.size else float(np.partition(arr, arr.size - keep)[arr.size - keep])
    threshold = max(threshold, 0.15)
    found: Dict[float, float] = {}
    for t, v in zip(times[1:], strengths):
        if v >= threshold:
            found[t] = max(found.get(t, 0.0), v)
    found[times[1]] = max(found.get(times[1], 0.0), strengths[0])
    if len(chord_hints) >= 2:
        templates = _chord_templates()
        for prev, cur in zip(chord_hints, chord_hints[1:]):
            if cur.chord == prev.chord:
                continue
            a, b = templates.get(prev.chord), templates.get(cur.chord)
            sim = 0.0 if a is None or b is None else float(np.clip(np.dot(a, b), -1.0, 1.0))
            found[cur.time] = max(found.get(cur.time, 0.0), float(np.clip(1.0 - sim, 0.0, 1.0)))
    if not found:
        return []
    top = max(found.values()) or 1.0
    return [ChordChangePoint(time=float(t), strength=float(v / top)) for t, v in sorted(found.items())]


def _scale_for_key(key: str) -> List[int]:
    root, _, mode = key.partition(" ")
    steps = (0, 2, 4, 5, 7, 9, 11) if mode.strip().lower().startswith("major") else (0, 2, 3, 5, 7, 8, 10)
    return [(PITCH_CLASS_NAMES.index(root) + k) % 12 for k in steps]


def _generate_midi(chroma: np.ndarray, beat_result: BeatAnalysis, key: KeyEstimate, rng: np.random.Generator, *, name: str,
                   octave: int = 0, start_offset: float = 0.0) -> MidiSuggestion:
    import pandas as pd

    scale = _scale_for_key(key.key)
    beats_ = [max(0.0, b - start_offset) for b in beat_result.beat_times[:8]] or [0.0, 0.5, 1.0, 1.5]
    duration = np.median(np.diff(beats_)) if len(beats_) > 1 else 0.5
    rows = []
    for t in beats_:  # two draws per note, in the reference's order: scale degree, then velocity offset
        degree = int(scale[int(rng.integers(0, len(scale)))])
        velocity = int(np.clip(96 + rng.integers(-12, 12), 20, 127))
        rows.append({"start": float(t), "duration": float(duration), "pitch": int(60 + degree + octave * 12),
                     "velocity": velocity, "channel": 0})
    return MidiSuggestion(name=name, notes=pd.DataFrame(rows, columns=["start", "duration", "pitch", "velocity", "channel"]))


def analyse_harmony(audio: AudioInput, beat_result: BeatAnalysis, downbeat_result: Optional[DownbeatAnalysis], *,
                    seed: int) -> HarmonyAnalysis:
    """Reference signature (harmony.py:132-139); see the module docstring for the chroma_cqt deviation."""
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_harmony expects an AudioInput instance")
    seed_everything(seed)
    rng = deterministic_rng(seed)
    with runtime.frontend_session():
        balance = _spectral_balance(audio)
        image = _stereo_image(audio)
        cqt_like = _chroma_cqt(audio.samples, audio.sample_rate)
        stft_chroma = chroma_stft(audio.samples, audio.sample_rate)[0]
    keys = _rank_keys(*_score_keys([cqt_like, stft_chroma]))
    profiles = _beat_profiles(cqt_like, beat_result)  # shared by the two consumers (the reference builds them twice)
    hints = _estimate_chords(cqt_like, beat_result, rng, profiles)
    changes = _detect_chord_changes(cqt_like, beat_result, hints, profiles)
    if downbeat_result and downbeat_result.downbeat_times:
        offset = downbeat_result.downbeat_times[0]
    else:
        offset = beat_result.beat_times[0] if beat_result.beat_times else 0.0
    hook = _generate_midi(cqt_like, beat_result, keys.best, rng, name="hook", start_offset=offset)
    bass = _generate_midi(cqt_like, beat_result, keys.best, rng, name="bass", octave=-1, start_offset=offset)
    return HarmonyAnalysis(spectral_balance=balance, stereo_image=image, primary_key=keys.best, secondary_key=keys.second_best,
                           chord_hints=hints, chord_change_points=changes, hook_suggestion=hook, bass_suggestion=bass)


__all__ = ["HarmonyAnalysis", "ChordChangePoint", "ChordHint", "KeyEstimation", "KeyEstimate", "MidiSuggestion",
           "SpectralBalance", "StereoImage", "analyse_harmony", "key_estimate"]
