"""Structural segmentation (mirror of the reference's ``analysis/structure.py``).

Device side (csrc/): the magnitude spectrogram and mel power spectrogram (structure.py:48-59), the
spectral flux on linear mel power in float64 (structure.py:195, a reference quirk: SURVEY Appendix B)
and the per-frame sums of the harmonic and percussive HPSS components (structure.py:52 as consumed at
:143-144 and :212-213; csrc/hpss.cu).  Host side, on those (13 .. 128) x T arrays and T-length curves:
the MFCC self-similarity, the novelty mix, peak picking, boundary refinement, beat snapping, labelling
and classification of structure.py:61-342, restated here with the same decisions.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np
import scipy.fft
import scipy.ndimage

from .. import hostlogic, runtime
from ..utils import AudioInput, seed_everything
from .beats import BeatAnalysis


@dataclass(slots=True)
class StructuralSegment:
    label: str
    category: str
    start: float
    end: float
    confidence: float
    percussive_energy: float
    harmonic_energy: float
    percussive_ratio: float


@dataclass(slots=True)
class StructureAnalysis:
    segments: List[StructuralSegment]
    novelty_curve: List[float]


@dataclass(slots=True)
class StructureFrontend:
    magnitude: np.ndarray | None  # (1 + n_fft/2, T) float32  -- structure.py:48-51 (None: left on the device)
    mel: np.ndarray | None        # (128, T) float32 power    -- structure.py:53-59 (None: left on the device)
    log_mel: np.ndarray | None    # (128, T) float64          -- structure.py:192
    spectral_flux: np.ndarray     # (T,) float64              -- structure.py:193-194
    harmonic_curve: np.ndarray | None = None    # (T,) sum over bins of hpss(magnitude)[0]
    percussive_curve: np.ndarray | None = None  # (T,) sum over bins of hpss(magnitude)[1]
    mfcc: np.ndarray | None = None              # (13, T) float64 mfcc(S=log_mel) from csrc/mfcc.cu -- structure.py:199
    self_similarity: np.ndarray | None = None   # (T,) float64 MFCC self-similarity from csrc/novelty.cu -- structure.py:199-210


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: float = 80.0) -> np.ndarray:
    """librosa.power_to_db(ref=1.0) on the host for the (128, T) log-mel the MFCC stage consumes."""
    S = np.asarray(S)
    out = 10.0 * np.log10(np.maximum(amin, S))
    return np.maximum(out, out.max() - top_db) if out.size else out


def structure_frontend(audio: AudioInput, *, frame_length: int = 2048, hop_length: int = 512,
                       magnitude: bool = True, matrices: bool = True) -> StructureFrontend:
    """Device outputs the structure stage consumes.  ``matrices=False`` (what analyse_structure uses) leaves the
    (128, T) mel matrix in HBM as well and brings back its 13-row cepstrum instead."""
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_structure expects an AudioInput instance")
    # matrices=False: the self-similarity curve is formed on the device too, so not even the cepstrum travels
    outs = ("flux_linear", "hpss_harmonic", "hpss_percussive") + (("mfcc", "mel") if matrices else ("self_similarity",)) + \
           (("magnitude",) if magnitude else ())
    res = runtime.frontend(np.asarray(audio.samples, dtype=np.float32), audio.sample_rate, n_fft=frame_length,
                           hop=hop_length, outputs=outs)
    if not matrices:
        return StructureFrontend(magnitude=res["magnitude"] if (magnitude and "magnitude" in res) else None, mel=None,
                                 log_mel=None, spectral_flux=np.asarray(res["flux_linear"], dtype=float),
                                 harmonic_curve=np.asarray(res["hpss_harmonic"]), percussive_curve=np.asarray(res["hpss_percussive"]),
                                 self_similarity=np.asarray(res["self_similarity"], dtype=float))
    mel64 = np.asarray(res["mel"], dtype=float)
    # inside a session the fused run holds every output; the 64 MB magnitude is only copied back when asked for
    return StructureFrontend(magnitude=res["magnitude"] if (magnitude and "magnitude" in res) else None, mel=res["mel"],
                             log_mel=power_to_db(mel64 + 1e-9), spectral_flux=np.asarray(res["flux_linear"], dtype=float),
                             harmonic_curve=np.asarray(res["hpss_harmonic"]), percussive_curve=np.asarray(res["hpss_percussive"]),
                             mfcc=np.asarray(res["mfcc"]))


# ------------------------------------------------------------------------------ host logic (structure.py:61-342)
def _unit_range(curve: np.ndarray) -> np.ndarray:
    if curve.size == 0:
        return curve
    lo, hi = float(np.min(curve)), float(np.max(curve))
    return np.zeros_like(curve) if hi - lo < 1e-9 else (curve - lo) / (hi - lo)


def novelty_curves(log_mel: np.ndarray | None, spectral_flux: np.ndarray, percussive_curve: np.ndarray,
                   harmonic_curve: np.ndarray, *, hop_length: int, sample_rate: int, context_seconds: float = 2.0,
                   mfcc: np.ndarray | None = None, self_similarity: np.ndarray | None = None) -> Tuple[np.ndarray, np.ndarray]:
    """(smoothed combined novelty, normalised energy novelty): structure.py:182-224 from the device curves.

    ``self_similarity``: the MFCC self-similarity curve when the device formed it (csrc/novelty.cu); else it is derived
    here from ``mfcc`` (the (13, T) cepstrum, device or host), itself derived from ``log_mel`` when absent."""
    if self_similarity is not None:
        return _mix_novelty(np.asarray(self_similarity, dtype=float), spectral_flux, percussive_curve, harmonic_curve,
                            hop_length=hop_length, sample_rate=sample_rate)
    if mfcc is None:
        # librosa.feature.mfcc(S=log_mel, n_mfcc=13): orthonormal DCT-II along the mel axis, first 13 rows
        mfcc = scipy.fft.dct(np.asarray(log_mel, dtype=float), axis=0, type=2, norm="ortho")[:13]
    frames = mfcc.shape[1]
    mfcc = scipy.ndimage.gaussian_filter1d(mfcc, sigma=1.0, axis=1)
    context = max(2, int(round(context_seconds * sample_rate / float(hop_length))))
    self_similarity = np.zeros(frames, dtype=float)
    if frames > 2 * context:
        # means of every `context`-frame window at once (the reference loops over frames, structure.py:203-210);
        # window j covers frames [j, j + context): left of frame f is window f - context, right is window f
        means = np.lib.stride_tricks.sliding_window_view(mfcc, context, axis=1).mean(axis=2)
        unit = means / (np.linalg.norm(means, axis=0) + 1e-9)
        f = np.arange(context, frames - context)
        self_similarity[f] = 1.0 - np.sum(unit[:, f - context] * unit[:, f], axis=0)
    return _mix_novelty(self_similarity, spectral_flux, percussive_curve, harmonic_curve, hop_length=hop_length,
                        sample_rate=sample_rate)


def _mix_novelty(self_similarity: np.ndarray, spectral_flux: np.ndarray, percussive_curve: np.ndarray,
                 harmonic_curve: np.ndarray, *, hop_length: int, sample_rate: int) -> Tuple[np.ndarray, np.ndarray]:
    """structure.py:211-224: percussive-ratio energy novelty and the 0.5 / 0.3 / 0.2 mix, smoothed."""
    frames = self_similarity.shape[0]
    perc = np.asarray(percussive_curve) if np.size(percussive_curve) else np.zeros(frames)
    harm = np.asarray(harmonic_curve) if np.size(harmonic_curve) else np.zeros(frames)
    ratio = perc / (perc + harm + 1e-9)
    ratio = scipy.ndimage.gaussian_filter1d(ratio, sigma=max(1.0, 0.5 * sample_rate / float(hop_length)))
    energy_novelty = _unit_range(np.abs(np.diff(ratio, prepend=ratio[0])))
    combined = 0.5 * _unit_range(np.asarray(spectral_flux, dtype=float)) + 0.3 * _unit_range(self_similarity) + 0.2 * energy_novelty
    return scipy.ndimage.gaussian_filter1d(combined, sigma=1.5), energy_novelty


def _refine(peaks: np.ndarray, energy_novelty: np.ndarray, radius: int) -> np.ndarray:
    radius = max(1, radius)
    out = []
    for idx in (int(v) for v in peaks):
        lo, hi = max(0, idx - radius), min(energy_novelty.shape[0], idx + radius + 1)
        out.append(idx if hi <= lo else lo + int(np.argmax(energy_novelty[lo:hi])))
    return np.asarray(out, dtype=int)


def _space_frames(peaks: np.ndarray, novelty: np.ndarray, min_spacing: int) -> np.ndarray:
    kept: List[int] = []
    for idx in (int(v) for v in np.sort(peaks)):
        if kept and idx - kept[-1] < min_spacing:
            if novelty[idx] > novelty[kept[-1]]:
                kept[-1] = idx
        else:
            kept.append(idx)
    return np.asarray(kept, dtype=int)


def _space_times(times: Sequence[float], frames: Sequence[int], novelty: np.ndarray, min_seconds: float) -> np.ndarray:
    times, frames = np.asarray(times, dtype=float), np.asarray(frames, dtype=int)
    if times.size <= 2:
        return np.ones(times.shape, dtype=bool)
    kept = [0]
    for idx in range(1, len(times) - 1):
        prev = kept[-1]
        if times[idx] - times[prev] < min_seconds:
            if prev != 0 and novelty[frames[idx]] > novelty[frames[prev]]:
                kept[-1] = idx
        else:
            kept.append(idx)
    kept.append(len(times) - 1)
    mask = np.zeros(times.shape, dtype=bool)
    mask[kept] = True
    return mask


def _classify(ratios: Sequence[float], perc: Sequence[float], harm: Sequence[float]) -> List[str]:
    ratios = np.asarray(ratios, dtype=float)
    total = np.asarray(perc, dtype=float) + np.asarray(harm, dtype=float)
    if total.size == 0:
        return []
    median = float(np.median(total))
    out = []
    for i, (ratio, energy) in enumerate(zip(ratios, total)):
        if i == 0:
            out.append("intro")
        elif i == len(ratios) - 1:
            out.append("outro")
        elif energy < 0.5 * median and ratio < 0.35:
            out.append("breakdown")
        elif ratio > 0.65 and energy >= 0.75 * median:
            out.append("drop")
        elif ratio > 0.45:
            out.append("groove")
        elif ratio < 0.35:
            out.append("breakdown")
        else:
            out.append("bridge")
    return out


def boundaries_from_curves(novelty: np.ndarray, energy_novelty: np.ndarray, beat_times: Sequence[float], *, sample_rate: int,
                           hop_length: int) -> Tuple[np.ndarray, np.ndarray]:
    """Section boundary frames and times: structure.py:86-130 (peak pick, refine, spacing, beat snapping)."""
    fps = sample_rate / float(hop_length)
    min_seconds = 8.0
    min_frames = max(1, int(round(min_seconds * fps)))
    peaks = hostlogic.peak_pick(novelty, pre_max=8, post_max=8, pre_avg=32, post_avg=32, delta=np.std(novelty) * 0.4,
                                wait=min_frames)
    peaks = _refine(peaks, energy_novelty, int(round(fps * 3.0))) if peaks.size else peaks
    peaks = _space_frames(peaks, novelty, min_frames) if peaks.size else peaks
    frames = np.asarray(np.unique(np.concatenate(([0], peaks, [len(novelty) - 1]))), dtype=int)
    times = hostlogic.frames_to_time(frames, sample_rate, hop_length)
    if len(beat_times):
        beats = np.asarray(beat_times)
        times = np.maximum.accumulate(np.asarray([float(beats[int(np.argmin(np.abs(beats - t)))]) for t in times]))
    keep = _space_times(times, frames, novelty, min_seconds)
    return frames[keep], np.asarray(times)[keep]


def segments_from_curves(frontend: StructureFrontend, beat_result: BeatAnalysis, *, sample_rate: int, hop_length: int,
                         duration: float) -> StructureAnalysis:
    probe = next(a for a in (frontend.mel, frontend.mfcc, frontend.self_similarity) if a is not None)
    if probe.size == 0:
        raise ValueError("not enough values to unpack (expected 2, got 0)")  # the reference's behaviour on empty audio
    # the cepstrum computed next to the mel matrix on the device is used when the matrix itself stayed there
    novelty, energy_novelty = novelty_curves(frontend.log_mel, frontend.spectral_flux, frontend.percussive_curve,
                                             frontend.harmonic_curve, hop_length=hop_length, sample_rate=sample_rate,
                                             mfcc=frontend.mfcc if frontend.log_mel is None else None,
                                             self_similarity=frontend.self_similarity)
    frames, times = boundaries_from_curves(novelty, energy_novelty, beat_result.beat_times, sample_rate=sample_rate,
                                           hop_length=hop_length)
    labels = [chr(ord("A") + i % 26) for i in range(len(frames) - 1)]
    perc_e, harm_e, ratios, segments = [], [], [], []
    peak = float(np.max(novelty)) + 1e-9
    for i, start in enumerate(frames[:-1]):
        end = frames[i + 1]
        window = novelty[start:end]
        pe = float(np.sum(frontend.percussive_curve[start:end], dtype=np.float64))
        he = float(np.sum(frontend.harmonic_curve[start:end], dtype=np.float64))
        perc_e.append(pe)
        harm_e.append(he)
        ratios.append(float(pe / (pe + he + 1e-9)))
        segments.append(StructuralSegment(
            label=labels[i], category="", start=float(times[i]), end=float(times[i + 1]),
            confidence=float(np.clip((float(np.mean(window)) if window.size else 0.0) / peak, 0.0, 1.0)),
            percussive_energy=pe, harmonic_energy=he, percussive_ratio=ratios[-1]))
    for seg, cat in zip(segments, _classify(ratios, perc_e, harm_e)):
        seg.category = cat
    return StructureAnalysis(segments=segments, novelty_curve=novelty.tolist())


def analyse_structure(audio: AudioInput | str, beat_result: BeatAnalysis, *, seed: int, frame_length: int = 2048,
                      hop_length: int = 512) -> StructureAnalysis:
    """Detect structural boundaries (reference signature: structure.py:34-41)."""
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_structure expects an AudioInput instance")
    seed_everything(seed)
    fe = structure_frontend(audio, frame_length=frame_length, hop_length=hop_length, magnitude=False, matrices=False)
    return segments_from_curves(fe, beat_result, sample_rate=audio.sample_rate, hop_length=hop_length,
                                duration=float(audio.duration))
