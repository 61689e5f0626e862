"""numpy restatement of the librosa 0.10.2.post1 routines the reference calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  librosa is an un-vendored
third-party dependency of the reference (pin: /root/reference/pyproject.toml:15);
its source is not on this machine, so each function restates the published
algorithm and names the reference call site that reaches
it.  dtype notes follow numpy 1.26.4 semantics (the reference's pin): real FFTs
run in float64 regardless of the input dtype, so inputs are promoted explicitly
here because numpy >= 2 would otherwise transform float32 in float32.
"""

from __future__ import annotations

import numpy as np
import scipy.fft
import scipy.signal

# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------


def tiny(x) -> float:
    """librosa.util.tiny: smallest normal number of x's (real) float type."""
    x = np.asarray(x)
    if np.issubdtype(x.dtype, np.floating) or np.issubdtype(x.dtype, np.complexfloating):
        dtype = x.dtype
    else:
        dtype = np.dtype(np.float32)
    return float(np.finfo(dtype).tiny)


def normalize(S, norm=np.inf, axis=0):
    """librosa.util.normalize(S, norm, axis, threshold=None, fill=None).

    Used by estimate_bpm (tempo.py:50, norm=inf on a 1-d slice), spectral_centroid
    (norm=1), chroma_stft / tempogram (norm=inf, axis=-2), filters.chroma (norm=2).
    """
    S = np.asarray(S)
    threshold = tiny(S)
    mag = np.abs(S).astype(float)
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == 1:
        length = np.sum(mag, axis=axis, keepdims=True)
    elif norm == 2:
        length = np.sum(mag**2, axis=axis, keepdims=True) ** 0.5
    else:  # pragma: no cover - not reached by the reference
        raise ValueError(norm)
    small = length < threshold
    length[small] = 1.0
    out = np.empty_like(S)
    out[:] = S / length
    return out


def frame(x, frame_length, hop_length):
    """librosa.util.frame along the last axis -> (..., frame_length, n_frames) view."""
    x = np.asarray(x)
    n = x.shape[-1]
    if n < frame_length:
        raise ValueError(f"Input is too short (n={n}) for frame_length={frame_length}")
    win = np.lib.stride_tricks.sliding_window_view(x, frame_length, axis=-1)
    win = win[..., ::hop_length, :]
    return np.moveaxis(win, -1, -2)


def fft_frequencies(sr, n_fft):
    """librosa.fft_frequencies (features.py:81, stereo.py:100, harmony.py:255)."""
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def get_window(name, n):
    """scipy.signal.get_window(name, n, fftbins=True): periodic window, float64."""
    return scipy.signal.get_window(name, n, fftbins=True)


# --------------------------------------------------------------------------
# A.1  STFT
# --------------------------------------------------------------------------


def stft(y, n_fft=2048, hop_length=None, window="hann"):
    """librosa.stft(center=True, pad_mode="constant", win_length=n_fft).

    Reference call sites: features.py:79, stereo.py:95-96, structure.py:48,
    harmony.py:254, and inside melspectrogram / spectral_centroid /
    spectral_rolloff / chroma_stft.  Output: (1 + n_fft/2, 1 + N//hop),
    complex64 for float32 input (complex128 for float64 input), Fortran order
    as librosa allocates it.
    """
    y = np.asarray(y)
    if hop_length is None:
        hop_length = n_fft // 4
    w = get_window(window, n_fft)  # float64
    ypad = np.pad(y, (n_fft // 2, n_fft // 2), mode="constant")
    frames = frame(ypad, n_fft, hop_length)  # (n_fft, T) view
    out_dtype = np.complex64 if y.dtype == np.float32 else np.complex128
    T = frames.shape[-1]
    out = np.zeros((1 + n_fft // 2, T), dtype=out_dtype, order="F")
    # column blocks only bound memory; they have no numeric effect
    step = max(1, (2**25) // (n_fft * 8))
    for s in range(0, T, step):
        blk = frames[:, s : s + step].astype(np.float64) * w[:, None]
        out[:, s : s + step] = scipy.fft.rfft(blk, axis=0)
    return out


def spectrogram(y, n_fft, hop_length, power):
    """librosa.core.spectrum._spectrogram: |stft| ** power, in y's real dtype."""
    S = np.abs(stft(y, n_fft=n_fft, hop_length=hop_length))
    if power == 1:
        return S
    return S**power


# --------------------------------------------------------------------------
# A.2  mel filterbank and mel spectrogram
# --------------------------------------------------------------------------


def hz_to_mel(f):
    f = np.asanyarray(f, dtype=float)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if f.ndim:
        big = f >= min_log_hz
        mels[big] = min_log_mel + np.log(f[big] / min_log_hz) / logstep
    elif f >= min_log_hz:
        mels = min_log_mel + np.log(f / min_log_hz) / logstep
    return mels


def mel_to_hz(m):
    m = np.asanyarray(m, dtype=float)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    if m.ndim:
        big = m >= min_log_mel
        freqs[big] = min_log_hz * np.exp(logstep * (m[big] - min_log_mel))
    elif m >= min_log_mel:
        freqs = min_log_hz * np.exp(logstep * (m - min_log_mel))
    return freqs


def mel_frequencies(n_mels, fmin, fmax):
    return mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels))


def filters_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(htk=False, norm="slaney", dtype=float32)."""
    if fmax is None:
        fmax = float(sr) / 2
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = fft_frequencies(sr, n_fft)
    mel_f = mel_frequencies(n_mels + 2, fmin, fmax)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2 : n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def melspectrogram(y, sr, n_fft=2048, hop_length=512, power=2.0, n_mels=128, fmax=None):
    """librosa.feature.melspectrogram (structure.py:53; inside onset_strength)."""
    S = spectrogram(y, n_fft, hop_length, power)
    basis = filters_mel(sr, n_fft, n_mels=n_mels, fmax=fmax)
    return np.einsum("ft,mf->mt", S, basis, optimize=True)


# --------------------------------------------------------------------------
# A.4  dB conversions
# --------------------------------------------------------------------------


def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    S = np.asarray(S)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, np.abs(ref)))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def mfcc(S, n_mfcc=13):
    """librosa.feature.mfcc(S=log_power_mel, n_mfcc=13, dct_type=2, norm="ortho", lifter=0) as called at
    analysis/structure.py:199: the first n_mfcc rows of the orthonormal DCT-II along the mel axis."""
    import scipy.fft

    return scipy.fft.dct(np.asarray(S), axis=-2, type=2, norm="ortho")[..., :n_mfcc, :]


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    magnitude = np.abs(np.asarray(S))
    power = np.square(magnitude)
    return power_to_db(power, ref=np.abs(ref) ** 2, amin=amin**2, top_db=top_db)


# --------------------------------------------------------------------------
# A.3  onset strength
# --------------------------------------------------------------------------


def onset_strength(y=None, sr=22050, S=None, hop_length=512, n_fft=2048, lag=1,
                   aggregate=np.mean, center=True):
    """librosa.onset.onset_strength (tempo.py:19 with y; structure.py:195 with S)."""
    if S is None:
        S = np.abs(melspectrogram(y, sr, n_fft=n_fft, hop_length=hop_length, fmax=0.5 * sr))
        S = power_to_db(S)
    S = np.atleast_2d(S)
    env = S[..., lag:] - S[..., :-lag]
    env = np.maximum(0.0, env)
    env = aggregate(env, axis=-2)
    pad_width = lag
    if center:
        pad_width += n_fft // (2 * hop_length)
    env = np.pad(env, (int(pad_width), 0), mode="constant")
    if center:
        env = env[: S.shape[-1]]
    return env


# --------------------------------------------------------------------------
# A.8  autocorrelation and tempogram
# --------------------------------------------------------------------------


def autocorrelate(y, axis=-1):
    """librosa.autocorrelate (tempo.py:38).  float64 result under numpy 1.26."""
    y = np.asarray(y)
    n = y.shape[axis]
    n_pad = scipy.fft.next_fast_len(2 * n - 1, real=True)
    spec = np.fft.rfft(y.astype(np.float64), n=n_pad, axis=axis)
    powspec = spec.real**2 + spec.imag**2
    ac = np.fft.irfft(powspec, n=n_pad, axis=axis)
    sl = [slice(None)] * ac.ndim
    sl[axis] = slice(n)
    return ac[tuple(sl)]


def tempogram(y=None, sr=22050, onset_envelope=None, hop_length=512, win_length=384):
    """librosa.feature.tempogram (report.py:260), center=True, hann, norm=inf."""
    ac_window = get_window("hann", win_length)
    if onset_envelope is None:
        onset_envelope = onset_strength(y=y, sr=sr, hop_length=hop_length)
    n = onset_envelope.shape[-1]
    padded = np.pad(onset_envelope, (win_length // 2, win_length // 2),
                    mode="linear_ramp", end_values=[0, 0])
    odf = frame(padded, win_length, 1)[..., :n]
    return normalize(autocorrelate(odf * ac_window[:, None], axis=-2), norm=np.inf, axis=-2)


def tempo_frequencies(n_bins, hop_length, sr):
    out = np.zeros(n_bins, dtype=float)
    out[0] = np.inf
    out[1:] = 60.0 * sr / (hop_length * np.arange(1.0, n_bins))
    return out


# --------------------------------------------------------------------------
# A.5 / A.6 / A.11  spectral summaries and RMS
# --------------------------------------------------------------------------


def spectral_centroid(y, sr, n_fft=2048, hop_length=512):
    S = spectrogram(y, n_fft, hop_length, 1)
    freq = fft_frequencies(sr, n_fft)[:, None]
    return np.sum(freq * normalize(S, norm=1, axis=-2), axis=-2, keepdims=True)


def spectral_rolloff(y, sr, n_fft=2048, hop_length=512, roll_percent=0.85):
    S = spectrogram(y, n_fft, hop_length, 1)
    freq = fft_frequencies(sr, n_fft)[:, None]
    total = np.cumsum(S, axis=-2)
    threshold = roll_percent * total[-1, :]
    ind = np.where(total < threshold[None, :], np.nan, 1)
    return np.nanmin(ind * freq, axis=-2, keepdims=True)


def rms(y, frame_length=2048, hop_length=512):
    """librosa.feature.rms(center=True, pad_mode="constant") -> (1, 1 + N//hop) float32."""
    y = np.asarray(y)
    ypad = np.pad(y, (frame_length // 2, frame_length // 2), mode="constant")
    x = frame(ypad, frame_length, hop_length)
    sq = np.asfortranarray(np.square(np.abs(x)).astype(np.float32))
    power = np.mean(sq, axis=-2, keepdims=True)
    return np.sqrt(power)


# --------------------------------------------------------------------------
# A.10  chroma_stft (tuning estimation + chroma filterbank)
# --------------------------------------------------------------------------


def hz_to_octs(frequencies, tuning=0.0, bins_per_octave=12):
    A440 = 440.0 * 2.0 ** (tuning / bins_per_octave)
    return np.log2(np.asanyarray(frequencies) / (float(A440) / 16))


def _parabolic_interpolation(S):
    """0.10.2 stencil along axis -2: -b/a unless |b| >= |a|; edges 0."""
    shift = np.zeros_like(S)
    a = S[2:] + S[:-2] - 2 * S[1:-1]
    b = (S[2:] - S[:-2]) / 2
    with np.errstate(divide="ignore", invalid="ignore"):
        val = -b / a
    val = np.where(np.abs(b) >= np.abs(a), 0, val)
    shift[1:-1] = val
    return shift


def localmax(x):
    """librosa.util.localmax along axis -2 (first row never a max)."""
    out = np.zeros(x.shape, dtype=bool)
    out[1:-1] = (x[1:-1] > x[:-2]) & (x[1:-1] >= x[2:])
    out[-1] = x[-1] > x[-2]
    return out


def piptrack(S, sr, fmin=150.0, fmax=4000.0, threshold=0.1):
    S = np.abs(S)
    n_fft = 2 * (S.shape[-2] - 1)
    fmin = np.maximum(fmin, 0)
    fmax = np.minimum(fmax, float(sr) / 2)
    fft_freqs = fft_frequencies(sr, n_fft)
    avg = np.gradient(S, axis=-2)
    shift = _parabolic_interpolation(S)
    dskew = 0.5 * avg * shift
    pitches = np.zeros_like(S)
    mags = np.zeros_like(S)
    freq_mask = ((fmin <= fft_freqs) & (fft_freqs < fmax))[:, None]
    ref_value = threshold * np.max(S, axis=-2, keepdims=True)
    idx = np.nonzero(freq_mask & localmax(S * (S > ref_value)))
    pitches[idx] = (idx[0] + shift[idx]) * float(sr) / n_fft
    mags[idx] = S[idx] + dskew[idx]
    return pitches, mags


def pitch_tuning(frequencies, resolution=0.01, bins_per_octave=12):
    frequencies = np.atleast_1d(frequencies)
    frequencies = frequencies[frequencies > 0]
    if not np.any(frequencies):
        return 0.0
    residual = np.mod(bins_per_octave * hz_to_octs(frequencies), 1.0)
    residual[residual >= 0.5] -= 1.0
    bins = np.linspace(-0.5, 0.5, int(np.ceil(1.0 / resolution)) + 1)
    counts, tuning = np.histogram(residual, bins)
    return float(tuning[np.argmax(counts)])


def estimate_tuning(S, sr, bins_per_octave=12):
    pitch, mag = piptrack(S, sr)
    pitch_mask = pitch > 0
    if pitch_mask.any():
        threshold = np.median(mag[pitch_mask])
    else:
        threshold = 0.0
    return pitch_tuning(pitch[(mag >= threshold) & pitch_mask], bins_per_octave=bins_per_octave)


def filters_chroma(sr, n_fft, tuning=0.0, n_chroma=12, ctroct=5.0, octwidth=2.0):
    """librosa.filters.chroma(norm=2, base_c=True, dtype=float32)."""
    frequencies = np.linspace(0, sr, n_fft, endpoint=False)[1:]
    frqbins = n_chroma * hz_to_octs(frequencies, tuning=tuning, bins_per_octave=n_chroma)
    frqbins = np.concatenate(([frqbins[0] - 1.5 * n_chroma], frqbins))
    binwidthbins = np.concatenate((np.maximum(frqbins[1:] - frqbins[:-1], 1.0), [1]))
    D = np.subtract.outer(frqbins, np.arange(0, n_chroma, dtype="d")).T
    n_chroma2 = np.round(float(n_chroma) / 2)
    D = np.remainder(D + n_chroma2 + 10 * n_chroma, n_chroma) - n_chroma2
    wts = np.exp(-0.5 * (2 * D / np.tile(binwidthbins, (n_chroma, 1))) ** 2)
    wts = normalize(wts, norm=2, axis=0)
    wts *= np.tile(np.exp(-0.5 * (((frqbins / n_chroma - ctroct) / octwidth) ** 2)), (n_chroma, 1))
    wts = np.roll(wts, -3 * (n_chroma // 12), axis=0)
    return np.ascontiguousarray(wts[:, : int(1 + n_fft / 2)], dtype=np.float32)


def chroma_stft(y, sr, n_fft=2048, hop_length=512, n_chroma=12, return_tuning=False):
    """librosa.feature.chroma_stft(norm=inf, tuning=None) (harmony.py:108,149)."""
    S = spectrogram(y, n_fft, hop_length, 2)
    tuning = estimate_tuning(S, sr, bins_per_octave=n_chroma)
    fb = filters_chroma(sr, n_fft, tuning=tuning, n_chroma=n_chroma)
    raw = np.einsum("cf,ft->ct", fb, S, optimize=True)
    out = normalize(raw, norm=np.inf, axis=-2)
    if return_tuning:
        return out, tuning
    return out


# --------------------------------------------------------------------------
# A.12  HPSS (section 8f rank 1; consumed by structure.py:52)
# --------------------------------------------------------------------------


def softmask(X, X_ref, power=2.0):
    """librosa.util.softmask(split_zeros=True)."""
    Z = np.maximum(X, X_ref).astype(X.dtype)
    bad = Z < np.finfo(X.dtype).tiny
    Z[bad] = 1
    mask = (X / Z) ** power
    ref_mask = (X_ref / Z) ** power
    good = ~bad
    mask[good] /= mask[good] + ref_mask[good]
    mask[bad] = 0.5
    return mask


def _median_filter_reflect(S, kernel_size, axis):
    """scipy.ndimage.median_filter(S, size=kernel_size along `axis`, mode="reflect").

    The scipy in this image (1.18.1) returns a wrong median for lines of exactly two samples when the first is the
    larger one (its 1-D fast path; lengths 1, 3 ... 40 agree with the definition, checked against explicit symmetric
    padding).  The reference pins scipy 1.11.4, which has no such path, so that one length is evaluated from the
    definition: np.pad(mode="symmetric") is scipy's "reflect".
    """
    import scipy.ndimage

    if S.shape[axis] != 2:
        size = [1] * S.ndim
        size[axis] = kernel_size
        return scipy.ndimage.median_filter(S, size=tuple(size), mode="reflect")
    pad = [(0, 0)] * S.ndim
    pad[axis] = (kernel_size // 2, kernel_size // 2)
    win = np.lib.stride_tricks.sliding_window_view(np.pad(S, pad, mode="symmetric"), kernel_size, axis=axis)
    return np.median(win, axis=-1).astype(S.dtype)


def hpss(S, kernel_size=31, power=2.0, margin=1.0):
    harm = np.empty_like(S)
    harm[:] = _median_filter_reflect(S, kernel_size, axis=1)
    perc = np.empty_like(S)
    perc[:] = _median_filter_reflect(S, kernel_size, axis=0)
    mask_harm = softmask(harm, perc * margin, power=power)
    mask_perc = softmask(perc, harm * margin, power=power)
    return S * mask_harm, S * mask_perc
