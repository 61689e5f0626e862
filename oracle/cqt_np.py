"""numpy restatement of ``librosa.feature.chroma_cqt`` (librosa 0.10.2.post1) as the reference calls it.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: librosa is not installable in this
image and the reference holds no vectors for this path; additionally librosa's octave decimator is
libsoxr's ``soxr_hq`` (a third-party C library, not on this machine), so the decimator here is a STATED,
SWAPPABLE stage: a zero-phase Kaiser-windowed-sinc low-pass built from soxr's published HQ band edges
(pass band to 0.913 of the new Nyquist, stop band from 1.0, 20-bit = 126.4 dB rejection).  What checks
this file: ``brute_force_cqt`` below (direct float64 correlation of the full-rate signal with every
constant-Q atom: no decimation, no FFT, no sparsification) and the property tests in
``tests/test_oracle_cqt.py``.

Reference call sites: /root/reference/src/track_analyser/harmony.py:107 (``key_estimate``) and :148
(``analyse_harmony``): ``librosa.feature.chroma_cqt(y=y, sr=sr)`` with every default, i.e.
hop_length=512, fmin=C1, n_chroma=12, n_octaves=7, bins_per_octave=36, norm=inf, threshold=0.0,
tuning=None (estimated from y), cqt_mode="full" -> ``librosa.cqt`` = ``vqt(gamma=0, filter_scale=1,
norm=1, sparsity=0.01, window="hann", scale=True, pad_mode="constant", res_type="soxr_hq")``.
"""

from __future__ import annotations

import numpy as np
import scipy.signal

from . import librosa_np as lr

C1_HZ = 32.70319566257483  # librosa.note_to_hz("C1") = 440 * 2 ** ((24 - 69) / 12)
HANN_BANDWIDTH = 1.50018310546875  # librosa.filters.WINDOW_BANDWIDTHS["hann"]

# ---- decimator (the swappable stage) ---------------------------------------------------------------------------

SOXR_HQ_PASSBAND_END = 0.913    # soxr.h: quality HQ, "0dB pt. bandwidth to preserve; nyquist = 1"
SOXR_HQ_STOPBAND_BEGIN = 1.0    # soxr.h: "aliasing/imaging control"
SOXR_HQ_REJECTION_DB = 21 * 20.0 * np.log10(2.0)   # (bits + 1) * 6.02 dB with 20 bits of precision


def decimator_taps():
    """Odd-length symmetric FIR for 2:1 decimation, unit DC gain, float64.  Design: Kaiser window method on the
    band edges above (Kaiser's length and beta formulas), cut-off at the centre of the transition band."""
    att = SOXR_HQ_REJECTION_DB
    beta = 0.1102 * (att - 8.7)
    width = np.pi * (SOXR_HQ_STOPBAND_BEGIN - SOXR_HQ_PASSBAND_END) / 2.0   # rad/sample at the input rate
    n = int(np.ceil((att - 7.95) / (2.285 * width))) + 1
    if n % 2 == 0:
        n += 1
    half = n // 2
    fc = 0.25 * (SOXR_HQ_PASSBAND_END + SOXR_HQ_STOPBAND_BEGIN) / 2.0      # cycles/sample at the input rate
    m = np.arange(-half, half + 1, dtype=float)
    h = 2.0 * fc * np.sinc(2.0 * fc * m) * np.kaiser(n, beta)
    return h / np.sum(h)


def decimate2(y, taps=None):
    """librosa.resample(y, orig_sr=2, target_sr=1, res_type="soxr_hq", scale=True) (librosa/core/audio.py) with the
    stated decimator: out[m] = sqrt(2) * sum_j h[j] * y[2m + j] (zero outside the signal), ceil(n/2) samples,
    float32 out like soxr's float32 engine and ``np.asarray(y_hat, dtype=y.dtype)``."""
    h = decimator_taps() if taps is None else taps
    y = np.asarray(y)
    n_out = (y.shape[-1] + 1) // 2
    full = scipy.signal.fftconvolve(y.astype(np.float64), h, mode="full")
    half = len(h) // 2
    out = full[half: half + 2 * n_out: 2]
    if out.shape[-1] < n_out:
        out = np.pad(out, (0, n_out - out.shape[-1]))
    return (out / np.sqrt(0.5)).astype(y.dtype)


# ---- librosa.filters.wavelet_lengths / wavelet -----------------------------------------------------------------


def et_relative_bw(bins_per_octave):
    """librosa.core.constantq.__et_relative_bw."""
    r = 2.0 ** (1.0 / bins_per_octave)
    return np.atleast_1d((r**2 - 1) / (r**2 + 1))


def wavelet_lengths(freqs, sr, filter_scale=1.0, gamma=0.0, alpha=None):
    """librosa.filters.wavelet_lengths(window="hann") -> (lengths float64, filter_cutoff)."""
    freqs = np.asarray(freqs, dtype=float)
    alpha = np.asarray(alpha, dtype=float)
    Q = float(filter_scale) / alpha
    filter_cutoff = max(freqs * (1 + 0.5 * HANN_BANDWIDTH / Q) + 0.5 * gamma)
    lengths = Q * sr / (freqs + gamma / alpha)
    return lengths, filter_cutoff


def wavelet(freqs, sr, alpha):
    """librosa.filters.wavelet(window="hann", filter_scale=1, pad_fft=True, norm=1, dtype=complex64, gamma=0)."""
    lengths, _ = wavelet_lengths(freqs, sr, alpha=alpha)
    filters = []
    for ilen, freq in zip(lengths, freqs):
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        ang = t * 2 * np.pi * freq / sr
        sig = np.cos(ang) + 1j * np.sin(ang)             # util.phasor
        sig = sig * lr.get_window("hann", len(sig))       # __float_window("hann")(len(sig))
        sig = sig / np.sum(np.abs(sig))                   # util.normalize(norm=1)
        filters.append(sig)
    max_len = int(2.0 ** (np.ceil(np.log2(max(lengths)))))
    out = np.zeros((len(filters), max_len), dtype=np.complex64)
    for i, f in enumerate(filters):
        lpad = (max_len - len(f)) // 2                    # util.pad_center
        out[i, lpad: lpad + len(f)] = f
    return out, lengths


def sparsify_rows(x, quantile=0.01):
    """librosa.util.sparsify_rows -> dense complex64 array with the dropped entries zeroed."""
    mags = np.abs(x)
    norms = np.sum(mags, axis=1, keepdims=True)
    mag_sort = np.sort(mags, axis=1)
    cumulative_mag = np.cumsum(mag_sort / norms, axis=1)
    threshold_idx = np.argmin(cumulative_mag < quantile, axis=1)
    out = np.zeros(x.shape, dtype=np.complex64)
    for i, j in enumerate(threshold_idx):
        keep = mags[i] >= mag_sort[i, j]
        out[i, keep] = x[i, keep]
    return out


def vqt_filter_fft(sr, freqs, alpha):
    """librosa.core.constantq.__vqt_filter_fft(filter_scale=1, norm=1, sparsity=0.01, hop_length=None)."""
    basis, lengths = wavelet(freqs, sr, alpha)
    n_fft = basis.shape[1]
    basis *= (lengths[:, np.newaxis] / float(n_fft))      # in place on complex64
    fft_basis = np.fft.fft(basis.astype(np.complex128), n=n_fft, axis=1)[:, : (n_fft // 2) + 1]
    return sparsify_rows(fft_basis, quantile=0.01), n_fft, lengths


def early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves):
    c1 = max(0, int(np.ceil(np.log2(nyquist / filter_cutoff)) - 1) - 1)
    twos = 0
    h = hop_length
    while h > 0 and h % 2 == 0:
        twos += 1
        h //= 2
    return min(c1, max(0, twos - n_octaves + 1))


def stft_ones(y, n_fft, hop_length):
    """librosa.stft(window="ones", center=True, pad_mode="constant") -> complex64 (float64 transform)."""
    y = np.asarray(y)
    ypad = np.pad(y, (n_fft // 2, n_fft // 2))
    frames = lr.frame(ypad, n_fft, hop_length)
    T = frames.shape[-1]
    out = np.zeros((1 + n_fft // 2, T), dtype=np.complex64)
    step = max(1, (2**24) // (n_fft * 8))
    for s in range(0, T, step):
        out[:, s: s + step] = np.fft.rfft(frames[:, s: s + step].astype(np.float64), axis=0)
    return out


# ---- estimate_tuning(y=...) -------------------------------------------------------------------------------------


def estimate_tuning_y(y, sr, bins_per_octave):
    """librosa.estimate_tuning(y=y, sr=sr, bins_per_octave=bpo): piptrack on the MAGNITUDE spectrogram
    (_spectrogram(power=1), n_fft=2048, hop 512), unlike chroma_stft which hands it the power spectrogram."""
    S = lr.spectrogram(np.asarray(y), 2048, 512, 1)
    return lr.estimate_tuning(S, sr, bins_per_octave=bins_per_octave)


# ---- cqt ----------------------------------------------------------------------------------------------------------


def cqt_plan(sr, hop_length=512, n_bins=252, bins_per_octave=36, tuning=0.0, fmin=None):
    """The static part of ``vqt``: frequencies, early down-sampling count and, per octave, (rate, hop, freqs)."""
    n_octaves = int(np.ceil(float(n_bins) / bins_per_octave))
    n_filters = min(bins_per_octave, n_bins)
    fmin = C1_HZ if fmin is None else fmin
    fmin = fmin * 2.0 ** (tuning / bins_per_octave)
    freqs = fmin * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave)   # interval_frequencies("equal")
    alpha = et_relative_bw(bins_per_octave)
    lengths, filter_cutoff = wavelet_lengths(freqs, sr, alpha=alpha)
    nyquist = sr / 2.0
    if filter_cutoff > nyquist:
        raise ValueError("wavelet basis with max frequency would exceed the Nyquist frequency")
    early = early_downsample_count(nyquist, filter_cutoff, hop_length, n_octaves)
    sr0 = sr / float(2**early)
    hop0 = hop_length // (2**early)
    octaves = []
    my_sr, my_hop, stage = sr0, hop0, early
    for i in range(n_octaves):
        sl = slice(-n_filters, None) if i == 0 else slice(-n_filters * (i + 1), -n_filters * i)
        octaves.append(dict(sr=my_sr, hop=my_hop, freqs=freqs[sl], stage=stage))
        if my_hop % 2 == 0:
            my_hop //= 2
            my_sr /= 2.0
            stage += 1
    lengths_final, _ = wavelet_lengths(freqs, sr0, alpha=alpha)
    return dict(freqs=freqs, alpha=alpha, early=early, sr0=sr0, octaves=octaves, lengths=lengths_final,
                n_bins=n_bins, n_filters=n_filters)


def cqt(y, sr, hop_length=512, n_bins=252, bins_per_octave=36, tuning=None, decimate=decimate2):
    """librosa.cqt(y, sr, hop_length, fmin=None, n_bins, bins_per_octave, tuning) -> complex64 (n_bins, T)."""
    y = np.asarray(y, dtype=np.float32)
    if tuning is None:
        tuning = estimate_tuning_y(y, sr, bins_per_octave)
    plan = cqt_plan(sr, hop_length, n_bins, bins_per_octave, tuning)
    signals = {0: y}

    def signal(stage):
        if stage not in signals:
            signals[stage] = decimate(signal(stage - 1))
        return signals[stage]

    resp = []
    for o in plan["octaves"]:
        fft_basis, n_fft, _ = vqt_filter_fft(o["sr"], o["freqs"], plan["alpha"])
        fft_basis = (fft_basis * np.float32(np.sqrt(plan["sr0"] / o["sr"]))).astype(np.complex64)
        D = stft_ones(signal(o["stage"]), n_fft, o["hop"])
        resp.append((fft_basis @ D).astype(np.complex64))
    max_col = min(r.shape[-1] for r in resp)
    V = np.empty((n_bins, max_col), dtype=np.complex64)
    end = n_bins
    for r in resp:                                            # __trim_stack
        n_oct = r.shape[0]
        if end < n_oct:
            V[:end] = r[-end:, :max_col]
        else:
            V[end - n_oct: end] = r[:, :max_col]
        end -= n_oct
    V /= np.sqrt(plan["lengths"])[:, None]                    # scale=True
    return V


def cq_to_chroma(n_input, bins_per_octave=36, n_chroma=12):
    """librosa.filters.cq_to_chroma(fmin=None -> C1, base_c=True, window=None, dtype=float32)."""
    n_merge = bins_per_octave // n_chroma
    m = np.repeat(np.eye(n_chroma), n_merge, axis=1)
    m = np.roll(m, -(n_merge // 2), axis=1)
    n_oct = int(np.ceil(float(n_input) / bins_per_octave))
    m = np.tile(m, n_oct)[:, :n_input]
    midi_0 = np.mod(12 * (np.log2(C1_HZ) - np.log2(440.0)) + 69, 12)
    roll = int(np.round(midi_0 * (n_chroma / 12.0)))
    return np.roll(m, roll, axis=0).astype(np.float32)


def chroma_cqt(y, sr, hop_length=512, return_parts=False, decimate=decimate2):
    """librosa.feature.chroma_cqt(y=y, sr=sr) (harmony.py:107,148) -> float32 (12, T)."""
    y = np.asarray(y, dtype=np.float32)
    tuning = estimate_tuning_y(y, sr, 36)
    C = np.abs(cqt(y, sr, hop_length=hop_length, n_bins=252, bins_per_octave=36, tuning=tuning, decimate=decimate))
    raw = np.einsum("cf,ft->ct", cq_to_chroma(C.shape[0]), C, optimize=True)
    raw[raw < 0.0] = 0.0
    chroma = lr.normalize(raw, norm=np.inf, axis=-2)
    if return_parts:
        return chroma, C, tuning
    return chroma


# ---- independent cross-check ---------------------------------------------------------------------------------------


def brute_force_cqt(y, sr, hop_length=512, n_bins=252, bins_per_octave=36, tuning=0.0, frames=None):
    """|CQT| from the definition, for checking the multi-rate restatement above: every bin is the float64 inner
    product of the FULL-RATE signal with its Hann-windowed complex exponential of length Q*sr/f centred on frame
    t*hop, L1-normalised, times sqrt(length) (librosa's ``scale=True`` convention).  No decimation, no FFT, no basis
    sparsification: agreement is limited by the 1 % sparsification of the spectral basis (about 1e-2 relative), which
    is what tests/test_oracle_cqt.py allows."""
    y = np.asarray(y, dtype=np.float64)
    fmin = C1_HZ * 2.0 ** (tuning / bins_per_octave)
    freqs = fmin * 2.0 ** (np.arange(n_bins, dtype=float) / bins_per_octave)
    alpha = et_relative_bw(bins_per_octave)
    lengths, _ = wavelet_lengths(freqs, sr, alpha=alpha)
    T = 1 + len(y) // hop_length
    frames = np.arange(T) if frames is None else np.asarray(frames)
    out = np.zeros((n_bins, len(frames)))
    for k, (ilen, f) in enumerate(zip(lengths, freqs)):
        t = np.arange(-ilen // 2, ilen // 2, dtype=float)
        atom = np.exp(-1j * t * 2 * np.pi * f / sr) * lr.get_window("hann", len(t))
        atom /= np.sum(np.abs(atom))
        off = -int(np.floor(ilen / 2))   # the spectral product H[j]*X[j] correlates with the time-reversed atom
        pad = len(t)
        ypad = np.pad(y, (pad, pad))
        for j, fr in enumerate(frames):
            s = fr * hop_length + off + pad
            out[k, j] = np.abs(np.dot(ypad[s: s + len(t)], atom)) * np.sqrt(ilen)
    return out
