"""``install()``: run the UNMODIFIED reference package on the B200 frontend.

The reference (``track_analyser``) reaches its hot path through 22 librosa entry points, ``pyloudnorm.Meter`` and
``resampy.resample`` (call sites listed per function below).  ``install()`` puts small modules
named ``librosa`` / ``pyloudnorm`` / ``resampy`` / ``audioread`` into ``sys.modules`` -- only where the real ones are not
importable -- whose functions answer exactly those calls from the kernels of ``libta_b200.so`` (through
``runtime.frontend``) plus the host-side decisions in ``hostlogic.py``.  The reference's own beat, structure, harmony,
loudness, feature and stereo code then runs byte for byte as written, on arrays the GPU produced.

Two things are replaced at function level instead, because their librosa calls cannot be answered array for array:
``track_analyser.stereo.frequency_dependent_width`` multiplies COMPLEX left / right spectra (stereo.py:95-110; the kernels
keep |mid|^2 and |side|^2 sums, not phases), and ``pipeline.analyse_track`` is wrapped in a ``frontend_session`` so the
>= 11 identical STFT requests of one call share one fused run.

``shim.stft`` returns the MAGNITUDE as a real float32 array: every remaining call site takes ``np.abs`` of it at once
(features.py:79-80, analysis/structure.py:48-51, harmony.py:254).  Arrays handed out by a shim are remembered (weakly, by
identity) together with the device results they came from, so that ``autocorrelate(onset_env)``, ``decompose.hpss(magnitude)``
and ``onset_strength(S=mel)`` are served by the device outputs of the same run when the reference passes them straight on,
and computed from the array itself otherwise.
"""

from __future__ import annotations

import sys
import types
import weakref

import numpy as np

from . import hostlogic, runtime

_registry: dict = {}   # id(array) -> (weakref to the array, TrackResult, sample rate, n_fft, hop)


def _remember(arr: np.ndarray, res, sr: int, n_fft: int, hop: int) -> np.ndarray:
    try:
        _registry[id(arr)] = (weakref.ref(arr, lambda _r, k=id(arr), reg=_registry: reg.pop(k, None)), res, sr, n_fft, hop)
    except TypeError:  # pragma: no cover - not weak-referenceable
        pass
    return arr


def _recall(arr):
    hit = _registry.get(id(arr))
    if hit is not None and hit[0]() is arr:
        return hit
    return None


def _mono32(y) -> np.ndarray:
    y = np.asarray(y)
    if y.ndim > 1:   # librosa.to_mono: mean over the leading axes
        y = np.mean(y, axis=tuple(range(y.ndim - 1)))
    return np.ascontiguousarray(y, dtype=np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# host-side pieces of librosa the reference calls on small arrays (restated from the published algorithms)
# ---------------------------------------------------------------------------------------------------------------------
def power_to_db(S, ref=1.0, amin=1e-10, top_db=80.0):
    """librosa.power_to_db (analysis/structure.py:194): 10 log10(max(amin, S)) - 10 log10(max(amin, ref)), floored top_db below the max."""
    S = np.asarray(S)
    ref_value = ref(S) if callable(ref) else np.abs(ref)
    log_spec = 10.0 * np.log10(np.maximum(amin, S))
    log_spec -= 10.0 * np.log10(np.maximum(amin, ref_value))
    if top_db is not None:
        log_spec = np.maximum(log_spec, log_spec.max() - top_db)
    return log_spec


def amplitude_to_db(S, ref=1.0, amin=1e-5, top_db=80.0):
    """librosa.amplitude_to_db (analysis/loudness.py:42)."""
    magnitude = np.abs(np.asarray(S))
    ref_value = ref(magnitude) if callable(ref) else np.abs(ref)
    return power_to_db(np.square(magnitude), ref=ref_value**2, amin=amin**2, top_db=top_db)


def normalize(S, norm=np.inf, axis=0):
    """librosa.util.normalize(norm=inf) as tempo.py:50 calls it on a 1-d slice."""
    S = np.asarray(S)
    mag = np.abs(S).astype(float)
    if norm == np.inf:
        length = np.max(mag, axis=axis, keepdims=True)
    elif norm == 1:
        length = np.sum(mag, axis=axis, keepdims=True)
    elif norm == 2:
        length = np.sum(mag**2, axis=axis, keepdims=True) ** 0.5
    else:
        raise ValueError(f"unsupported norm {norm!r}")
    length[length < hostlogic.tiny(S)] = 1.0
    out = np.empty_like(S)
    out[:] = S / length
    return out


def _onset_strength_from_S(S, n_fft=2048, hop_length=512, lag=1, aggregate=np.mean):
    """librosa.onset.onset_strength(S=...) (analysis/structure.py:195): lag-1 difference, half-wave rectification, mean over
    the rows, left pad of lag + n_fft // (2 hop) frames, trimmed to the input's frame count."""
    S = np.atleast_2d(np.asarray(S))
    env = np.maximum(0.0, S[..., lag:] - S[..., :-lag])
    env = (aggregate or np.mean)(env, axis=-2)
    pad = lag + n_fft // (2 * hop_length)
    env = np.pad(env, (int(pad), 0), mode="constant")
    return env[: S.shape[-1]]


def _autocorrelate_host(y):
    """librosa.autocorrelate: irfft(|rfft(y, n_pad)|^2)[:n] in float64 (numpy 1.26 transforms float32 input in float64)."""
    import scipy.fft

    y = np.asarray(y)
    n = y.shape[-1]
    n_pad = scipy.fft.next_fast_len(2 * n - 1, real=True)
    spec = np.fft.rfft(y.astype(np.float64), n=n_pad)
    return np.fft.irfft(spec.real**2 + spec.imag**2, n=n_pad)[:n]


# ---------------------------------------------------------------------------------------------------------------------
# the shim functions (one per reference call site)
# ---------------------------------------------------------------------------------------------------------------------
def stft(y, n_fft=2048, hop_length=None, win_length=None, window="hann", center=True, dtype=None, pad_mode="constant", **_):
    """features.py:79, analysis/structure.py:48, harmony.py:254 -> |STFT| (see the module docstring), float32 (1 + n_fft/2, T)."""
    if not center or pad_mode != "constant" or (win_length not in (None, n_fft)):
        raise NotImplementedError("the frontend computes centred, zero-padded, full-window STFTs (librosa's defaults)")
    hop = n_fft // 4 if hop_length is None else int(hop_length)
    y32 = _mono32(y) if np.ndim(y) > 1 else np.ascontiguousarray(y, dtype=np.float32)
    sr = _stft_rate.get("sr", 22_050)
    res = runtime.frontend(y32, sr, n_fft=n_fft, hop=hop, n_mels=128 if n_fft == 2048 and hop == 512 else 0,
                           outputs=("magnitude",), window=window)
    return _remember(np.asarray(res["magnitude"]), res, sr, n_fft, hop)


_stft_rate: dict = {}   # librosa.stft takes no sample rate; magnitude does not depend on it, the session key does


def fft_frequencies(*, sr=22_050, n_fft=2048):
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)


def frames_to_time(frames, *, sr=22_050, hop_length=512, n_fft=None):
    return hostlogic.frames_to_time(frames, sr, hop_length)


def time_to_frames(times, *, sr=22_050, hop_length=512, n_fft=None):
    return hostlogic.time_to_frames(times, sr, hop_length)


def midi_to_hz(notes):
    """librosa.midi_to_hz (the reference's tests build their chords with it): 440 * 2^((m - 69) / 12)."""
    return 440.0 * (2.0 ** ((np.asanyarray(notes) - 69.0) / 12.0))


def hz_to_midi(frequencies):
    return 12.0 * (np.log2(np.asanyarray(frequencies)) - np.log2(440.0)) + 69.0


def tempo_frequencies(n_bins, *, hop_length=512, sr=22_050):
    out = np.zeros(int(n_bins), dtype=float)
    out[0] = np.inf
    out[1:] = 60.0 * sr / (hop_length * np.arange(1.0, n_bins))
    return out


def autocorrelate(y, *, max_size=None, axis=-1):
    """tempo.py:38: the device's float64 autocorrelation when ``y`` is the envelope a shim returned, else from ``y`` itself."""
    hit = _recall(y)
    if hit is not None and "autocorr" in hit[1] and max_size is None and np.ndim(y) == 1:
        return np.asarray(hit[1]["autocorr"])
    return _autocorrelate_host(y)


def resample(y, *, orig_sr, target_sr, res_type="soxr_hq", **_):
    """io.py:44,51 / utils.py:61,68 (only reached when resampy is missing; the shim provides resampy, so this is a courtesy)."""
    from .resample import resample as rs

    return rs(np.asarray(y, dtype=np.float32), int(orig_sr), int(target_sr))


def onset_strength(*, y=None, sr=22_050, S=None, lag=1, max_size=1, ref=None, detrend=False, center=True, feature=None,
                   aggregate=None, hop_length=512, n_fft=2048, **_):
    """tempo.py:19 (y=..., the dB mel flux: device) and analysis/structure.py:195 (S=linear mel power)."""
    if S is not None:
        hit = _recall(S)
        if hit is not None and "flux_linear" in hit[1] and hit[3] == 2048:
            return np.asarray(hit[1]["flux_linear"], dtype=float)
        return _onset_strength_from_S(S, n_fft=n_fft, hop_length=hop_length, lag=lag, aggregate=aggregate)
    y32 = _mono32(y)
    res = runtime.frontend(y32, int(sr), n_fft=2048, hop=int(hop_length), outputs=("onset_env", "autocorr"))
    return _remember(np.asarray(res["onset_env"]), res, int(sr), 2048, int(hop_length))


def onset_detect(*, y=None, sr=22_050, onset_envelope=None, hop_length=512, backtrack=False, energy=None, units="frames",
                 normalize=True, **kwargs):
    """tempo.py:81,100."""
    if onset_envelope is None:
        onset_envelope = onset_strength(y=y, sr=sr, hop_length=hop_length)
    return hostlogic.onset_detect(np.asarray(onset_envelope), int(sr), int(hop_length), backtrack=backtrack, units=units)


def peak_pick(x, *, pre_max, post_max, pre_avg, post_avg, delta, wait):
    """analysis/structure.py:89."""
    return hostlogic.peak_pick(np.asarray(x), pre_max, post_max, pre_avg, post_avg, delta, wait)


def buf_to_float(x, *, n_bytes=2, dtype=np.float32):
    scale = 1.0 / float(1 << ((8 * n_bytes) - 1))
    return scale * np.frombuffer(x, f"<i{n_bytes:d}").astype(dtype)


def spectral_centroid(*, y=None, sr=22_050, S=None, n_fft=2048, hop_length=512, **_):
    """features.py:97 -> (1, T) float64."""
    res = runtime.frontend(_mono32(y), int(sr), n_fft=n_fft, hop=hop_length, outputs=("centroid",))
    return np.asarray(res["centroid"], dtype=np.float64)[None, :]


def spectral_rolloff(*, y=None, sr=22_050, S=None, n_fft=2048, hop_length=512, roll_percent=0.85, **_):
    """features.py:116 -> (1, T) float64."""
    res = runtime.frontend(_mono32(y), int(sr), n_fft=n_fft, hop=hop_length, roll_percent=roll_percent, outputs=("rolloff_bin",))
    return np.fft.rfftfreq(n=n_fft, d=1.0 / sr)[np.asarray(res["rolloff_bin"])][None, :]


def chroma_stft(*, y=None, sr=22_050, **_):
    """harmony.py:108,149."""
    return np.asarray(runtime.frontend(_mono32(y), int(sr), outputs=("chroma", "tuning"))["chroma"])


def chroma_cqt(*, y=None, sr=22_050, **_):
    """harmony.py:107,148."""
    from .harmony import _chroma_cqt

    return np.asarray(_chroma_cqt(_mono32(y), int(sr)))


def melspectrogram(*, y=None, sr=22_050, n_fft=2048, hop_length=512, power=2.0, n_mels=128, **_):
    """analysis/structure.py:53-59."""
    if power != 2.0:
        raise NotImplementedError("the frontend's mel projection is on the power spectrogram (power=2.0)")
    res = runtime.frontend(_mono32(y), int(sr), n_fft=n_fft, hop=hop_length, n_mels=n_mels, outputs=("mel", "flux_linear"))
    return _remember(np.asarray(res["mel"]), res, int(sr), n_fft, hop_length)


def mfcc(*, y=None, sr=22_050, S=None, n_mfcc=20, dct_type=2, norm="ortho", lifter=0, **_):
    """analysis/structure.py:199: the first n_mfcc rows of the orthonormal DCT-II of the given log-mel matrix."""
    import scipy.fft

    if S is None:
        raise NotImplementedError("the reference calls mfcc with S=log_mel")
    return scipy.fft.dct(np.asarray(S), axis=-2, type=dct_type, norm=norm)[..., :n_mfcc, :]


def rms(*, y=None, S=None, frame_length=2048, hop_length=512, center=True, pad_mode="constant", **_):
    """analysis/loudness.py:39 -> (1, 1 + N // hop) float32, from the device's centred frame mean squares."""
    y32 = _mono32(y)
    sr = _stft_rate.get("sr", 44_100)
    plan = runtime.get_plan(sr)
    for seconds, key in ((3.0, "rms_short"), (plan.meter_block, "rms_momentary")):
        if plan.rms_frames(seconds) == (int(frame_length), int(hop_length)):
            ms = runtime.frontend(y32, sr, outputs=(key,))[key]
            return np.sqrt(np.asarray(ms, dtype=np.float64)).astype(np.float32)[None, :]
    block = frame_length / float(sr)   # another window: a plan whose momentary frame is this one
    plan = runtime.get_plan(sr, meter_block=block)
    if plan.rms_frames(block) != (int(frame_length), int(hop_length)):
        raise NotImplementedError("rms frames other than the reference's (frame, frame // 2) windows")
    ms = runtime.frontend(y32, sr, meter_block=block, outputs=("rms_momentary",))["rms_momentary"]
    return np.sqrt(np.asarray(ms, dtype=np.float64)).astype(np.float32)[None, :]


def tempogram(*, y=None, sr=22_050, onset_envelope=None, hop_length=512, win_length=384, **_):
    """report.py:260."""
    res = runtime.frontend(_mono32(y), int(sr), hop=hop_length, outputs=("tempogram",))
    return np.asarray(res["tempogram"], dtype=float)


def hpss(S, *, kernel_size=31, power=2.0, mask=False, margin=1.0):
    """analysis/structure.py:52: the two component matrices, by the device's median kernels on the magnitude that is still
    resident when ``S`` is the array ``stft`` returned (re-uploaded otherwise)."""
    import ctypes as C

    import torch

    from . import _native as nat
    from . import engine

    if kernel_size != 31 or power != 2.0 or mask or margin != 1.0:
        raise NotImplementedError("the device kernels implement librosa.decompose.hpss's defaults")
    S = np.asarray(S)
    n_fft = 2 * (S.shape[0] - 1)
    T = S.shape[1]
    hit = _recall(S)
    sr = hit[2] if hit else _stft_rate.get("sr", 44_100)
    hop = hit[4] if hit else 512
    plan = runtime.get_plan(sr, n_fft, hop, 128 if (n_fft, hop) == (2048, 512) else 0)
    dev = torch.device(f"cuda:{plan.device}")
    # a one-track batch of T frames: (T - 1) * hop samples describe the layout, the PCM itself is not read
    n = (T - 1) * hop
    ld = engine.frame_pitch(T)
    pcm = torch.zeros(4, dtype=torch.float32, device=dev)
    batch = engine.DeviceBatch(plan, pcm, np.zeros(1, np.int64), np.asarray([n], np.int64), 1)
    mag = torch.zeros(S.shape[0] * ld, dtype=torch.float32, device=dev)
    mag.view(S.shape[0], ld)[:, :T].copy_(torch.from_numpy(np.ascontiguousarray(S, dtype=np.float32)))
    scratch, harm, perc = (torch.empty_like(mag) for _ in range(3))
    hs, ps = (torch.empty(ld, dtype=torch.float32, device=dev) for _ in range(2))
    ws = engine.workspace(plan, batch)
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    nat.check(plan.lib.ta_hpss_components(plan._h, C.byref(batch.c_batch), C.c_void_p(mag.data_ptr()), C.c_void_p(scratch.data_ptr()),
                                          C.c_void_p(harm.data_ptr()), C.c_void_p(perc.data_ptr()), C.c_void_p(hs.data_ptr()),
                                          C.c_void_p(ps.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), stream))
    out = [np.ascontiguousarray(t.view(S.shape[0], ld)[:, :T].cpu().numpy()) for t in (harm, perc)]
    return out[0], out[1]


class Meter:
    """pyloudnorm.Meter(rate, block_size=...).integrated_loudness(data) (analysis/loudness.py:60-61).  No ``loudness_range``
    attribute: pyloudnorm 0.1.1, the reference's pin, has none (loudness.py:62-68 takes its percentile branch)."""

    def __init__(self, rate, filter_class="K-weighting", block_size=0.400):
        if filter_class != "K-weighting":
            raise NotImplementedError("the device kernel implements the K-weighting filter chain")
        self.rate, self.block_size = int(rate), float(block_size)

    def integrated_loudness(self, data):
        x = np.asarray(data)
        if x.ndim > 1:
            raise NotImplementedError("the reference meters mono signals")
        if x.shape[0] < self.block_size * self.rate:
            raise ValueError("Audio must have length greater than the block size.")
        res = runtime.frontend(np.ascontiguousarray(x, dtype=np.float32), self.rate, meter_block=self.block_size, outputs=("lufs",))
        return float(res["lufs"])


class NoBackendError(Exception):
    """audioread.NoBackendError: the shim decodes nothing (WAV files go through the soundfile-shaped reader of io.py)."""


def _audio_open(path):
    raise NoBackendError(f"no decoder for {path!r}: the B200 frontend reads RIFF/WAVE only")


def _module(name: str, **members) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(members)
    m.__ta_b200_shim__ = True
    return m


def build_shims() -> dict:
    """{module name: module object} for librosa (+ submodules), pyloudnorm, resampy, audioread."""
    from .resample import resample as rs

    util = _module("librosa.util", normalize=normalize, peak_pick=peak_pick, buf_to_float=buf_to_float)
    onset = _module("librosa.onset", onset_strength=onset_strength, onset_detect=onset_detect)
    feature = _module("librosa.feature", spectral_centroid=spectral_centroid, spectral_rolloff=spectral_rolloff,
                      chroma_stft=chroma_stft, chroma_cqt=chroma_cqt, melspectrogram=melspectrogram, mfcc=mfcc, rms=rms,
                      tempogram=tempogram)
    decompose = _module("librosa.decompose", hpss=hpss)
    librosa = _module("librosa", stft=stft, midi_to_hz=midi_to_hz, hz_to_midi=hz_to_midi, fft_frequencies=fft_frequencies, frames_to_time=frames_to_time,
                      time_to_frames=time_to_frames, tempo_frequencies=tempo_frequencies, autocorrelate=autocorrelate,
                      resample=resample, power_to_db=power_to_db, amplitude_to_db=amplitude_to_db, util=util, onset=onset,
                      feature=feature, decompose=decompose, __version__="0.10.2.post1+ta_b200")
    librosa.__path__ = []   # a package: "from librosa import util" and "import librosa.onset" both resolve
    return {
        "librosa": librosa, "librosa.util": util, "librosa.onset": onset, "librosa.feature": feature, "librosa.decompose": decompose,
        "pyloudnorm": _module("pyloudnorm", Meter=Meter),
        "resampy": _module("resampy", resample=lambda x, sr_orig, sr_new, **_: rs(np.asarray(x, dtype=np.float32), int(sr_orig), int(sr_new))),
        "audioread": _module("audioread", audio_open=_audio_open, NoBackendError=NoBackendError),
    }


def install(reference_package: str = "track_analyser", *, force: bool = False):
    """Make ``import track_analyser`` (the unmodified reference) run on the B200 frontend.  Returns the reference package.

    Shim modules are registered for every third-party module above that is not importable (all of them with ``force``);
    then the reference is imported, ``stereo.frequency_dependent_width`` is pointed at the frontend's band-energy
    implementation, ``io.load_audio`` at the WAV reader of ``io.py`` when neither soundfile nor a real audioread exists,
    and ``pipeline.analyse_track`` is wrapped in a ``frontend_session``."""
    import functools
    import importlib
    import importlib.util

    shims = build_shims()
    installed = []
    for name in ("librosa", "pyloudnorm", "resampy", "audioread"):
        real = False
        if not force and name not in sys.modules:
            try:
                real = importlib.util.find_spec(name) is not None
            except (ImportError, ValueError):
                real = False
        elif not force:
            real = not getattr(sys.modules[name], "__ta_b200_shim__", False)
        if real:
            continue
        for mod_name, mod in shims.items():
            if mod_name == name or mod_name.startswith(name + "."):
                sys.modules[mod_name] = mod
        installed.append(name)
    ref = importlib.import_module(reference_package)
    ref_stereo = importlib.import_module(reference_package + ".stereo")
    ref_pipeline = importlib.import_module(reference_package + ".pipeline")
    ref_io = importlib.import_module(reference_package + ".io")
    ref_utils = importlib.import_module(reference_package + ".utils")
    from . import io as our_io
    from . import stereo as our_stereo

    if "librosa" in installed:
        ref_stereo.frequency_dependent_width = our_stereo.frequency_dependent_width
    if "audioread" in installed and getattr(ref_io, "sf", None) is None:
        ref_io.load_audio = our_io.load_audio
        ref_utils.load_audio = our_io.load_audio
    if not getattr(ref_pipeline.analyse_track, "__ta_b200_wrapped__", False):
        inner = ref_pipeline.analyse_track

        @functools.wraps(inner)
        def analyse_track(source, *args, **kwargs):
            with runtime.frontend_session():
                audio = source if isinstance(source, ref_utils.AudioInput) else ref_utils.coerce_audio(source)
                _stft_rate["sr"] = int(audio.sample_rate)
                if audio.stereo_samples is not None:
                    runtime.alias_mono_to_stereo(np.asarray(audio.samples), np.asarray(audio.stereo_samples))
                return inner(audio, *args, **kwargs)

        analyse_track.__ta_b200_wrapped__ = True
        ref_pipeline.analyse_track = analyse_track
        ref.analyse_track = analyse_track
    ref.__ta_b200_installed__ = tuple(installed)
    return ref
