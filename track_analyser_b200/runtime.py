"""Plan registry and the per-call frontend cache.

The reference recomputes the same mono 2048/512 Hann STFT at least eleven times
per ``analyse_track``.  Inside a ``frontend_session()``
every module-level function asks this cache for the results of ONE fused GPU run
keyed on the sample buffer; outside a session each call runs the kernels afresh.
"""

from __future__ import annotations

import collections
import contextlib
import threading

import numpy as np

from . import engine

_MAX_PLANS = 16  # least-recently-used plans beyond this are closed (each owns device tables and a workspace)
_plans: "collections.OrderedDict" = collections.OrderedDict()
_plans_lock = threading.Lock()
_local = threading.local()

# What a session computes on the first request for a (buffer, plan): the default 2048/512/128 plan serves ~11 call
# sites of analyse_track, so it runs everything once; any other plan (e.g. the 4096/1024 one of _spectral_balance)
# only runs the STFT-feature stage plus what was asked for.
_LIGHT_OUTPUTS = ("ltas", "centroid", "rolloff_bin", "frame_max", "band_energy")
_CQT_OUTPUTS = ("chroma_cqt", "cqt_tuning", "cqt_mag")


def get_plan(sample_rate: int, n_fft: int = 2048, hop: int = 512, n_mels: int = 128, *, roll_percent: float = 0.85,
             meter_block: float = 0.4, device: int | None = None, window: str = "hann") -> engine.Plan:
    """``window``: a scipy.signal.get_window name, as librosa.stft(window=...) takes it (features.py:66-79)."""
    import torch

    dev = torch.cuda.current_device() if (device is None and torch.cuda.is_available()) else device
    key = (dev, int(sample_rate), int(n_fft), int(hop), int(n_mels), float(roll_percent), float(meter_block), str(window))
    with _plans_lock:
        plan = _plans.get(key)
        if plan is None:
            table = None
            if window != "hann":
                import scipy.signal

                table = scipy.signal.get_window(window, int(n_fft), fftbins=True)
            plan = engine.Plan(sample_rate, n_fft, hop, n_mels, device=dev, roll_percent=roll_percent,
                               meter_block=meter_block, window=table)
            _plans[key] = plan
            while len(_plans) > _MAX_PLANS:
                _plans.popitem(last=False)  # dropped here; the plan closes when its last user lets go of it
        else:
            _plans.move_to_end(key)
    return plan


def _fingerprint(x: np.ndarray):
    """Identity of a sample buffer inside one session: address, layout and a strided probe of its contents.  The session
    keeps a reference to every buffer it has fingerprinted (``_pin``), so a temporary cannot be freed and its address
    reused by different data while the session lives; buffers must not be modified in place inside a session."""
    x = np.asarray(x)
    probe = x.reshape(-1)[:: max(1, x.size // 4096)]
    return (x.__array_interface__["data"][0], x.shape, x.strides, str(x.dtype), hash(probe.tobytes()))


def _pin(cache: dict, fp, x: np.ndarray) -> None:
    cache.setdefault("__pinned__", {})[fp] = x


@contextlib.contextmanager
def frontend_session():
    """Scope inside which identical frontend requests share one GPU computation."""
    prev = getattr(_local, "cache", None)
    _local.cache = {} if prev is None else prev
    try:
        yield _local.cache
    finally:
        _local.cache = prev


def alias_mono_to_stereo(mono: np.ndarray, stereo: np.ndarray) -> bool:
    """Inside a session: let requests for ``mono`` be served by the fused run on ``stereo``.

    Valid when mono is exactly the float32 mean of the two channels (utils.coerce_audio builds it that way,
    utils.py:116): the stereo kernel's mid spectrum IS the mono spectrum, its K-weighting and true peak run on
    the same mono mix.  Returns False (and does nothing) when that does not hold."""
    cache = getattr(_local, "cache", None)
    mono = np.asarray(mono)
    stereo = np.asarray(stereo)
    if cache is None or stereo.ndim != 2 or stereo.shape[0] != 2 or mono.shape != stereo.shape[1:]:
        return False
    if mono.dtype != np.float32 or stereo.dtype != np.float32:
        return False
    # np.mean over two float32 rows is (left + right) / 2 in float32; halving is exact, so it is compared as
    # (left + right) * 0.5 in cache-sized pieces instead of materialising the full mean
    tmp = np.empty(min(mono.shape[0], 1 << 18), dtype=np.float32)
    for a in range(0, mono.shape[0], tmp.shape[0]):
        b = min(mono.shape[0], a + tmp.shape[0])
        t = tmp[: b - a]
        np.add(stereo[0, a:b], stereo[1, a:b], out=t)
        t *= np.float32(0.5)
        if not np.array_equal(t, mono[a:b]):
            return False
    cache.setdefault("__alias__", {})[_fingerprint(mono)] = _fingerprint(stereo)
    cache.setdefault("__alias_buf__", {})[_fingerprint(stereo)] = stereo
    _pin(cache, _fingerprint(mono), mono)
    return True


@contextlib.contextmanager
def precomputed_session(results: dict):
    """Scope in which every ``frontend()`` request is answered from ``results`` and nothing touches the GPU.

    ``results``: ``{(n_fft, hop, n_mels, channels): TrackResult}`` of ONE track, produced by a batched run elsewhere
    (``pipeline.analyse_tracks`` runs the kernels on whole chunks in the parent and the per-track host stages in worker
    processes).  A request for the mono view of a stereo track falls back to the stereo entry when there is no mono one
    (mono == mid exactly, utils.py:116).  Requests outside what was precomputed raise instead of computing."""
    prev = getattr(_local, "precomputed", None)
    _local.precomputed = results
    try:
        yield results
    finally:
        _local.precomputed = prev


def is_precomputed() -> bool:
    """True inside a ``precomputed_session`` (the caller must not create plans or touch the GPU)."""
    return getattr(_local, "precomputed", None) is not None


def _from_precomputed(pre: dict, x: np.ndarray, n_fft, hop, n_mels, roll_percent, meter_block, window, want):
    if float(roll_percent) != 0.85 or float(meter_block) != 0.4 or window != "hann":
        raise RuntimeError("precomputed session: only the default roll_percent / meter_block / window are available")
    ch = 2 if (x.ndim == 2 and x.shape[0] == 2) else 1
    hit = pre.get((int(n_fft), int(hop), int(n_mels), ch))
    if hit is None and ch == 1:
        hit = pre.get((int(n_fft), int(hop), int(n_mels), 2))
    if hit is None:
        raise RuntimeError(f"precomputed session: no result for n_fft={n_fft} hop={hop} n_mels={n_mels} channels={ch}")
    missing = [o for o in (want or ()) if o not in hit]
    if missing:
        raise RuntimeError(f"precomputed session: outputs {missing} were not computed for n_fft={n_fft} hop={hop}")
    return hit


def _plan_outputs(plan: engine.Plan, n_mels: int, n_samples: int, outs) -> tuple:
    """Drop what this plan / track cannot produce from a requested output set."""
    outs = tuple(outs)
    if n_samples < plan.meter_block * plan.sample_rate:
        outs = tuple(o for o in outs if o not in ("kw_blocks", "lufs"))
    if n_mels == 0:
        outs = tuple(o for o in outs if o not in ("mel", "onset_env", "autocorr", "flux_linear", "tempogram", "mfcc", "self_similarity"))
    if not plan.cqt_ok:
        outs = tuple(o for o in outs if o not in _CQT_OUTPUTS)
    return outs


def frontend(samples: np.ndarray, sample_rate: int, *, n_fft: int = 2048, hop: int = 512, n_mels: int = 128,
             roll_percent: float = 0.85, meter_block: float = 0.4, outputs=None, window: str = "hann") -> engine.TrackResult:
    """Run (or fetch) the fused frontend for one track given as (N,), (1, N) or (2, N) float32."""
    x = np.asarray(samples, dtype=np.float32)
    if x.ndim == 2 and x.shape[0] == 1:
        x = x[0]
    cache = getattr(_local, "cache", None)
    want = tuple(outputs) if outputs is not None else None
    pre = getattr(_local, "precomputed", None)
    if pre is not None:
        return _from_precomputed(pre, x, n_fft, hop, n_mels, roll_percent, meter_block, window, want)
    plan = get_plan(sample_rate, n_fft, hop, n_mels, roll_percent=roll_percent, meter_block=meter_block, window=window)
    everything = tuple(o for o in engine.ALL_OUTPUTS if o != "cqt_mag")
    key = None
    if cache is not None:
        fp = _fingerprint(x)
        _pin(cache, fp, x)
        src = None
        if fp in cache.get("__alias__", {}):   # mono view of a stereo buffer that is (or will be) analysed
            src = cache["__alias__"][fp]
            fp = src
        key = (fp, int(sample_rate), n_fft, hop, n_mels, float(roll_percent), float(meter_block), str(window))
        hit = cache.get(key)
        if hit is None and src is not None:
            x = cache["__alias_buf__"][src]   # run on the stereo buffer; the result also serves the mono requests
        if hit is not None and (want is None or all(o in hit for o in _plan_outputs(plan, n_mels, x.shape[-1], want))):
            return hit
        default_plan = (n_fft, hop, n_mels, window) == (2048, 512, 128, "hann")
        if want is None or default_plan or hit is not None:
            outs = everything + (tuple(want) if want else ())   # one run serves every later request of the session
        else:
            outs = tuple(dict.fromkeys(_LIGHT_OUTPUTS + want))
    else:
        outs = everything if want is None else want
    outs = _plan_outputs(plan, n_mels, x.shape[-1], outs)
    resident = None
    if cache is not None:  # the PCM of this buffer may already be in HBM for another plan of the same session
        pcm_key = ("__pcm__", _fingerprint(x), plan.device)
        resident = cache.get(pcm_key)
        if resident is None:
            resident = cache[pcm_key] = engine.upload(plan, [x])
    res = engine.analyse_batch(plan, [x], outs, lazy=cache is not None, resident=resident)[0]
    if cache is not None:
        cache[key] = res
    return res
