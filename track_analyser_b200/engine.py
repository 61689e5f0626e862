"""Device-side driver: plans, batch packing and the fused frontend call.

PyTorch is used only for memory ownership (pinned host staging, device buffers)
and for the current CUDA stream; every number is produced by the kernels in
``csrc/`` through the C ABI (``include/ta_b200.h``).  Tracks are independent
(reference: pipeline.py:32 takes one source), so a batch is a ragged list of
planar float32 arrays processed by one launch sequence.
"""

from __future__ import annotations

import ctypes as C
import os
import sys
import threading
import time
from dataclasses import dataclass, field
from typing import Iterable, Sequence

import numpy as np
import torch

from . import _native as nat

_TRACE_FETCH = os.environ.get("TA_TRACE_FETCH", "0") not in ("", "0")  # debugging aid: log every lazy D2H fetch

ALL_OUTPUTS = (
    "magnitude", "mel", "onset_env", "autocorr", "flux_linear", "ltas", "centroid", "rolloff_bin",
    "band_energy", "moments", "kw_blocks", "lufs", "rms_momentary", "rms_short", "frame_max", "chroma", "tuning",
    "tempogram", "true_peak", "hpss_harmonic", "hpss_percussive", "mfcc", "self_similarity", "chroma_cqt", "cqt_tuning", "cqt_mag",
)
# SURVEY section 8a (the north-star frontend): what bench.py measures.  true_peak, the HPSS curves and the MFCC are
# section-8f "next" rows.
_NEXT_ROWS = ("true_peak", "hpss_harmonic", "hpss_percussive", "mfcc", "self_similarity", "chroma_cqt", "cqt_tuning", "cqt_mag")
FRONTEND_OUTPUTS = tuple(o for o in ALL_OUTPUTS if o not in _NEXT_ROWS)
# What the host-side stages of pipeline.analyse_track consume: neither the magnitude (HPSS runs on the device), nor the
# mel matrix (its only host consumer, the MFCC, runs on the device) nor the plot-only tempogram leave the GPU.
ANALYSIS_OUTPUTS = tuple(o for o in ALL_OUTPUTS if o not in ("magnitude", "mel", "tempogram", "cqt_mag", "mfcc"))
# the bench / default frontend: everything the per-track analysis consumes except the plot-only tempogram (and the
# constant-Q chroma, which only the 2048/512 plan of the harmony stage can produce: ask for it explicitly)
CORE_OUTPUTS = tuple(o for o in ALL_OUTPUTS if o not in ("tempogram", "chroma_cqt", "cqt_tuning", "cqt_mag"))
DEFAULT_OUTPUTS = tuple(o for o in CORE_OUTPUTS if o != "magnitude")


N_CQT_BINS = 252  # 7 octaves x 36 bins (librosa.feature.chroma_cqt defaults)
N_MFCC = 13     # TA_N_MFCC
N_MOMENTS = 10  # TA_N_MOMENTS: sum L, R, L^2, R^2, LR, mid^2, side^2, n, sum |L|, sum |R|
STAGE_NAMES = ("stft_mel_features", "onset_flux", "autocorrelation", "tempogram", "chroma_stft", "time_domain_loudness")


def available_outputs(plan: "Plan", outputs: Iterable[str] = ALL_OUTPUTS) -> tuple:
    """``outputs`` without what ``plan`` cannot produce: the mel-derived ones when it has no mel bands, the constant-Q
    ones unless it is librosa's chroma_cqt configuration (n_fft 2048, hop 512, basis below the Nyquist frequency)."""
    drop = set()
    if plan.n_mels == 0:
        drop |= {"mel", "onset_env", "autocorr", "flux_linear", "tempogram", "mfcc", "self_similarity"}
    if not plan.cqt_ok:
        drop |= {"chroma_cqt", "cqt_tuning", "cqt_mag"}
    return tuple(o for o in outputs if o not in drop)


def frame_count(n_samples: int, hop: int) -> int:
    return 1 + n_samples // hop


def frame_pitch(n_frames: int) -> int:
    return (n_frames + 31) & ~31


@dataclass
class TrackResult:
    """Per-track host copies of the frontend outputs (numpy, reference shapes).

    With a ``loader`` the device buffers stay alive and an output is copied to the host the first time it is
    asked for (``runtime.frontend_session``: one fused run serves many consumers, most of which want KB-sized
    outputs, not the 64 MB magnitude)."""

    n_samples: int
    n_frames: int
    data: dict = field(default_factory=dict)
    channels: int = 1
    loader: object = None   # callable(key) -> numpy value, or None
    available: tuple = ()

    def __getitem__(self, key):
        if key not in self.data:
            if self.loader is None or key not in self.available:
                raise KeyError(key)
            self.data[key] = self.loader(key)
        return self.data[key]

    def __contains__(self, key):
        return key in self.data or key in self.available


class PlanGeometry:
    """The host-side numbers of a plan that the result layout depends on (no device state): what a worker process needs to
    cut a batch's output arrays into per-track views."""

    def __init__(self, sample_rate: int, n_fft: int, hop: int, n_mels: int, meter_block: float = 0.4, tempogram_win: int = 384):
        self.sample_rate, self.n_fft, self.hop, self.n_mels = int(sample_rate), int(n_fft), int(hop), int(n_mels)
        self.meter_block, self.tempogram_win = float(meter_block), int(tempogram_win)
        self.n_bins = self.n_fft // 2 + 1

    def geometry(self) -> "PlanGeometry":
        return PlanGeometry(self.sample_rate, self.n_fft, self.hop, self.n_mels, self.meter_block, self.tempogram_win)

    # ---- loudness framing helpers ---------------------------------------------------
    def rms_frames(self, seconds: float) -> tuple[int, int]:
        frame = max(1024, int(round(self.sample_rate * seconds)))
        if frame % 2:
            frame += 1
        return frame, max(1, frame // 2)

    def kw_block_count(self, n_samples: int) -> int:
        T_g = self.meter_block
        if n_samples < T_g * self.sample_rate:
            return 0
        return int(np.round(((n_samples / self.sample_rate - T_g) / (T_g * 0.25))) + 1)


class BatchGeometry:
    """Frame counts and pitches of a batch (host arrays only): the other half of what ``_cut`` needs."""

    def __init__(self, n_samples, n_frames, pitch, pitch_off, channels, cqt=None):
        self.n_samples, self.n_frames, self.pitch, self.pitch_off = n_samples, n_frames, pitch, pitch_off
        self.channels, self.n_tracks, self._cqt = channels, len(n_samples), cqt

    def cqt_layout(self):
        return self._cqt


class Plan(PlanGeometry):
    """Owns a ``ta_plan`` (twiddles, window, mel filterbank, biquads) on one device."""

    def __init__(self, sample_rate: int, n_fft: int = 2048, hop: int = 512, n_mels: int = 128,
                 device: int | None = None, fmin: float = 0.0, fmax: float | None = None,
                 roll_percent: float = 0.85, meter_block: float = 0.4, n_chroma: int = 12,
                 tempogram_win: int = 384, window=None):
        """``window``: None (periodic Hann) or n_fft float64 values, e.g. scipy.signal.get_window(name, n_fft, fftbins=True)."""
        self.tempogram_win = int(tempogram_win)
        if not torch.cuda.is_available():
            raise RuntimeError("track_analyser_b200 needs a CUDA device (B200); there is no CPU fallback")
        self.lib = nat.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.sample_rate, self.n_fft, self.hop, self.n_mels = int(sample_rate), int(n_fft), int(hop), int(n_mels)
        self.roll_percent, self.meter_block = float(roll_percent), float(meter_block)
        desc = nat.PlanDesc(self.device, self.sample_rate, self.n_fft, self.hop, self.n_mels, int(n_chroma),
                            int(tempogram_win), 0, float(fmin), float(fmax) if fmax else 0.0,
                            self.roll_percent, self.meter_block)
        handle = C.c_void_p()
        if window is None:
            nat.check(self.lib.ta_plan_create(C.byref(desc), C.byref(handle)))
        else:
            w = np.ascontiguousarray(window, dtype=np.float64)
            if w.shape != (int(n_fft),):
                raise ValueError(f"window must hold n_fft = {n_fft} values")
            nat.check(self.lib.ta_plan_create_window(C.byref(desc), w.ctypes.data_as(C.POINTER(C.c_double)), C.byref(handle)))
        self._h = handle
        self.n_bins = self.n_fft // 2 + 1
        self._ws = None  # cached workspace tensor

    def close(self):
        if getattr(self, "_h", None):
            self.lib.ta_plan_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    # ---- tables (tests) -------------------------------------------------------
    def table(self, which: str) -> np.ndarray:
        spec = {"window": (0, np.float32, (self.n_fft,)), "mel": (1, np.float32, (self.n_mels, self.n_bins)),
                "freqs": (2, np.float64, (self.n_bins,)), "biquads": (3, np.float64, (2, 6))}[which]
        out = np.empty(spec[2], dtype=spec[1])
        nat.check(self.lib.ta_plan_table(self._h, spec[0], out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    # ---- constant-Q outputs ------------------------------------------------------
    @property
    def cqt_ok(self) -> bool:
        """Whether this plan can produce the constant-Q outputs (librosa's chroma_cqt defaults: n_fft 2048, hop 512, and a
        sample rate its basis fits under)."""
        if getattr(self, "_cqt_ok", None) is None:
            self._cqt_ok = self.n_fft == 2048 and self.hop == 512 and int(self.lib.ta_cqt_frame_count(self._h, 0)) >= 0
        return self._cqt_ok

    def cqt_frame_count(self, n_samples: int) -> int:
        """Frames of librosa.cqt / chroma_cqt for a track of ``n_samples`` (can differ by one from 1 + n // hop)."""
        n = int(self.lib.ta_cqt_frame_count(self._h, int(n_samples)))
        if n < 0:
            nat.check(n)
        return n


class DeviceBatch:
    """A ragged batch of tracks resident in HBM plus its host metadata."""

    def __init__(self, plan: Plan, pcm: torch.Tensor, offsets: np.ndarray, n_samples: np.ndarray, channels: int):
        self.plan, self.pcm, self.channels = plan, pcm, channels
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.n_samples = np.ascontiguousarray(n_samples, dtype=np.int64)
        self.n_tracks = len(self.n_samples)
        self.n_frames = 1 + self.n_samples // plan.hop
        self.pitch = (self.n_frames + 31) & ~31
        self.pitch_off = np.concatenate([[0], np.cumsum(self.pitch)]).astype(np.int64)
        self.total_pitch = int(self.pitch_off[-1])
        self.c_batch = nat.Batch(self.n_tracks, channels, pcm.data_ptr() if hasattr(pcm, "data_ptr") else pcm.ctypes.data,
                                 self.offsets.ctypes.data_as(C.POINTER(C.c_int64)),
                                 self.n_samples.ctypes.data_as(C.POINTER(C.c_int64)))

    def cqt_layout(self):
        """(frames, pitch, pitch offsets) of the constant-Q outputs (their frame count is librosa's, see ta_cqt_frame_count)."""
        if getattr(self, "_cqt", None) is None:
            frames = np.asarray([self.plan.cqt_frame_count(int(n)) for n in self.n_samples], dtype=np.int64)
            pitch = (frames + 31) & ~31
            self._cqt = (frames, pitch, np.concatenate([[0], np.cumsum(pitch)]).astype(np.int64))
        return self._cqt

    def geometry(self, with_cqt: bool = False) -> BatchGeometry:
        return BatchGeometry(self.n_samples, self.n_frames, self.pitch, self.pitch_off, self.channels,
                             self.cqt_layout() if with_cqt else None)

    def rebind(self, plan: Plan) -> "DeviceBatch":
        """The same resident PCM described for another plan (frame counts and pitches follow the plan's hop)."""
        if plan is self.plan:
            return self
        if plan.device != self.plan.device:
            raise ValueError("the batch lives on another device than the plan")
        return DeviceBatch(plan, self.pcm, self.offsets, self.n_samples, self.channels)


_staging: dict = {}
_staging_lock = threading.Lock()


def _pinned_staging(n: int) -> torch.Tensor:
    """Reusable pinned host buffer (cudaHostAlloc of tens of MB costs more than the copy it speeds up)."""
    buf = _staging.get("buf")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1 << 20), dtype=torch.float32, pin_memory=True)
        _staging["buf"] = buf
    return buf[:n]


def pack_host(tracks: Sequence[np.ndarray], channels: int, pinned: bool = True, reuse: bool = False):
    """Pack float32 tracks ((C, N) planar or (N,)) into one flat host tensor; offsets are 4-aligned.

    ``reuse``: stage through one process-wide pinned buffer (the caller must have finished with the previous
    contents, i.e. synchronise the copy before the next call)."""
    n_samples, offsets, total = [], [], 0
    for t in tracks:
        n = t.shape[-1]
        n_samples.append(n)
        offsets.append(total)
        total += (channels * n + 3) & ~3
    host = _pinned_staging(max(total, 4)) if (pinned and reuse) else torch.empty(max(total, 4), dtype=torch.float32, pin_memory=pinned)
    hv = host.numpy()
    for t, off, n in zip(tracks, offsets, n_samples):
        hv[off: off + channels * n] = np.ascontiguousarray(t, dtype=np.float32).reshape(-1)
    return host, np.asarray(offsets, dtype=np.int64), np.asarray(n_samples, dtype=np.int64)


_pack_pool = None
_COPY_PIECE = 2 << 20   # floats per host-to-device copy (8 MB)


def upload(plan: Plan, tracks: Sequence[np.ndarray], out: "torch.Tensor | None" = None, wait: bool = True) -> DeviceBatch:
    """Host tracks -> one flat device buffer (``out`` if it is a float32 tensor on the plan's device that is large enough,
    else a new one).  A track that already sits in pinned memory is copied from where it is; pageable ones go through the
    process-wide pinned staging buffer, filled by a few threads (numpy's copy releases the GIL and one core moves
    ~12 GB/s, a fifth of what the link takes).  ``wait=False``: when every track is pinned, return with the copies
    enqueued on the current stream instead of waiting for them (the staging buffer is not involved then)."""
    global _pack_pool
    tracks = [np.asarray(t, dtype=np.float32) for t in tracks]
    chans = {1 if t.ndim == 1 else t.shape[0] for t in tracks}
    if len(chans) != 1 or next(iter(chans)) not in (1, 2):
        raise ValueError("a batch must hold tracks that are all mono (N,) / (1, N) or all stereo (2, N)")
    channels = next(iter(chans))
    n_samples = np.asarray([t.shape[-1] for t in tracks], dtype=np.int64)
    sizes = (channels * n_samples + 3) & ~3
    offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    total = int(sizes.sum())
    if out is not None and out.dtype == torch.float32 and out.is_cuda and out.device.index == plan.device and out.numel() >= max(total, 4):
        dev = out[: max(total, 4)]
    else:
        dev = torch.empty(max(total, 4), dtype=torch.float32, device=f"cuda:{plan.device}")
    flat = [torch.from_numpy(t.reshape(-1)) if t.flags.c_contiguous else None for t in tracks]
    pinned = [f is not None and f.numel() > 0 and f.is_pinned() for f in flat]
    with _staging_lock:  # one process-wide pinned staging buffer: fill, copy, and wait before anyone refills it
        pageable = [i for i, p in enumerate(pinned) if not p]
        if pageable:
            host = _pinned_staging(max(total, 4)).numpy()

            def fill(i):
                off, n = int(offsets[i]), channels * int(n_samples[i])
                host[off: off + n] = np.ascontiguousarray(tracks[i], dtype=np.float32).reshape(-1)

            if len(pageable) > 1:
                if _pack_pool is None:
                    import concurrent.futures as cf

                    _pack_pool = cf.ThreadPoolExecutor(max_workers=4)
                list(_pack_pool.map(fill, pageable))
            else:
                fill(pageable[0])
        for i, f in enumerate(flat):
            off, n = int(offsets[i]), channels * int(n_samples[i])
            if n == 0:
                continue
            src = f if pinned[i] else torch.from_numpy(host[off: off + n])
            # pieces of 8 MB: the small descriptor uploads of a frontend run on another stream queue behind whatever the
            # copy engine is busy with, and must not wait for a whole 64 MB track
            for a in range(0, n, _COPY_PIECE):
                b = min(n, a + _COPY_PIECE)
                dev[off + a: off + b].copy_(src[a:b], non_blocking=True)
        if wait or pageable:
            torch.cuda.current_stream(dev.device).synchronize()
    return DeviceBatch(plan, dev, offsets, n_samples, channels)


PCM_FORMATS = {"s16": (nat.TA_PCM_S16, 2), "s24": (nat.TA_PCM_S24, 3), "s32": (nat.TA_PCM_S32, 4), "f32": (nat.TA_PCM_F32, 4)}


def decode_pcm(raw: torch.Tensor, fmt: str, channels: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """Interleaved PCM bytes already on the device (a WAV data chunk: uint8 / int16 / int32 / float32 tensor) ->
    planar float32 ``(channels, n_frames)`` on the same device, by the library's decode kernel (io.py:72-79)."""
    code, width = PCM_FORMATS[fmt]
    nbytes = raw.numel() * raw.element_size()
    n_frames = nbytes // (width * channels)
    if out is None:
        out = torch.empty((channels, n_frames), dtype=torch.float32, device=raw.device)
    assert out.is_cuda and out.dtype == torch.float32 and out.numel() >= channels * n_frames
    stream = C.c_void_p(torch.cuda.current_stream(raw.device).cuda_stream)
    nat.check(nat.load().ta_decode_pcm(C.c_void_p(raw.data_ptr()), code, channels, n_frames, C.c_void_p(out.data_ptr()), stream))
    return out


class FrontendBuffers:
    """Device output buffers (torch-owned) for one DeviceBatch and the matching C struct."""

    def __init__(self, batch: DeviceBatch, outputs: Iterable[str], true_peak_oversample: int = 8):
        plan = batch.plan
        dev = batch.pcm.device
        P, nt, B, M = batch.total_pitch, batch.n_tracks, plan.n_bins, plan.n_mels
        outputs = set(outputs)
        unknown = outputs - set(ALL_OUTPUTS)
        if unknown:
            raise ValueError(f"unknown outputs {sorted(unknown)}")
        self.requested = tuple(o for o in ALL_OUTPUTS if o in outputs)  # what the caller asked for (the rest are intermediates)
        if "self_similarity" in outputs:
            outputs.add("mfcc")
        if outputs & {"onset_env", "autocorr", "flux_linear", "mfcc"}:
            outputs.add("mel")
        if outputs & {"autocorr", "tempogram"}:
            outputs.add("onset_env")
        if "chroma" in outputs:
            outputs |= {"magnitude", "frame_max", "tuning"}
        if "tuning" in outputs:
            outputs |= {"magnitude", "frame_max", "chroma"}
        if outputs & {"onset_env", "autocorr", "flux_linear", "tempogram"}:
            outputs.add("mel")
        if outputs & {"hpss_harmonic", "hpss_percussive"}:
            outputs |= {"hpss_harmonic", "hpss_percussive", "magnitude"}
        if "rolloff_bin" in outputs:
            outputs.add("magnitude")
        if outputs & {"chroma_cqt", "cqt_tuning", "cqt_mag"}:
            outputs |= {"chroma_cqt", "cqt_tuning", "magnitude", "frame_max"}
        Pc = int(batch.cqt_layout()[2][-1]) if "chroma_cqt" in outputs else 0
        max_ns = int(batch.n_samples.max()) if nt else 0
        self.kw_pitch = max(1, plan.kw_block_count(max_ns))
        self.rms_pitch = 1 + max_ns // plan.rms_frames(plan.meter_block)[1]
        shapes = {
            "magnitude": ((B * P,), torch.float32), "mel": ((M * P,), torch.float32),
            "onset_env": ((P,), torch.float32), "autocorr": ((P,), torch.float64),
            "flux_linear": ((P,), torch.float64), "ltas": ((nt, B), torch.float64),
            "centroid": ((P,), torch.float64), "rolloff_bin": ((P,), torch.int32),
            "band_energy": ((nt, 2, B), torch.float64), "moments": ((nt, N_MOMENTS), torch.float64),
            "kw_blocks": ((nt, self.kw_pitch), torch.float64), "lufs": ((nt,), torch.float64),
            "rms_momentary": ((nt, self.rms_pitch), torch.float64), "rms_short": ((nt, self.rms_pitch), torch.float64),
            "frame_max": ((P,), torch.float32), "chroma": ((12 * P,), torch.float32), "tuning": ((nt,), torch.float64),
            "tempogram": ((plan.tempogram_win * P,), torch.float32), "true_peak": ((nt,), torch.float32),
            "hpss_harmonic": ((P,), torch.float32), "hpss_percussive": ((P,), torch.float32),
            "mfcc": ((N_MFCC * P,), torch.float64), "self_similarity": ((P,), torch.float64),
            "chroma_cqt": ((12 * Pc,), torch.float32), "cqt_tuning": ((nt,), torch.float64),
            "cqt_mag": ((N_CQT_BINS * Pc,), torch.float32),
        }
        self.t = {k: torch.empty(shapes[k][0], dtype=shapes[k][1], device=dev) for k in ALL_OUTPUTS if k in outputs}
        self.c_out = nat.FrontendOut()
        for k in ALL_OUTPUTS:
            setattr(self.c_out, k, self.t[k].data_ptr() if k in self.t else None)
        # device-only scratch of the HPSS kernels (time-direction medians); never copied to the host
        self.hpss_scratch = torch.empty(B * P, dtype=torch.float32, device=dev) if "hpss_harmonic" in outputs else None
        self.c_out.hpss_scratch = self.hpss_scratch.data_ptr() if self.hpss_scratch is not None else None
        # device-only scratch of the constant-Q chain (decimated signals, peak lists, descriptors)
        self.cqt_scratch = None
        if "chroma_cqt" in outputs:
            need = plan.lib.ta_cqt_scratch_bytes(plan._h, C.byref(batch.c_batch))
            if need == 0:
                nat.check(nat.TA_ERR_UNSUPPORTED)
            self.cqt_scratch = torch.empty(need, dtype=torch.uint8, device=dev)
            self.c_out.cqt_scratch = self.cqt_scratch.data_ptr()
            self.c_out.cqt_scratch_bytes = need
        self.c_out.true_peak_oversample = int(true_peak_oversample)
        self.c_out.kw_pitch = self.kw_pitch
        self.c_out.rms_pitch = self.rms_pitch
        self.outputs = outputs

    def bytes_d2h(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.t.values())


def workspace(plan: Plan, batch: DeviceBatch) -> torch.Tensor:
    need = plan.lib.ta_workspace_bytes(plan._h, C.byref(batch.c_batch))
    if need == 0:
        nat.check(nat.TA_ERR_INVALID)
    if plan._ws is None or plan._ws.numel() < need or plan._ws.device != batch.pcm.device:
        plan._ws = torch.empty(need, dtype=torch.uint8, device=batch.pcm.device)
    return plan._ws


def run_device(plan: Plan, batch: DeviceBatch, bufs: FrontendBuffers, stage: str = "frontend") -> None:
    """Enqueue the fused frontend (or one stage) on torch's current stream."""
    ws = workspace(plan, batch)
    stream = C.c_void_p(torch.cuda.current_stream(batch.pcm.device).cuda_stream)
    fn = {"frontend": plan.lib.ta_frontend_run, "stft": plan.lib.ta_stft_features,
          "time": plan.lib.ta_time_domain}[stage]
    nat.check(fn(plan._h, C.byref(batch.c_batch), C.byref(bufs.c_out), C.c_void_p(ws.data_ptr()), ws.numel(), stream))


def run_device_profiled(plan: Plan, batch: DeviceBatch, bufs: FrontendBuffers) -> list[float]:
    """Fused frontend with per-stage device times (ms): see STAGE_NAMES."""
    ws = workspace(plan, batch)
    stream = C.c_void_p(torch.cuda.current_stream(batch.pcm.device).cuda_stream)
    ms = (C.c_float * len(STAGE_NAMES))()
    nat.check(plan.lib.ta_frontend_run_profiled(plan._h, C.byref(batch.c_batch), C.byref(bufs.c_out),
                                                C.c_void_p(ws.data_ptr()), ws.numel(), stream, ms))
    return list(ms)


def launch_count() -> int:
    return int(nat.load().ta_launch_count())


def _cut(plan: Plan, batch: DeviceBatch, i: int, k: str, h: np.ndarray):
    """Slice the host copy ``h`` of output ``k`` down to track ``i`` in the reference's shape."""
    B, M = plan.n_bins, plan.n_mels
    T, ld, po = int(batch.n_frames[i]), int(batch.pitch[i]), int(batch.pitch_off[i])
    ns = int(batch.n_samples[i])
    if k == "magnitude":
        return h[B * po: B * (po + ld)].reshape(B, ld)[:, :T]
    if k == "mel":
        return h[M * po: M * (po + ld)].reshape(M, ld)[:, :T]
    if k == "chroma":
        return h[12 * po: 12 * (po + ld)].reshape(12, ld)[:, :T]
    if k == "tempogram":
        W = plan.tempogram_win
        return h[W * po: W * (po + ld)].reshape(W, ld)[:, :T]
    if k == "mfcc":
        return h[N_MFCC * po: N_MFCC * (po + ld)].reshape(N_MFCC, ld)[:, :T]
    if k in ("chroma_cqt", "cqt_mag"):
        frames, pitch, off = batch.cqt_layout()
        rows, Tc, ldc, poc = (12 if k == "chroma_cqt" else N_CQT_BINS), int(frames[i]), int(pitch[i]), int(off[i])
        return h[rows * poc: rows * (poc + ldc)].reshape(rows, ldc)[:, :Tc]
    if k in ("tuning", "lufs", "true_peak", "cqt_tuning"):
        return float(h[i])
    if k in ("onset_env", "autocorr", "flux_linear", "centroid", "rolloff_bin", "frame_max", "hpss_harmonic",
             "hpss_percussive", "self_similarity"):
        return h[po: po + T]
    if k == "kw_blocks":
        return h[i, : plan.kw_block_count(ns)]
    if k == "rms_momentary":
        return h[i, : 1 + ns // plan.rms_frames(plan.meter_block)[1]]
    if k == "rms_short":
        return h[i, : 1 + ns // plan.rms_frames(3.0)[1]]
    if k == "ltas":
        return (h[i] / T).astype(np.float32)
    return h[i]


def download(batch: DeviceBatch, bufs: FrontendBuffers) -> list[TrackResult]:
    """Copy results to the host and cut them into per-track numpy arrays of reference shape."""
    plan = batch.plan
    # only what was asked for travels: buffers that exist because another output needs them (the magnitude behind the
    # chroma / HPSS / roll-off kernels, the mel behind the onset envelope) stay on the device
    host = {k: bufs.t[k].cpu().numpy() for k in bufs.requested}
    out = []
    for i in range(batch.n_tracks):
        r = TrackResult(n_samples=int(batch.n_samples[i]), n_frames=int(batch.n_frames[i]), channels=batch.channels)
        for k, h in host.items():
            r.data[k] = _cut(plan, batch, i, k, h)
        out.append(r)
    return out


def download_track(batch: DeviceBatch, bufs: FrontendBuffers, i: int, keys: Iterable[str]) -> TrackResult:
    """Host copy of a few outputs of ONE track of a resident batch (slices on the device first: the matrices of a large
    batch are gigabytes).  Supports the (rows, T) matrices and the per-frame series."""
    plan = batch.plan
    rows = {"magnitude": plan.n_bins, "mel": plan.n_mels, "chroma": 12, "tempogram": plan.tempogram_win, "mfcc": N_MFCC}
    T, ld, po = int(batch.n_frames[i]), int(batch.pitch[i]), int(batch.pitch_off[i])
    r = TrackResult(n_samples=int(batch.n_samples[i]), n_frames=T, channels=batch.channels)
    for k in keys:
        t = bufs.t[k]
        if k in rows:
            r.data[k] = t[rows[k] * po: rows[k] * (po + ld)].cpu().numpy().reshape(rows[k], ld)[:, :T]
        elif t.ndim == 1 and t.numel() == batch.total_pitch:
            r.data[k] = t[po: po + T].cpu().numpy()
        else:
            raise KeyError(f"download_track handles the (rows, T) matrices and the per-frame series, not {k!r}")
    return r


def lazy_results(batch: DeviceBatch, bufs: FrontendBuffers) -> list[TrackResult]:
    """Results whose outputs are copied to the host on first access (the device buffers stay referenced)."""
    plan = batch.plan
    host: dict = {}

    def fetch(k):
        if k not in host:
            t0 = time.perf_counter() if _TRACE_FETCH else 0.0
            host[k] = bufs.t[k].cpu().numpy()
            if _TRACE_FETCH:
                print(f"[ta fetch] {k}: {host[k].nbytes} B in {(time.perf_counter() - t0) * 1e3:.2f} ms (n_fft {plan.n_fft})",
                      file=sys.stderr)
        return host[k]

    out = []
    for i in range(batch.n_tracks):
        out.append(TrackResult(n_samples=int(batch.n_samples[i]), n_frames=int(batch.n_frames[i]), channels=batch.channels,
                               loader=(lambda k, i=i: _cut(plan, batch, i, k, fetch(k))), available=tuple(bufs.t)))
    return out


def analyse_batch(plan: Plan, tracks: Sequence[np.ndarray], outputs: Iterable[str] = DEFAULT_OUTPUTS,
                  lazy: bool = False, resident: DeviceBatch | None = None, true_peak_oversample: int = 8) -> list[TrackResult]:
    """Host arrays in, host results out: H2D copy, fused frontend, D2H copy (on first access with ``lazy``).

    ``resident``: a DeviceBatch that already holds exactly these tracks (uploaded for another plan); its PCM is
    reused instead of being packed and copied again."""
    batch = upload(plan, tracks) if resident is None else resident.rebind(plan)
    bufs = FrontendBuffers(batch, outputs, true_peak_oversample=true_peak_oversample)
    run_device(plan, batch, bufs)
    return lazy_results(batch, bufs) if lazy else download(batch, bufs)


def _output_specs(plan: Plan, batch: DeviceBatch, kw_pitch: int, rms_pitch: int, Pc: int) -> dict:
    """name -> (element count, numpy dtype) of every ta_frontend_out array (include/ta_b200.h)."""
    P, nt, B, M = batch.total_pitch, batch.n_tracks, plan.n_bins, plan.n_mels
    f32, f64, i32 = np.float32, np.float64, np.int32
    return {
        "magnitude": (B * P, f32), "mel": (M * P, f32), "onset_env": (P, f32), "autocorr": (P, f64), "flux_linear": (P, f64),
        "ltas": (nt * B, f64), "centroid": (P, f64), "rolloff_bin": (P, i32), "band_energy": (nt * 2 * B, f64),
        "moments": (nt * N_MOMENTS, f64), "kw_blocks": (nt * kw_pitch, f64), "lufs": (nt, f64),
        "rms_momentary": (nt * rms_pitch, f64), "rms_short": (nt * rms_pitch, f64), "frame_max": (P, f32),
        "chroma": (12 * P, f32), "tuning": (nt, f64), "tempogram": (plan.tempogram_win * P, f32), "true_peak": (nt, f32),
        "hpss_harmonic": (P, f32), "hpss_percussive": (P, f32), "mfcc": (N_MFCC * P, f64), "self_similarity": (P, f64),
        "chroma_cqt": (12 * Pc, f32), "cqt_tuning": (nt, f64), "cqt_mag": (N_CQT_BINS * Pc, f32),
    }


def analyse_host(plan: Plan, tracks: Sequence[np.ndarray], outputs: Iterable[str] = DEFAULT_OUTPUTS,
                 true_peak_oversample: int = 8) -> list[TrackResult]:
    """numpy in, numpy out through ``ta_frontend_run_host``: the library owns every device buffer of the call (no torch
    tensor is created here).  Same results as ``analyse_batch``; meant for callers that bind the C ABI without torch."""
    tracks = [np.asarray(t, dtype=np.float32) for t in tracks]
    chans = {1 if t.ndim == 1 else t.shape[0] for t in tracks}
    if len(chans) != 1 or next(iter(chans)) not in (1, 2):
        raise ValueError("a batch must hold tracks that are all mono (N,) / (1, N) or all stereo (2, N)")
    channels = next(iter(chans))
    host, offsets, n_samples = pack_host(tracks, channels, pinned=False)
    pcm = host.numpy()
    batch = DeviceBatch(plan, pcm, offsets, n_samples, channels)
    outputs = tuple(outputs)
    unknown = set(outputs) - set(ALL_OUTPUTS)
    if unknown:
        raise ValueError(f"unknown outputs {sorted(unknown)}")
    max_ns = int(batch.n_samples.max()) if batch.n_tracks else 0
    kw_pitch = max(1, plan.kw_block_count(max_ns))
    rms_pitch = 1 + max_ns // plan.rms_frames(plan.meter_block)[1]
    Pc = int(batch.cqt_layout()[2][-1]) if set(outputs) & {"chroma_cqt", "cqt_mag", "cqt_tuning"} else 0
    specs = _output_specs(plan, batch, kw_pitch, rms_pitch, Pc)
    host_out = {k: np.empty(specs[k][0], dtype=specs[k][1]) for k in outputs}
    c_out = nat.FrontendOut()
    for k, v in host_out.items():
        setattr(c_out, k, v.ctypes.data)
    c_out.kw_pitch, c_out.rms_pitch, c_out.true_peak_oversample = kw_pitch, rms_pitch, int(true_peak_oversample)
    nat.check(plan.lib.ta_frontend_run_host(plan._h, C.byref(batch.c_batch), C.byref(c_out), None))
    for k in ("ltas", "band_energy", "moments", "kw_blocks", "rms_momentary", "rms_short"):
        if k in host_out:
            host_out[k] = host_out[k].reshape(batch.n_tracks, *({"band_energy": (2, plan.n_bins)}.get(k, (-1,))))
    out = []
    for i in range(batch.n_tracks):
        r = TrackResult(n_samples=int(batch.n_samples[i]), n_frames=int(batch.n_frames[i]), channels=channels)
        for k, h in host_out.items():
            r.data[k] = _cut(plan, batch, i, k, h)
        out.append(r)
    return out


class HostPipeline:
    """Multi-buffered host -> HBM -> host streaming of equal-length tracks.

    The public end-to-end path for large batches: per chunk of ``chunk_tracks``
    tracks it copies pinned host PCM to the device on a copy stream, runs the fused
    frontend on a compute stream and copies every requested output back into
    pinned host buffers on a third stream, so PCIe transfers in both directions
    overlap the kernels of the neighbouring chunks.  ``n_buffers`` (default 3) sets
    of device / pinned buffers rotate, so the host only waits when it is a whole
    buffer set ahead of the device.
    """

    def __init__(self, plan: Plan, n_samples: int, channels: int, chunk_tracks: int,
                 outputs: Iterable[str] = ALL_OUTPUTS, pcm16: bool = False, n_buffers: int = 3):
        """``pcm16``: the host tracks are interleaved int16 PCM (what a 16-bit WAV file holds, n_samples * channels
        values each); they are copied as such -- half the PCIe bytes -- and converted to planar float32 by the decode
        kernel on the copy stream."""
        self.plan, self.n_samples, self.channels, self.chunk = plan, int(n_samples), int(channels), int(chunk_tracks)
        self.pcm16 = bool(pcm16)
        self.nbuf = nb = max(2, int(n_buffers))
        dev = torch.device(f"cuda:{plan.device}")
        self.dev_raw = ([torch.empty(self.chunk * channels * self.n_samples, dtype=torch.int16, device=dev) for _ in range(nb)]
                        if self.pcm16 else None)
        self.stride = (channels * self.n_samples + 3) & ~3
        offsets = np.arange(self.chunk, dtype=np.int64) * self.stride
        ns = np.full(self.chunk, self.n_samples, dtype=np.int64)
        self.dev_in = [torch.empty(self.chunk * self.stride, dtype=torch.float32, device=dev) for _ in range(nb)]
        self.batches = [DeviceBatch(plan, d, offsets, ns, channels) for d in self.dev_in]
        self.requested = tuple(outputs)
        self.bufs = [FrontendBuffers(b, outputs) for b in self.batches]
        # only what was asked for is copied back; buffers that exist because another output needs them (the magnitude
        # for chroma / HPSS, the mel for the onset envelope ...) stay on the device
        self.host_out = [{k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in bf.t.items()
                          if k in self.requested} for bf in self.bufs]
        self.s_copy, self.s_comp, self.s_out = (torch.cuda.Stream(dev) for _ in range(3))
        self.ev_h2d = [torch.cuda.Event() for _ in range(nb)]
        self.ev_comp = [torch.cuda.Event() for _ in range(nb)]
        self.ev_d2h = [torch.cuda.Event() for _ in range(nb)]
        self.h2d_bytes_per_track = channels * self.n_samples * (2 if self.pcm16 else 4)
        self.d2h_bytes_per_chunk = sum(t.numel() * t.element_size() for t in self.host_out[0].values())
        # partial final chunks reuse the full-size buffers with a shorter batch view
        self._ws = workspace(plan, self.batches[0])
        self._tails = []

    def run(self, host_tracks: Sequence[torch.Tensor], consume=None) -> int:
        """Process ``host_tracks`` (pinned 1-d float32 tensors of C*N samples each).

        ``consume(chunk_index, first_track, n_tracks, host_out_dict)`` is called once the
        chunk's results are in pinned host memory.  Returns the number of chunks.
        """
        n = len(host_tracks)
        nb = self.nbuf
        n_chunks = (n + self.chunk - 1) // self.chunk
        pending = [None] * nb
        cur = torch.cuda.current_stream()
        for s in (self.s_copy, self.s_comp, self.s_out):
            s.wait_stream(cur)

        def finish(b):
            if pending[b] is not None:
                self.ev_d2h[b].synchronize()
                if consume is not None:
                    consume(*pending[b], self.host_out[b])
                pending[b] = None

        for ci in range(n_chunks):
            b = ci % nb
            finish(b)  # the host has consumed buffer b's previous results
            first = ci * self.chunk
            cnt = min(self.chunk, n - first)
            with torch.cuda.stream(self.s_copy):
                self.s_copy.wait_event(self.ev_comp[b])
                per = self.channels * self.n_samples
                for j in range(cnt):
                    if self.pcm16:
                        raw = self.dev_raw[b][j * per: (j + 1) * per]
                        raw.copy_(host_tracks[first + j], non_blocking=True)
                        decode_pcm(raw, "s16", self.channels, out=self.dev_in[b][j * self.stride: j * self.stride + per])
                    else:
                        self.dev_in[b][j * self.stride: j * self.stride + host_tracks[first + j].numel()].copy_(
                            host_tracks[first + j], non_blocking=True)
                self.ev_h2d[b].record(self.s_copy)
            with torch.cuda.stream(self.s_comp):
                self.s_comp.wait_event(self.ev_h2d[b])
                self.s_comp.wait_event(self.ev_d2h[b])
                batch = self.batches[b]
                if cnt != self.chunk:
                    batch = DeviceBatch(self.plan, self.dev_in[b], batch.offsets[:cnt], batch.n_samples[:cnt], self.channels)
                    self._tails.append(batch)  # keep host metadata alive until the stream has consumed it
                run_device(self.plan, batch, self.bufs[b])
                self.ev_comp[b].record(self.s_comp)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(self.ev_comp[b])
                for k, h in self.host_out[b].items():
                    h.copy_(self.bufs[b].t[k], non_blocking=True)
                self.ev_d2h[b].record(self.s_out)
            pending[b] = (ci, first, cnt)
        for k in range(nb):   # drain in submission order
            finish((n_chunks + k) % nb)
        cur.wait_stream(self.s_out)
        self._tails.clear()
        return n_chunks
