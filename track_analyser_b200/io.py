"""Minimal WAV decoding (mirror of ``io.load_audio``'s return contract, io.py:56-139).

Container parsing stays on the host; sample conversion has a device
kernel (``engine.decode_pcm``) and so has resampling (``resample.resample``).  This reader exists so
``analyse_track(path)`` works on PCM16/24/32 and float32 WAV files.
Returns ``(samples, sample_rate, metadata)`` like the reference (io.py:56-139): with ``mono=True`` (the
default) a 1-d float32 array (the channel mean), with ``mono=False`` always planar ``(channels, N)`` --
``(1, N)`` for a mono file, which is what makes ``coerce_audio`` attach ``stereo_samples`` of shape
``(1, N)`` to mono files (utils.py:106-108).
"""

from __future__ import annotations

import struct

import numpy as np


def load_audio(path: str, target_sr=None, mono: bool = True):
    with open(path, "rb") as fh:
        raw = fh.read()
    if raw[:4] != b"RIFF" or raw[8:12] != b"WAVE":
        raise RuntimeError(f"Could not decode audio file {path}: only RIFF/WAVE is supported")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(raw):
        cid, size = raw[pos:pos + 4], struct.unpack("<I", raw[pos + 4:pos + 8])[0]
        body = raw[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
            fmt_body = body
        elif cid == b"data":
            data = body
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise RuntimeError(f"Could not decode audio file {path}: missing fmt/data chunk")
    tag, channels, sr, _, _, bits = fmt
    if tag == 0xFFFE:  # WAVE_FORMAT_EXTENSIBLE: the first two bytes of the SubFormat GUID are the real format tag
        if len(fmt_body) < 26:
            raise RuntimeError(f"Could not decode audio file {path}: truncated WAVE_FORMAT_EXTENSIBLE header")
        tag = struct.unpack("<H", fmt_body[24:26])[0]
    if channels < 1:
        raise RuntimeError(f"Could not decode audio file {path}: no channels")
    width = bits // 8
    data = data[: (len(data) // (width * channels)) * width * channels] if width else data  # whole frames only
    if tag == 1 and bits == 16:
        x = np.frombuffer(data, dtype="<i2").astype(np.float32) / 32768.0
    elif tag == 1 and bits == 24:
        b = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v - (1 << 24), v)
        x = v.astype(np.float32) / 8388608.0
    elif tag == 1 and bits == 32:
        x = (np.frombuffer(data, dtype="<i4").astype(np.float64) / 2147483648.0).astype(np.float32)
    elif tag == 3 and bits == 32:
        x = np.frombuffer(data, dtype="<f4").astype(np.float32)
    else:
        raise RuntimeError(f"Could not decode audio file {path}: unsupported WAV format {tag}/{bits}")
    x = x[: (x.size // channels) * channels].reshape(-1, channels).T
    x = np.ascontiguousarray(x, dtype=np.float32)
    if target_sr is not None and int(target_sr) != int(sr):  # io.py:126-128
        from .resample import resample

        x = resample(x, int(sr), int(target_sr))
        sr = int(target_sr)
    if mono:  # io.py:129-138: mean over channels, squeezed to 1-d
        samples = np.ascontiguousarray(x[0] if channels == 1 else np.mean(x, axis=0), dtype=np.float32)
    else:
        samples = x
    meta = {"path": path, "sample_rate": int(sr), "channels": int(channels), "frames": int(x.shape[1]),
            "duration": float(x.shape[1]) / float(sr)}
    return samples, int(sr), meta
