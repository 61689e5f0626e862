"""numpy restatement of ``resampy.resample(x, sr_orig, sr_new)`` (band-limited sinc interpolation, "kaiser_best").

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  resampy is an un-vendored dependency of the reference
(pins: /root/reference/pyproject.toml:16 ``resampy==0.4.3``, requirements.txt:4 ``resampy==0.4.2``); the reference
reaches it at utils.py:55-70 (``_resample``: one call per channel, default filter) from ``coerce_audio``
(utils.py:91-96, 112-116, 141-143) and ``load_audio`` (io.py:126-128).

PARITY UNPINNED: resampy is not installed in this image, the reference holds no golden vectors for its output
(tests/test_cli.py only checks that a resampled 22.05 kHz file produces reports), and the interpolation window is
rebuilt here from the filter's published design parameters instead of being read from resampy's packaged
``kaiser_best.npz``.  Sanity pin (tests/test_oracle_crosschecks.py): torchaudio's ``sinc_interp_kaiser`` with the
same design -- whose default Kaiser beta is this very number -- agrees to < 5e-4; band-limited interpolation
properties in tests/test_host_logic.py.  What is restated, from the published algorithm (J. O. Smith's band-limited interpolation
as implemented in resampy 0.4 ``core.resample`` / ``interpn.resample_f``) [UPSTREAM-RECALL]:

* window: ``sinc_window(num_zeros=64, precision=9, rolloff=0.9475937167399596)`` tapered by the right half of a
  Kaiser window with ``beta=14.769656459379492`` -- 64 * 512 + 1 = 32 769 float64 samples, 512 per zero crossing;
* ``sample_ratio = sr_new / sr_orig``; output length ``int(n * sample_ratio)``; the window is scaled by the ratio
  when down-sampling; ``interp_delta = diff(window)`` with a trailing 0; ``scale = min(1, ratio)``;
* per output sample t at time ``t * (1 / ratio)``: left wing over x[n], x[n-1], ... then right wing over x[n+1],
  x[n+2], ..., each tap weighted by the linearly interpolated window, accumulated IN PLACE in an output array of the
  input's dtype -- for float32 input every tap rounds the running sum to float32 (float64 weight * float32 sample,
  added in float64, stored as float32).

The loops over taps are vectorised over output samples; the accumulation order per output sample is the reference
implementation's.
"""

from __future__ import annotations

import functools

import numpy as np
import scipy.signal

KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)


def sinc_window(num_zeros, precision, rolloff, beta):
    """Right half of the windowed sinc, ``num_zeros * 2**precision + 1`` samples (resampy.filters.sinc_window)."""
    num_bits = 2 ** precision
    n = num_bits * num_zeros
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, num_zeros, num=n + 1, endpoint=True))
    taper = scipy.signal.windows.kaiser(2 * n + 1, beta)[n:]
    return taper * sinc_win, num_bits


@functools.lru_cache(maxsize=None)
def kaiser_best():
    win, num_table = sinc_window(**KAISER_BEST)
    win.setflags(write=False)
    return win, num_table


def output_length(n: int, sr_orig: int, sr_new: int) -> int:
    return int(n * (float(sr_new) / float(sr_orig)))


def resample(x: np.ndarray, sr_orig: int, sr_new: int) -> np.ndarray:
    """``resampy.resample(x, sr_orig, sr_new)`` along the last axis (1-d or (rows, n) input)."""
    x = np.asarray(x)
    if sr_orig <= 0 or sr_new <= 0:
        raise ValueError("Invalid sample rate")
    sample_ratio = float(sr_new) / float(sr_orig)
    n_out = int(x.shape[-1] * sample_ratio)
    if n_out < 1:
        raise ValueError(f"Input signal length={x.shape[-1]} is too small to resample from {sr_orig}->{sr_new}")
    dtype = np.float32 if np.issubdtype(x.dtype, np.integer) else x.dtype
    if x.ndim > 1:
        return np.stack([resample(row, sr_orig, sr_new) for row in x.reshape(-1, x.shape[-1])]).reshape(
            x.shape[:-1] + (n_out,)).astype(dtype, copy=False)
    interp_win, num_table = kaiser_best()
    if sample_ratio < 1:
        interp_win = interp_win * sample_ratio
    interp_delta = np.diff(interp_win, append=interp_win[-1])
    scale = min(1.0, sample_ratio)
    t_out = np.arange(n_out) * (1.0 / sample_ratio)
    index_step = int(scale * num_table)
    nwin, n_orig = interp_win.shape[0], x.shape[0]
    xd = x.astype(np.float64)
    y = np.zeros(n_out, dtype=dtype)

    n = t_out.astype(np.int64)  # int(time_register): truncation of a non-negative float
    frac = scale * (t_out - n)

    def wing(frac, count, sample_index):
        nonlocal y
        index_frac = frac * num_table
        offset = index_frac.astype(np.int64)
        eta = index_frac - offset
        for i in range(int(count.max()) if count.size else 0):
            live = np.nonzero(i < count)[0]
            j = offset[live] + i * index_step
            weight = interp_win[j] + eta[live] * interp_delta[j]
            y[live] = (y[live].astype(np.float64) + weight * xd[sample_index(live, i)]).astype(dtype)

    i_max = np.minimum(n + 1, (nwin - (frac * num_table).astype(np.int64)) // index_step)
    wing(frac, i_max, lambda live, i: n[live] - i)
    frac_r = scale - frac
    k_max = np.minimum(n_orig - n - 1, (nwin - (frac_r * num_table).astype(np.int64)) // index_step)
    wing(frac_r, k_max, lambda live, k: n[live] + k + 1)
    return y
