"""Sample-rate conversion on the device (mirror of the reference's ``utils._resample``, utils.py:55-70).

The reference converts every file input (and every ``AudioInput`` / ``(samples, sr)`` source whose rate differs from
the target) to 44.1 kHz with ``resampy.resample(x, sr_orig, sr_new)``: band-limited sinc interpolation with the
"kaiser_best" window, one call per channel.  Here the interpolation runs in ``csrc/resample.cu`` behind
``ta_resample``; this module builds the window the way ``resampy.filters.sinc_window`` does -- resampy ships the same
array as a data file -- and owns one resampler handle per (device, rate pair).

resampy is not installed in this image and the reference has no golden vectors for it, so this path is checked
bit for bit against ``oracle/resampy_np.py`` (a restatement of resampy 0.4's published algorithm), not against
resampy itself: see DESIGN.md ("parity unpinned" for this row).
"""

from __future__ import annotations

import ctypes as C
import functools

import numpy as np
import scipy.signal
import torch

from . import _native as nat

# resampy's "kaiser_best" design: 64 zero crossings, 2**9 table samples per crossing, roll-off and Kaiser beta as published
KAISER_BEST = dict(num_zeros=64, precision=9, rolloff=0.9475937167399596, beta=14.769656459379492)


@functools.lru_cache(maxsize=None)
def kaiser_best_window():
    """(half window float64 [num_zeros * 2**precision + 1], table samples per zero crossing)."""
    num_table = 2 ** KAISER_BEST["precision"]
    n = num_table * KAISER_BEST["num_zeros"]
    rolloff = KAISER_BEST["rolloff"]
    sinc_win = rolloff * np.sinc(rolloff * np.linspace(0, KAISER_BEST["num_zeros"], num=n + 1, endpoint=True))
    taper = scipy.signal.windows.kaiser(2 * n + 1, KAISER_BEST["beta"])[n:]
    return np.ascontiguousarray(taper * sinc_win, dtype=np.float64), num_table


class Resampler:
    """Device resampler for one rate pair (``ta_resampler``)."""

    def __init__(self, sr_orig: int, sr_new: int, device: int | None = None):
        if sr_orig <= 0 or sr_new <= 0:
            raise ValueError("Invalid sample rate")
        self.lib = nat.load()
        self.device = torch.cuda.current_device() if device is None else int(device)
        self.sr_orig, self.sr_new = int(sr_orig), int(sr_new)
        win, num_table = kaiser_best_window()
        self._h = C.c_void_p()
        nat.check(self.lib.ta_resampler_create(self.device, self.sr_orig, self.sr_new,
                                               win.ctypes.data_as(C.POINTER(C.c_double)), win.shape[0], num_table,
                                               C.byref(self._h)))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.ta_resampler_destroy(h)

    def out_len(self, n_in: int) -> int:
        return int(self.lib.ta_resampler_out_len(self._h, int(n_in)))

    def run_device(self, x: torch.Tensor) -> torch.Tensor:
        """(rows, n) or (n,) float32 CUDA tensor -> resampled tensor of the same rank on the same device."""
        assert x.is_cuda and x.dtype == torch.float32
        rows = x.reshape(-1, x.shape[-1]).contiguous()
        n_out = self.out_len(rows.shape[1])
        if n_out < 1:
            raise ValueError(f"Input signal length={rows.shape[1]} is too small to resample from "
                             f"{self.sr_orig}->{self.sr_new}")
        out = torch.empty((rows.shape[0], n_out), dtype=torch.float32, device=x.device)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        nat.check(self.lib.ta_resample(self._h, C.c_void_p(rows.data_ptr()), rows.shape[1], rows.shape[1], rows.shape[0],
                                       C.c_void_p(out.data_ptr()), n_out, stream))
        return out.reshape(x.shape[:-1] + (n_out,))


_resamplers: dict = {}


def get_resampler(sr_orig: int, sr_new: int, device: int | None = None) -> Resampler:
    dev = torch.cuda.current_device() if device is None else int(device)
    key = (dev, int(sr_orig), int(sr_new))
    if key not in _resamplers:
        _resamplers[key] = Resampler(sr_orig, sr_new, dev)
    return _resamplers[key]


def resample(samples: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """``utils._resample`` (utils.py:55-70): host array in, host array out, (n,) or (channels, n); float32 result."""
    if orig_sr == target_sr:
        return samples
    x = np.asarray(samples)
    if np.issubdtype(x.dtype, np.integer) or x.dtype != np.float32:
        x = x.astype(np.float32)  # the reference hands float32 to resampy (io.py:79, utils.py:87,109,126,137)
    r = get_resampler(orig_sr, target_sr)
    dev = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(x))).to(f"cuda:{r.device}")
    out = r.run_device(dev).cpu().numpy()
    return out[0] if x.ndim == 1 else out
