"""numpy/scipy restatement of pyloudnorm 0.1.1 ``Meter.integrated_loudness``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  pyloudnorm is an un-vendored
dependency of the reference (pin: /root/reference/pyproject.toml:19); the
reference reaches it at analysis/loudness.py:60-61.  Restated from the published
algorithm: two cascaded RBJ-style biquads applied with
``scipy.signal.lfilter`` (float64 state) and written back into the float32
working copy between stages, then 400 ms / 75 %-overlap gated block energies.
"""

from __future__ import annotations

import numpy as np
import scipy.signal


def k_weighting_coefficients(rate):
    """[(b, a)] for the high-shelf (4 dB, Q 1/sqrt2, 1500 Hz) then high-pass (Q 0.5, 38 Hz)."""
    out = []
    # high shelf
    G, Q, fc = 4.0, 1.0 / np.sqrt(2.0), 1500.0
    A = 10 ** (G / 40.0)
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    b0 = A * ((A + 1) + (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha)
    b1 = -2 * A * ((A - 1) + (A + 1) * np.cos(w0))
    b2 = A * ((A + 1) + (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha)
    a0 = (A + 1) - (A - 1) * np.cos(w0) + 2 * np.sqrt(A) * alpha
    a1 = 2 * ((A - 1) - (A + 1) * np.cos(w0))
    a2 = (A + 1) - (A - 1) * np.cos(w0) - 2 * np.sqrt(A) * alpha
    out.append((np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0))
    # high pass
    G, Q, fc = 0.0, 0.5, 38.0
    w0 = 2.0 * np.pi * (fc / rate)
    alpha = np.sin(w0) / (2.0 * Q)
    b0 = (1 + np.cos(w0)) / 2
    b1 = -(1 + np.cos(w0))
    b2 = (1 + np.cos(w0)) / 2
    a0 = 1 + alpha
    a1 = -2 * np.cos(w0)
    a2 = 1 - alpha
    out.append((np.array([b0, b1, b2]) / a0, np.array([a0, a1, a2]) / a0))
    return out


def k_weight(data, rate):
    """Filtered copy of mono ``data`` in its own dtype (float32 in the reference)."""
    x = np.array(data, copy=True)
    for b, a in k_weighting_coefficients(rate):
        x[:] = scipy.signal.lfilter(b, a, x)
    return x


def block_bounds(n_samples, rate, block_size=0.4):
    """(lower, upper) int arrays exactly as pyloudnorm evaluates them in Python floats."""
    T_g = block_size
    step = 0.25
    T = n_samples / rate
    num_blocks = int(np.round(((T - T_g) / (T_g * step))) + 1)
    lo = [int(T_g * (j * step) * rate) for j in range(num_blocks)]
    hi = [int(T_g * (j * step + 1) * rate) for j in range(num_blocks)]
    return np.asarray(lo, dtype=np.int64), np.asarray(hi, dtype=np.int64)


def gate(z, Gamma_a=-70.0):
    """BS.1770 absolute + relative gating of mono block energies ``z`` -> LUFS."""
    z = np.asarray(z, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        l = -0.691 + 10.0 * np.log10(z)
        J_g = [j for j, l_j in enumerate(l) if l_j >= Gamma_a]
        z_avg = np.mean([z[j] for j in J_g]) if J_g else np.nan
        Gamma_r = -0.691 + 10.0 * np.log10(z_avg) - 10.0
        J_g = [j for j, l_j in enumerate(l) if (l_j > Gamma_r and l_j > Gamma_a)]
        z_avg = np.nan_to_num(np.mean([z[j] for j in J_g]) if J_g else np.nan)
        return float(-0.691 + 10.0 * np.log10(z_avg))


def block_energies(data, rate, block_size=0.4):
    data = np.asarray(data)
    if data.ndim != 1:
        raise ValueError("oracle restates the mono path only")
    if data.shape[0] < block_size * rate:
        raise ValueError("Audio must have length greater than the block size.")
    x = k_weight(data, rate)
    lo, hi = block_bounds(x.shape[0], rate, block_size)
    z = np.zeros(len(lo))
    for j, (l, u) in enumerate(zip(lo, hi)):
        z[j] = (1.0 / (block_size * rate)) * np.sum(np.square(x[l:u]))
    return z


def integrated_loudness(data, rate, block_size=0.4):
    return gate(block_energies(data, rate, block_size))
