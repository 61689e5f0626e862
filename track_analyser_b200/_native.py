"""ctypes binding of ``libta_b200.so`` (the C ABI declared in ``include/ta_b200.h``).

There is no CPU fallback: importing this module without the built extension, or
creating a plan without a CUDA device, raises.  Build the library with
``python -c "import __graft_entry__ as g; g.build()"`` (or ``make -C
track_analyser_b200/csrc``).
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# TA_B200_LIB: another build of the same library (A/B timing of kernel variants); the default is the in-tree build
LIB_PATH = os.environ.get("TA_B200_LIB") or os.path.join(_HERE, "libta_b200.so")

TA_ABI_VERSION = 4
TA_OK = 0
TA_ERR_INVALID = -1
TA_ERR_CUDA = -2
TA_ERR_UNSUPPORTED = -3
TA_ERR_WORKSPACE = -4
TA_PCM_S16, TA_PCM_S24, TA_PCM_S32, TA_PCM_F32 = 1, 2, 3, 4


class NativeError(RuntimeError):
    """Raised when a C-ABI call returns a negative TA_ERR_* code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libta_b200 error {code}: {message}")
        self.code = code


class PlanDesc(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("sample_rate", C.c_int32),
        ("n_fft", C.c_int32),
        ("hop", C.c_int32),
        ("n_mels", C.c_int32),
        ("n_chroma", C.c_int32),
        ("tempogram_win", C.c_int32),
        ("reserved", C.c_int32),
        ("fmin", C.c_double),
        ("fmax", C.c_double),
        ("roll_percent", C.c_double),
        ("meter_block", C.c_double),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_tracks", C.c_int32),
        ("channels", C.c_int32),
        ("pcm", C.c_void_p),
        ("pcm_offset", C.POINTER(C.c_int64)),
        ("n_samples", C.POINTER(C.c_int64)),
    ]


class FrontendOut(C.Structure):
    _fields_ = [
        ("magnitude", C.c_void_p),
        ("mel", C.c_void_p),
        ("onset_env", C.c_void_p),
        ("autocorr", C.c_void_p),
        ("flux_linear", C.c_void_p),
        ("ltas", C.c_void_p),
        ("centroid", C.c_void_p),
        ("rolloff_bin", C.c_void_p),
        ("band_energy", C.c_void_p),
        ("moments", C.c_void_p),
        ("kw_blocks", C.c_void_p),
        ("lufs", C.c_void_p),
        ("rms_momentary", C.c_void_p),
        ("rms_short", C.c_void_p),
        ("frame_max", C.c_void_p),
        ("chroma", C.c_void_p),
        ("tuning", C.c_void_p),
        ("tempogram", C.c_void_p),
        ("true_peak", C.c_void_p),
        ("hpss_harmonic", C.c_void_p),
        ("hpss_percussive", C.c_void_p),
        ("hpss_scratch", C.c_void_p),
        ("mfcc", C.c_void_p),
        ("self_similarity", C.c_void_p),
        ("chroma_cqt", C.c_void_p),
        ("cqt_tuning", C.c_void_p),
        ("cqt_mag", C.c_void_p),
        ("cqt_scratch", C.c_void_p),
        ("kw_pitch", C.c_int32),
        ("rms_pitch", C.c_int32),
        ("cqt_scratch_bytes", C.c_uint64),
        ("true_peak_oversample", C.c_int32),
        ("reserved", C.c_int32),
    ]


# every symbol include/ta_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "ta_abi_version": (C.c_int, []),
    "ta_last_error": (C.c_char_p, []),
    "ta_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    "ta_plan_create_window": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_double), C.POINTER(C.c_void_p)]),
    "ta_plan_destroy": (None, [C.c_void_p]),
    "ta_plan_n_bins": (C.c_int, [C.c_void_p]),
    "ta_plan_table": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]),
    "ta_workspace_bytes": (C.c_size_t, [C.c_void_p, C.POINTER(Batch)]),
    "ta_frontend_run": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(FrontendOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_frontend_run_host": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(FrontendOut), C.c_void_p]),
    "ta_frontend_run_profiled": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(FrontendOut), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_float)]),
    "ta_launch_count": (C.c_uint64, []),
    "ta_stft_features": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(FrontendOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_onset_flux": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ta_autocorrelate": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_chroma_stft": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_cqt_frame_count": (C.c_int64, [C.c_void_p, C.c_int64]),
    "ta_cqt_scratch_bytes": (C.c_size_t, [C.c_void_p, C.POINTER(Batch)]),
    "ta_chroma_cqt": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_mono_mix_fingerprint": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "ta_decode_pcm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    "ta_hpss_curves": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_hpss_components": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_tempogram": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_time_domain": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.POINTER(FrontendOut), C.c_void_p, C.c_size_t, C.c_void_p]),
    "ta_resampler_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ta_resampler_destroy": (None, [C.c_void_p]),
    "ta_resampler_out_len": (C.c_int64, [C.c_void_p, C.c_int64]),
    "ta_resample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the extension once; fail loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension is not built and there is no CPU fallback. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` from the repository root."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    got = lib.ta_abi_version()
    if got != TA_ABI_VERSION:
        raise RuntimeError(f"libta_b200 ABI {got} does not match the binding ({TA_ABI_VERSION})")
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != TA_OK:
        msg = load().ta_last_error()
        raise NativeError(code, msg.decode("utf-8", "replace") if msg else "")
