"""B200-native spectral feature frontend behind track-analyser's Python API.

Module names, function signatures and dataclasses mirror the reference package
(``track_analyser``); the arithmetic runs in hand-written sm_100a kernels reached
through the C ABI in ``include/ta_b200.h``.  There is no CPU fallback.
"""

__version__ = "0.1.0"


def get_version() -> str:
    """Mirror of track_analyser.get_version (reference __init__.py:12-23)."""
    return __version__


def install(reference_package: str = "track_analyser", *, force: bool = False):
    """Run the unmodified reference package on the B200 frontend: see ``compat.install``."""
    from .compat import install as _install

    return _install(reference_package, force=force)
