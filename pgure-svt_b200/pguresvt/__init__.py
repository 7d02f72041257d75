"""B200-native PGURE-SVT: drop-in for the `pguresvt` Python package of tjof2/pgure-svt.

Exports the same two names as the reference (pguresvt/__init__.py:4-6).  All numerical work is done by
the CUDA library behind include/pguresvt_b200.h; importing this package on a machine without that
library raises — there is no CPU fallback.
"""
from .svt import SVT, mixed_noise_model


def release_cached():
    """Free the per-device handles (device buffers, page-locked staging) the one-shot entry points keep between calls."""
    from ._pguresvt import release_cached as _rc

    _rc()


__all__ = ["mixed_noise_model", "SVT", "release_cached"]
__version__ = "0.6.4+b200.2"
