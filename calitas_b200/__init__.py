"""calitas_b200 — B200-native engine for the CALITAS SearchReference / AlignToReference hot path.

The package is a thin ctypes layer over libcalitas_b200.so (CUDA kernels + C ABI, include/calitas_b200.h and
include/calitas_b200_tools.h).  There is no CPU fallback: importing works anywhere, but every call that computes needs the
built library and a CUDA device and fails loudly otherwise.
"""
from ._capi import Library, CalitasError, default_library, Engine, DEFAULT_COSTS, Limits  # noqa: F401

__all__ = ["Library", "CalitasError", "default_library", "Engine", "DEFAULT_COSTS", "Limits"]
