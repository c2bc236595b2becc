"""minimaloptix_b200 — B200-native render path behind MinimalOptiX's surface.

Python here is plumbing only: ctypes bindings of the C ABI (include/mox.h, include/mox_host.h)
used by the tests and bench.py.  The product is libmox.so (hand-written sm_100a CUDA) and
libmox_host.so / mox_cli (C++ host side).  There is no CPU fallback: `gpu()` raises if the
CUDA library is missing.
"""
import os

from . import structs
from ._binding import Backend, Context, MoxError, GPU_ONLY

_HERE = os.path.dirname(os.path.abspath(__file__))
GPU_LIB = os.environ.get("MOX_GPU_LIB") or os.path.join(_HERE, "libmox.so")  # override: A/B testing of builds

_gpu = None


def gpu():
    """The CUDA backend (libmox.so, prefix mox_).  Fails loudly when it is not built."""
    global _gpu
    if _gpu is None:
        _gpu = Backend(GPU_LIB, "mox_", extra=GPU_ONLY)
    return _gpu


__all__ = ["structs", "Backend", "Context", "MoxError", "gpu", "GPU_LIB"]
