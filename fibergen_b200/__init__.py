"""fibergen_b200 -- B200-native (sm_100a) Lippmann-Schwinger solve loop of fospald/fibergen.

Only what the hot path needs lives here: ``csrc/`` (hand-written CUDA kernels, the C ABI of
``include/fgb200.h`` and the host-side C++ mirror of the reference's LSSolver) and thin ctypes views
of both (``Context``, ``LSSolver``).  Importing the package does not load the CUDA library; the first
object you create does, and it raises if ``libfgb200.so`` was not built -- there is no CPU fallback.
"""
from .lib import FgbError, LIB_PATH, load          # noqa: F401
from .solver import Context, LSSolver, lame, pad, unpad, nzp_of   # noqa: F401

__all__ = ["Context", "LSSolver", "FgbError", "load", "lame", "pad", "unpad", "nzp_of", "LIB_PATH"]
