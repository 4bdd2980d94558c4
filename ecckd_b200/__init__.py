"""ecckd_b200 - B200-native (sm_100a) implementation of ecCKD's data-parallel hot path.

Only what the path needs lives here: ``csrc/`` (CUDA kernels + the C ABI declared in
``include/ecckd_b200.h``), the ctypes loader and a thin Python mirror of the
reference's operator interface (``api``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
