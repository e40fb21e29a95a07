"""islamcts_b200 -- B200-native engine for the simulate loop of LorenzoBonanni/IslaMcts.

Only the hot path lives here: ``csrc/`` (hand-written sm_100a CUDA kernels + the C ABI of
``include/isla_b200.h``), ``_lib``/``engine`` (ctypes binding; torch owns memory and streams) and the
host-side mirror of the reference's agent API (``agents``, ``utils``; also importable under the
reference's own package name ``islaMcts``).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from ._lib import IslaError  # noqa: F401

__all__ = ["IslaError", "_lib"]
__version__ = "0.1.0"
