"""mini_parallel_b200 -- B200-native Smith-Waterman scoring engine.

Drop-in for the alignment path of bmwoolf/mini_parallel (``gpu_align`` and its callers,
smith_waterman/src/aligner.rs:365-544).  The compute path is hand-written CUDA for sm_100a
behind the C ABI of ``include/swb200.h``; this package is the thin ctypes host mirror used by
the tests, ``bench.py`` and ``__graft_entry__``.  There is no CPU fallback: loading fails
loudly when ``libswb200.so`` has not been built (``make`` / ``__graft_entry__.build()``).
"""
from ._lib import load_library, LIB_PATH, SwbResult, RESULT_DTYPE  # noqa: F401
from .engine import Engine, MultiEngine, SwbError, device_count  # noqa: F401
from . import aligner  # noqa: F401

__all__ = ["Engine", "MultiEngine", "SwbError", "device_count", "load_library", "LIB_PATH", "SwbResult", "RESULT_DTYPE", "aligner"]
