"""redisearch_b200 — B200-native drop-in for RediSearch's query-time scoring hot path.

The product is two C-ABI shared libraries built from ``redisearch_b200/csrc`` (hand-written sm_100a
CUDA + C++ host code): ``lib/libvecsim_b200.so`` (FLAT KNN behind the VecSim C API,
``include/vecsim_b200.h``) and ``lib/libii_b200.so`` (posting-list intersection/union + BM25 behind
the QueryIterator / scorer surface, ``include/ii_b200.h``).  This Python package is only the thin
ctypes face used by tests, bench.py and __graft_entry__; it contains no compute and no fallback:
importing a binding whose library is missing raises.
"""
from ._lib import lib_path, load_library  # noqa: F401

__all__ = ["lib_path", "load_library"]
