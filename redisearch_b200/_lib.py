"""Locate and load the in-tree shared libraries.  No fallback: a missing library is an error."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBDIR = os.path.join(_HERE, "lib")
_cache = {}


def lib_path(name: str) -> str:
    return os.path.join(_LIBDIR, name)


def load_library(name: str) -> ctypes.CDLL:
    """dlopen ``redisearch_b200/lib/<name>``; build it first with ``__graft_entry__.build()``."""
    if name in _cache:
        return _cache[name]
    path = lib_path(name)
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the CUDA extension has not been built "
            f"(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
        )
    lib = ctypes.CDLL(path)  # RTLD_LOCAL: the checkers under oracle/ define some of the same C symbols
    _cache[name] = lib
    return lib
