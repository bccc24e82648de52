"""ctypes loaders for the CHECKERS under oracle/ (test infrastructure; never used by the product).

  port()        oracle/liboracle.so          our CPU restatement (plain C)
  ref_vecsim()  oracle/_ref/libvecsim_ref.so the reference's own VecSim sources, compiled in place
  ref_scorers() oracle/_ref/libscorers_ref.so the reference's src/ext/default.c
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

F32, BF16, F16, I8, U8 = 0, 2, 3, 4, 5
L2, IP, COS = 0, 1, 2
TIER_SCALAR, TIER_AVX512 = 0, 1
NP_DTYPE = {F32: np.float32, BF16: np.uint16, F16: np.uint16, I8: np.int8, U8: np.uint8}

_P, _SZ = C.c_void_p, C.c_size_t
_port = None
_ref = None


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def build_port():
    subprocess.run(["make", "-C", ODIR, "liboracle.so"], check=True, capture_output=True)


def port():
    global _port
    if _port is None:
        path = os.path.join(ODIR, "liboracle.so")
        if not os.path.exists(path):
            build_port()
        L = C.CDLL(path)
        L.orc_distance.restype = C.c_float
        L.orc_distance.argtypes = [C.c_int, C.c_int, _SZ, _P, _P, C.c_int]
        L.orc_normalize.argtypes = [_P, _SZ, C.c_int]
        L.orc_stored_size.restype = _SZ
        L.orc_stored_size.argtypes = [C.c_int, _SZ, C.c_int]
        L.orc_half_to_float.restype = C.c_float
        L.orc_half_to_float.argtypes = [C.c_uint16]
        L.orc_float_to_half.restype = C.c_uint16
        L.orc_float_to_half.argtypes = [C.c_float]
        L.orc_float_to_bf16.restype = C.c_uint16
        L.orc_float_to_bf16.argtypes = [C.c_float]
        L.orc_index_new.restype = _P
        L.orc_index_new.argtypes = [C.c_int, _SZ, C.c_int, C.c_int, C.c_int]
        L.orc_index_free.argtypes = [_P]
        L.orc_index_add.restype = C.c_int
        L.orc_index_add.argtypes = [_P, _P, _SZ]
        L.orc_index_add_bulk.argtypes = [_P, _P, _SZ, _SZ, _SZ]
        L.orc_index_delete.restype = C.c_int
        L.orc_index_delete.argtypes = [_P, _SZ]
        L.orc_index_size.restype = _SZ
        L.orc_index_size.argtypes = [_P]
        L.orc_index_topk.restype = _SZ
        L.orc_index_topk.argtypes = [_P, _P, _SZ, C.c_int, _P, _P]
        L.orc_index_range.restype = _SZ
        L.orc_index_range.argtypes = [_P, _P, C.c_double, C.c_int, _SZ, _P, _P]
        L.orc_index_distance_from.restype = C.c_double
        L.orc_index_distance_from.argtypes = [_P, _SZ, _P]
        L.orc_index_prefer_adhoc.restype = C.c_int
        L.orc_index_prefer_adhoc.argtypes = [_P, _SZ, _SZ, C.c_int]
        L.orc_index_all_sorted.restype = _SZ
        L.orc_index_all_sorted.argtypes = [_P, _P, _P, _P]
        L.orc_index_time_topk.restype = C.c_double
        L.orc_index_time_topk.argtypes = [_P, _P, _SZ, _SZ, _SZ, C.c_int, _P, _P]
        L.orc_mix64.restype = C.c_uint64
        L.orc_mix64.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_synth_f32.restype = C.c_float
        L.orc_synth_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_synth_rows.argtypes = [C.c_int, C.c_uint64, C.c_uint64, _SZ, _SZ, _P]
        _port = L
    return _port


def ref_vecsim():
    """None when oracle/_ref was not built (no /root/reference on this machine and no prebuilt .so)."""
    global _ref
    if _ref is None:
        path = os.path.join(ODIR, "_ref", "libvecsim_ref.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.Ref_IndexNew.restype = _P
        L.Ref_IndexNew.argtypes = [C.c_int, _SZ, C.c_int, C.c_int, _SZ]
        L.Ref_IndexFree.argtypes = [_P]
        L.Ref_AddVector.restype = C.c_int
        L.Ref_AddVector.argtypes = [_P, _P, _SZ]
        L.Ref_AddVectors.argtypes = [_P, _P, _SZ, _SZ, _P, _SZ]
        L.Ref_DeleteVector.restype = C.c_int
        L.Ref_DeleteVector.argtypes = [_P, _SZ]
        L.Ref_IndexSize.restype = _SZ
        L.Ref_IndexSize.argtypes = [_P]
        L.Ref_TopK.restype = _SZ
        L.Ref_TopK.argtypes = [_P, _P, _SZ, C.c_int, _SZ, _P, _P, _P]
        L.Ref_Range.restype = _SZ
        L.Ref_Range.argtypes = [_P, _P, C.c_double, C.c_int, _SZ, _P, _P, _P]
        L.Ref_GetDistanceFrom.restype = C.c_double
        L.Ref_GetDistanceFrom.argtypes = [_P, _SZ, _P]
        L.Ref_PreferAdHoc.restype = C.c_int
        L.Ref_PreferAdHoc.argtypes = [_P, _SZ, _SZ, C.c_int]
        L.Ref_BatchNew.restype = _P
        L.Ref_BatchNew.argtypes = [_P, _P]
        L.Ref_BatchNext.restype = _SZ
        L.Ref_BatchNext.argtypes = [_P, _SZ, C.c_int, _SZ, _P, _P]
        L.Ref_BatchHasNext.restype = C.c_int
        L.Ref_BatchHasNext.argtypes = [_P]
        L.Ref_BatchReset.argtypes = [_P]
        L.Ref_BatchFree.argtypes = [_P]
        L.Ref_Distance.restype = C.c_float
        L.Ref_Distance.argtypes = [C.c_int, C.c_int, _SZ, _P, _P]
        L.Ref_Distances.argtypes = [C.c_int, C.c_int, _SZ, _P, _SZ, _SZ, _P, _P]
        L.Ref_Normalize.argtypes = [_P, _SZ, C.c_int]
        L.Ref_TimeTopK.restype = C.c_double
        L.Ref_TimeTopK.argtypes = [_P, _P, _SZ, _SZ, _SZ, C.c_int, _P, _P]
        _ref = L
    return _ref


def host_has_avx512f() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return " avx512f" in f.read()
    except OSError:
        return False


# ---------------------------------------------------------------------------------------------
# convenience wrappers
# ---------------------------------------------------------------------------------------------
class PortIndex:
    def __init__(self, vtype, dim, metric, multi=False, tier=TIER_AVX512):
        self.L = port()
        self.h = self.L.orc_index_new(vtype, dim, metric, int(multi), tier)
        self.vtype, self.dim, self.metric = vtype, dim, metric

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_index_free(self.h)
            self.h = None

    def add(self, blob, label):
        blob = np.ascontiguousarray(blob)
        return self.L.orc_index_add(self.h, _p(blob), label)

    def add_many(self, blobs, label0=0):
        blobs = np.ascontiguousarray(blobs)
        self.L.orc_index_add_bulk(self.h, _p(blobs), blobs.strides[0], blobs.shape[0], label0)

    def delete(self, label):
        return self.L.orc_index_delete(self.h, label)

    def size(self):
        return self.L.orc_index_size(self.h)

    def topk(self, q, k, order=0):
        q = np.ascontiguousarray(q)
        cap = max(1, min(k, self.size()))
        labels = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        n = self.L.orc_index_topk(self.h, _p(q), k, order, _p(labels), _p(scores))
        return labels[:n].astype(np.int64), scores[:n]

    def range(self, q, radius, order=0):
        q = np.ascontiguousarray(q)
        cap = max(1, self.size())
        labels = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        n = self.L.orc_index_range(self.h, _p(q), radius, order, cap, _p(labels), _p(scores))
        return labels[:n].astype(np.int64), scores[:n]

    def distance_from(self, label, q):
        q = np.ascontiguousarray(q)
        return self.L.orc_index_distance_from(self.h, label, _p(q))

    def prefer_adhoc(self, subset, k, initial):
        return bool(self.L.orc_index_prefer_adhoc(self.h, subset, k, int(initial)))

    def all_sorted(self, q):
        q = np.ascontiguousarray(q)
        cap = max(1, self.size())
        labels = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        n = self.L.orc_index_all_sorted(self.h, _p(q), _p(labels), _p(scores))
        return labels[:n].astype(np.int64), scores[:n]


class RefIndex:
    """The reference's BruteForceIndex through oracle/ref_shim/ref_capi.cpp."""

    def __init__(self, vtype, dim, metric, multi=False, block_size=1024):
        self.L = ref_vecsim()
        self.h = self.L.Ref_IndexNew(vtype, dim, metric, int(multi), block_size)
        self.vtype, self.dim, self.metric = vtype, dim, metric

    def __del__(self):
        if getattr(self, "h", None):
            self.L.Ref_IndexFree(self.h)
            self.h = None

    def add(self, blob, label):
        blob = np.ascontiguousarray(blob)
        return self.L.Ref_AddVector(self.h, _p(blob), label)

    def add_many(self, blobs, label0=0):
        blobs = np.ascontiguousarray(blobs)
        self.L.Ref_AddVectors(self.h, _p(blobs), blobs.shape[0], blobs.strides[0], None, label0)

    def delete(self, label):
        return self.L.Ref_DeleteVector(self.h, label)

    def size(self):
        return self.L.Ref_IndexSize(self.h)

    def topk(self, q, k, order=0):
        q = np.ascontiguousarray(q)
        cap = max(1, k)
        ids = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        code = C.c_int(0)
        n = self.L.Ref_TopK(self.h, _p(q), k, order, cap, _p(ids), _p(scores), C.byref(code))
        return ids[:n].astype(np.int64), scores[:n]

    def range(self, q, radius, order=0):
        q = np.ascontiguousarray(q)
        cap = max(1, self.size())
        ids = np.empty(cap, dtype=np.uint64)
        scores = np.empty(cap, dtype=np.float64)
        code = C.c_int(0)
        n = self.L.Ref_Range(self.h, _p(q), radius, order, cap, _p(ids), _p(scores), C.byref(code))
        return ids[:n].astype(np.int64), scores[:n]

    def distance_from(self, label, q):
        q = np.ascontiguousarray(q)
        return self.L.Ref_GetDistanceFrom(self.h, label, _p(q))

    def prefer_adhoc(self, subset, k, initial):
        return bool(self.L.Ref_PreferAdHoc(self.h, subset, k, int(initial)))

    def batches(self, q, n, order=0):
        """Drain a batch iterator n results at a time; returns a list of (ids, scores)."""
        q = np.ascontiguousarray(q)
        it = self.L.Ref_BatchNew(self.h, _p(q))
        out = []
        while self.L.Ref_BatchHasNext(it):
            ids = np.empty(n, dtype=np.uint64)
            scores = np.empty(n, dtype=np.float64)
            m = self.L.Ref_BatchNext(it, n, order, n, _p(ids), _p(scores))
            out.append((ids[:m].astype(np.int64), scores[:m]))
            if m == 0:
                break
        self.L.Ref_BatchFree(it)
        return out


def synth_rows(vtype, seed, row0, nrows, dim):
    """Counter-based synthetic rows (oracle/vecsim_oracle.c orc_synth_rows)."""
    out = np.empty((nrows, dim), dtype=NP_DTYPE[vtype])
    port().orc_synth_rows(vtype, seed, row0, nrows, dim, _p(out))
    return out


def to_type(x32: np.ndarray, vtype: int) -> np.ndarray:
    """fp32 values -> blobs of `vtype` with the reference's conversions."""
    L = port()
    if vtype == F32:
        return np.ascontiguousarray(x32, dtype=np.float32)
    flat = np.ascontiguousarray(x32, dtype=np.float32).ravel()
    if vtype == F16:
        return np.array([L.orc_float_to_half(float(v)) for v in flat], dtype=np.uint16).reshape(x32.shape)
    if vtype == BF16:
        return np.array([L.orc_float_to_bf16(float(v)) for v in flat], dtype=np.uint16).reshape(x32.shape)
    if vtype == I8:
        return np.rint(127.0 * x32).astype(np.int8)
    return np.rint(127.5 * x32 + 127.5).astype(np.uint8)
