#!/usr/bin/env python
"""Side measurement: fp32 L2 batch (256 queries, k=10) over 10M x 768 on the coarse route vs the exact CUDA-core kernel."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from redisearch_b200 import vecsim as vs
    from redisearch_b200._lib import load_library

    rows, dim, k, nq = int(os.environ.get("ROWS", 10_000_000)), 768, 10, 256
    dev = torch.device("cuda", 0)
    L = vs.lib()
    S = load_library("libsynth_b200.so")
    S.Synth_FillRows.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    index = vs.VecSimIndex(vs.VecSimType_FLOAT32, dim, vs.VecSimMetric_L2)
    assert L.VecSimB200_Reserve(index.h, rows) == 0
    chunk = 1_000_000
    buf = torch.empty((chunk, dim), dtype=torch.float32, device=dev)
    done = 0
    while done < rows:
        n = min(chunk, rows - done)
        assert S.Synth_FillRows(buf.data_ptr(), dim * 4, 0, 42, done, n, dim, sp) == 0
        torch.cuda.synchronize()
        assert L.VecSimB200_AddVectorsDevice(index.h, buf.data_ptr(), n, done + 1) == n
        done += n
    q = torch.empty((nq, dim), dtype=torch.float32, device=dev)
    assert S.Synth_FillRows(q.data_ptr(), dim * 4, 0, 43, 0, nq, dim, sp) == 0
    ol = torch.empty((nq, k), dtype=torch.int64, device=dev)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=dev)
    res, keep = {}, None
    for mode in (1, 0):
        L.VecSimB200_SetCoarseMode(mode)
        steps = 10 if mode else 1
        for _ in range(2 if mode else 1):
            assert L.VecSimB200_TopKQueryBatchDevice(index.h, q.data_ptr(), nq, k, ol.data_ptr(), os_.data_ptr(), sp) == 0
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            assert L.VecSimB200_TopKQueryBatchDevice(index.h, q.data_ptr(), nq, k, ol.data_ptr(), os_.data_ptr(), sp) == 0
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        flags = None
        if mode:
            import numpy as np
            f = np.zeros(nq, dtype=np.uint32)
            if L.VecSimB200_LastCoarseFlags(index.h, f.ctypes.data, nq) == 0:
                flags = int(f.sum())
            keep = (ol.clone(), os_.clone())
        res["coarse" if mode else "exact_cuda_core"] = {"ms_per_batch": ms, "qps": nq / ms * 1000.0, "route": L.VecSimB200_LastBatchPath(index.h),
                                                        "queries_proven": flags}
    identical = bool((keep[0] == ol).all().item()) and bool((keep[1] == os_).all().item())
    print(json.dumps({"workload": f"FLAT {rows} x {dim} fp32 L2 k={k} batch={nq}", **res, "identical_ids_and_score_bits": identical}))


if __name__ == "__main__":
    main()
