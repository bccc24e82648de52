#!/usr/bin/env python
"""Side measurement (not the bench contract): one GPU's share of BASELINE.json configs[2] —
FLAT 6.25M x 768 fp16, inner product, k=100, batch=1024 — through VecSimB200_TopKQueryBatchDevice.
Prints one JSON line; used for profiles/, not by the driver."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch

    from redisearch_b200 import vecsim as vs
    from redisearch_b200._lib import load_library

    rows, dim, k, nq = int(os.environ.get("ROWS", 6_250_000)), 768, 100, 1024
    dev = torch.device("cuda", 0)
    L = vs.lib()
    S = load_library("libsynth_b200.so")
    S.Synth_FillRows.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    index = vs.VecSimIndex(vs.VecSimType_FLOAT16, dim, vs.VecSimMetric_IP)
    assert L.VecSimB200_Reserve(index.h, rows) == 0
    chunk = 1_000_000
    buf = torch.empty((chunk, dim), dtype=torch.float16, device=dev)
    done = 0
    while done < rows:
        n = min(chunk, rows - done)
        assert S.Synth_FillRows(buf.data_ptr(), dim * 2, vs.VecSimType_FLOAT16, 42, done, n, dim, sp) == 0
        torch.cuda.synchronize()
        assert L.VecSimB200_AddVectorsDevice(index.h, buf.data_ptr(), n, done + 1) == n
        done += n
    q = torch.empty((nq, dim), dtype=torch.float16, device=dev)
    assert S.Synth_FillRows(q.data_ptr(), dim * 2, vs.VecSimType_FLOAT16, 43, 0, nq, dim, sp) == 0
    ol = torch.empty((nq, k), dtype=torch.int64, device=dev)
    os_ = torch.empty((nq, k), dtype=torch.float32, device=dev)
    res = {}
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
        res["tensor" if mode else "cuda_core"] = {"ms_per_batch": ms, "qps": nq / ms * 1000.0, "path": L.VecSimB200_LastBatchPath(index.h),
                                                   "tflops": 2.0 * nq * rows * dim / ms / 1e9}
        if mode:
            keep = (ol.clone(), os_.clone())
    same = bool((keep[0] == ol).float().mean().item() > 0.999) and bool((keep[1] - os_).abs().max().item() < 1e-3)
    print(json.dumps({"workload": f"FLAT {rows} x {dim} fp16 IP k={k} batch={nq} (one GPU's shard of configs[2])", **res,
                      "routes_agree": same}))


if __name__ == "__main__":
    main()
