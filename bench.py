#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (contract in the task statement).

Workload (BASELINE.json configs[1]): FLAT 10M x 768 fp32, cosine, k=10, batch=256 queries, 1 x B200.
A "step" is one pass of the KNN hot path over one batch of 256 synthetic queries.

  value   queries/s with queries and corpus already resident in HBM (VecSimB200_TopKQueryBatchDevice),
          timed with CUDA events on the launching stream, max over ranks.
  e2e     the same through the host-facing C-ABI (VecSimB200_TopKQueryBatch): host query blobs in,
          host labels/scores out, H2D + D2H inside the timed region.
  roofline  the dominant kernel, timed with CUDA events inside the library.  The batched pass is tensor-bound
            (2*B*N*D flop against the measured dense bf16 peak in MEASURED_PEAKS.json); the HBM view
            (algorithmic bytes = N*D*4 per launch, and the bytes the fp16 shadow actually costs) sits beside it
            under roofline.hbm.  The single-query leg is HBM-bound.
  cpu_baseline  the reference's own brute-force code (oracle/_ref, built from /root/reference) or,
            if that library is absent, our C restatement, on a bounded sample.

N > 1 (torchrun): every rank owns its own 10M-row shard of an (N x 10M)-row corpus (weak scaling);
a step scans the local shard for the same 256 queries, all-gathers the per-shard top-k over NCCL and
merges on device (VecSimB200_MergeShardTopK).  value = N*256 / step time = throughput in units of
"query x 10M-row shard".

--impl reference: times the reference CPU implementation on the host cores (rank 0 only).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "KNN QPS @k=10 on 10M x 768 fp32 (cosine, batch=256)"
N_ROWS, DIM, K, BATCH = 10_000_000, 768, 10, 256
SEED_ROWS, SEED_QUERIES = 42, 43


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def load_tensor_peak():
    """Dense bf16/fp16 tensor throughput (TFLOP/s): the burst figure, for a kernel timed alone."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 2250.0, 2250.0, "fallback (nominal dense bf16, B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []
        self.stop = threading.Event()
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=3)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no nvidia-smi samples"]}
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU code on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_qps(sample_rows, queries_per_thread, threads):
    """Times BruteForceIndex::topKQuery (oracle/_ref = the reference's sources) — or our C restatement
    if that library was not built — on `sample_rows` x 768 cosine rows; scales QPS linearly to 10M rows
    (the scan is a streaming pass).  Returns (qps_at_10M, info)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np

    import oracle_lib as ol

    rows = ol.synth_rows(ol.F32, SEED_ROWS, 0, sample_rows, DIM)
    nq = queries_per_thread * threads
    qs = ol.synth_rows(ol.F32, SEED_QUERIES, 0, nq, DIM)
    ref = ol.ref_vecsim()
    if ref is not None:
        kind = "reference"
        ix = ol.RefIndex(ol.F32, DIM, ol.COS)
        ix.add_many(rows, 1)
        secs = ref.Ref_TimeTopK(ix.h, ol._p(qs), qs.strides[0], nq, K, threads, None, None)
    else:
        kind = "port"
        ix = ol.PortIndex(ol.F32, DIM, ol.COS, tier=ol.TIER_AVX512)
        ix.add_many(rows, 1)
        secs = ol.port().orc_index_time_topk(ix.h, ol._p(qs), qs.strides[0], nq, K, threads, None, None)
    qps_sample = nq / secs
    qps_full = qps_sample * sample_rows / N_ROWS
    info = {"value": qps_full, "unit": "queries/s", "cores": threads, "kind": kind,
            "sample": f"{nq} queries x {sample_rows} rows x {DIM} fp32 cosine, {threads} threads, one query per "
                      f"thread (brute_force.h:243-291); QPS scaled x{sample_rows}/{N_ROWS} to 10M rows",
            "sample_seconds": secs}
    del ix
    return qps_full, info


def run_reference_arm(args, rank):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = []
    info = None
    total = args.warmup + args.steps
    for i in range(total):
        t0 = time.perf_counter()
        qps, info = cpu_reference_qps(args.ref_sample_rows, 1, threads)
        if i >= args.warmup:
            steps.append((qps, time.perf_counter() - t0))
        if time.perf_counter() - t0 > 120:
            break
    qps = statistics.mean(s[0] for s in steps)
    line = {"impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": len(steps), "warmup": args.warmup, "ms_per_step": 1000.0 * BATCH / qps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "FLAT 10M x 768 fp32 cosine k=10 batch=256 (reference CPU path, bounded sample)"},
            "cpu_baseline": info,
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# second half of the metric: BM25 intersect docs/sec (BASELINE.json configs[3])
# ------------------------------------------------------------------------------------------------
POSTING_QUERIES = [(1, 2, 3), (1, 10, 100), (2, 5, 9), (3, 30, 300), (1, 100, 10000), (10, 20, 30), (4, 8, 16), (50, 60, 70)]


def cpu_postings_baseline(n_docs, threads):
    """3-term AND + BM25STD + top-10 on the host with our C restatement of the reference's Rust iterators
    (kind "port": the reference's posting path cannot be built here — no Rust toolchain)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np

    import oracle_lib as ol

    L = ol.postings()
    doc_len = np.zeros(n_docs + 1, dtype=np.uint32)
    # doc lengths via the same hash (vectorised replica of orc_synth_doclen is not needed for timing: constant cost)
    doc_len[1:] = 50 + (np.arange(1, n_docs + 1, dtype=np.uint64) * np.uint64(2654435761) % np.uint64(451)).astype(np.uint32)
    reps = max(1, threads // len(POSTING_QUERIES))
    terms, keep, postings = [], [], 0
    for q in POSTING_QUERIES:
        trio = []
        for r in q:
            ix = ol.InvIndex(ol.CODEC_FREQS_ONLY)
            postings += L.orc_ii_fill_synth(ix.h, n_docs, r)
            trio.append(ix)
        keep.append(trio)
    for _ in range(reps):
        for trio in keep:
            terms += [ix.h for ix in trio]
    nq = len(terms) // 3
    arr = (C.c_void_p * len(terms))(*terms)
    ids = np.zeros(nq * 10, dtype=np.uint64)
    sc = np.zeros(nq * 10, dtype=np.float64)
    hits = np.zeros(nq, dtype=np.uint64)
    secs = L.orc_time_search3(arr, nq, ol._p(doc_len), n_docs, float(doc_len[1:].mean()), 10, min(threads, nq), ol._p(ids), ol._p(sc), ol._p(hits))
    total = postings * reps
    return {"value": total / secs, "unit": "postings/s", "cores": min(threads, nq), "kind": "port",
            "sample": f"{nq} queries (the {len(POSTING_QUERIES)} rank triples x {reps}) over a {n_docs}-doc synthetic Zipf index, FreqsOnly blocks, "
                      f"reader+Intersection::read+BM25STD+top-10 per query, one query per thread; {total} input postings",
            "sample_seconds": secs}


def bench_postings(torch, dev, stream_ptr, n_docs, steps):
    import numpy as np

    from redisearch_b200 import postings as ps
    from redisearch_b200._lib import load_library

    S = load_library("libsynth_b200.so")
    S.Synth_DocFreq.restype = C.c_uint64
    S.Synth_DocFreq.argtypes = [C.c_uint64, C.c_uint64]
    S.Synth_Postings.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    S.Synth_DocLens.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p]
    S.Synth_EncodeFreqsOnlyBlocks.restype = C.c_size_t
    S.Synth_EncodeFreqsOnlyBlocks.argtypes = [C.c_void_p] * 8 + [C.POINTER(C.c_size_t)]
    L = ps.lib()
    chunks = (n_docs + 1023) // 1024
    scratch = torch.empty(2 * chunks + 16, dtype=torch.int32, device=dev)
    d_total = torch.zeros(4, dtype=torch.int32, device=dev)
    h_count = np.zeros(4, dtype=np.uint32)
    d_len = torch.empty(n_docs + 1, dtype=torch.int32, device=dev)
    assert S.Synth_DocLens(n_docs, d_len.data_ptr(), stream_ptr) == 0
    torch.cuda.synchronize()
    avg_len = float(d_len[1:].double().mean().item())
    dt = L.II_DocTable_FromDevice(n_docs, d_len.data_ptr(), None, None)
    assert dt
    ranks = sorted({r for q in POSTING_QUERIES for r in q})
    lists, host_lists = {}, {}
    for r in ranks:
        cap = int(S.Synth_DocFreq(n_docs, r) * 1.2) + 4096
        ids = torch.empty(cap, dtype=torch.int32, device=dev)
        fr = torch.empty(cap, dtype=torch.int32, device=dev)
        assert S.Synth_Postings(n_docs, r, ids.data_ptr(), fr.data_ptr(), scratch.data_ptr(), d_total.data_ptr(), h_count.ctypes.data, stream_ptr) == 0
        n = int(h_count[0])
        lists[r] = L.II_PostingList_FromDevice(ids.data_ptr(), fr.data_ptr(), n)
        host_lists[r] = (ids[:n].cpu().numpy().view(np.uint32).copy(), fr[:n].cpu().numpy().view(np.uint32).copy())
        assert lists[r]
    st = ps.II_IndexStats(n_docs, 0, avg_len)

    def run_query(q, handles):
        arr = (C.c_void_p * 3)(*handles)
        terms = (ps.II_TermParams * 3)(*[ps.II_TermParams(1.0, L.II_CalculateIDF(n_docs, len(host_lists[r][0])),
                                                          L.II_CalculateIDF_BM25(n_docs, len(host_lists[r][0]))) for r in q])
        ids = np.zeros(10, dtype=np.uint64)
        sc = np.zeros(10, dtype=np.float64)
        tot = C.c_size_t(0)
        got = L.II_SearchTopN(arr, 3, 0, ps.SCORER_BM25STD, terms, 1.0, C.byref(st), dt, 10, ids.ctypes.data, sc.ctypes.data, C.byref(tot))
        return ids[:got].copy(), sc[:got].copy(), tot.value

    in_postings = sum(len(host_lists[r][0]) for q in POSTING_QUERIES for r in q)
    for q in POSTING_QUERIES:  # warm-up
        run_query(q, [lists[r] for r in q])
    ps.stats(reset=True)
    dev_us, hits, results = 0.0, 0, {}
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for q in POSTING_QUERIES:
            results[q] = run_query(q, [lists[r] for r in q])
            s_ = ps.stats(reset=False)
            dev_us += s_.intersect_device_us + s_.score_device_us
    torch.cuda.synchronize()
    wall_seq = (time.perf_counter() - t0) / steps
    launches = ps.stats(reset=True).kernel_launches
    hits = sum(results[q][2] for q in POSTING_QUERIES)
    dev_s = dev_us * 1e-6 / steps
    # the same query set through the batch entry point: the 8 searches are spread over a pool of streams inside
    # the library (what a dispatch shim does with concurrent FT.SEARCHes)
    def term_params(q):
        return [(1.0, L.II_CalculateIDF(n_docs, len(host_lists[r][0])), L.II_CalculateIDF_BM25(n_docs, len(host_lists[r][0]))) for r in q]

    class _H:  # SearchBatch wants objects with a .h handle
        def __init__(self, h):
            self.h = h

    batch = ps.SearchBatch([([_H(lists[r]) for r in q], term_params(q)) for q in POSTING_QUERIES], 10)
    for _ in range(3):
        conc = batch.run(False, ps.SCORER_BM25STD, 1.0, n_docs, avg_len, dt)
    reps = max(steps, 5) * 4
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        conc = batch.run(False, ps.SCORER_BM25STD, 1.0, n_docs, avg_len, dt)
    wall = (time.perf_counter() - t0) / reps
    for q, r_ in zip(POSTING_QUERIES, conc):
        assert r_[0].tolist() == results[q][0].tolist() and r_[2] == results[q][2]
    # e2e: encoded IndexBlocks on the host -> decode (all cores) -> H2D -> AND + BM25STD + top-10 -> host
    enc = {}
    for r in ranks:
        ids, fr = host_lists[r]
        n = len(ids)
        nb = n // 100 + 2
        out = np.zeros(n * 9 + 64, dtype=np.uint8)
        first, last = np.zeros(nb, dtype=np.uint64), np.zeros(nb, dtype=np.uint64)
        bn, off = np.zeros(nb, dtype=np.uint16), np.zeros(nb + 1, dtype=np.uint64)
        nblocks = C.c_size_t(0)
        S.Synth_EncodeFreqsOnlyBlocks(ids.ctypes.data, fr.ctypes.data, n, out.ctypes.data, first.ctypes.data, last.ctypes.data,
                                      bn.ctypes.data, off.ctypes.data, C.byref(nblocks))
        views = (ps.II_BlockView * nblocks.value)()
        base = out.ctypes.data
        for b in range(nblocks.value):
            views[b] = ps.II_BlockView(int(first[b]), int(last[b]), int(bn[b]), C.cast(base + int(off[b]), C.POINTER(C.c_uint8)), int(off[b + 1] - off[b]))
        enc[r] = (views, nblocks.value, out, int(off[nblocks.value]))
    enc_bytes = sum(enc[r][3] for q in POSTING_QUERIES for r in q)
    def e2e_pass(on_device):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        dec_us = 0.0
        for q in POSTING_QUERIES:
            hs = []
            for r in q:
                h = L.II_PostingList_FromBlocks(enc[r][0], enc[r][1], ps.CODEC_FREQS_ONLY, 0, on_device)
                dec_us += ps.stats(reset=False).decode_host_us
                hs.append(h)
            e_ids, e_sc, _ = run_query(q, hs)
            assert e_ids.tolist() == results[q][0].tolist()
            for h in hs:
                L.II_PostingList_Free(h)
        return time.perf_counter() - t0, dec_us

    e2e_pass(1)  # warm the pinned staging
    e2e_wall, decode_us = e2e_pass(1)
    e2e_wall_host, decode_us_host = e2e_pass(0)
    alg_bytes = in_postings * 8 + hits * 16
    peak, _ = load_peaks()
    for h in lists.values():
        L.II_PostingList_Free(h)
    L.II_DocTable_Free(dt)
    return {
        "metric": "BM25 intersect docs/sec", "value": in_postings / wall, "unit": "input postings/s",
        "matched_docs_per_s": hits / wall, "ms_per_query_set": wall * 1000.0, "gpu_launches": int(launches),
        "api": "II_SearchTopNBatch (8 queries per call, pool of 8 streams)",
        "sequential": {"value": in_postings / wall_seq, "ms_per_query_set": wall_seq * 1000.0,
                       "note": "II_SearchTopN, one query at a time"},
        "config": {"workload": f"3-term AND + BM25STD + top-10 over a {n_docs}-doc synthetic Zipf index, {len(POSTING_QUERIES)} queries "
                               f"(rank triples {POSTING_QUERIES}), postings resident in HBM, one II_SearchTopNBatch call per set", "input_postings": in_postings,
                   "matched_docs": hits},
        "e2e": {"value": in_postings / e2e_wall, "unit": "input postings/s", "h2d_bytes_per_step": in_postings * 8,
                "d2h_bytes_per_step": len(POSTING_QUERIES) * 10 * 16, "encoded_bytes": enc_bytes,
                "host_gather_ms": decode_us / 1000.0, "ms_per_query_set": e2e_wall * 1000.0,
                "note": "FreqsOnly IndexBlocks on the host -> II_PostingList_FromBlocks (gather to pinned, H2D of the encoded bytes, "
                        "decode_blocks_kernel) -> II_SearchTopN, one query at a time",
                "host_decode_variant": {"value": in_postings / e2e_wall_host, "ms_per_query_set": e2e_wall_host * 1000.0,
                                        "host_decode_ms": decode_us_host / 1000.0}},
        "roofline": {"bound": "hbm", "achieved": alg_bytes / dev_s / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_bytes / dev_s / 1e9 / peak, "traffic": None, "kernel": "intersect_kernel + gather_kernel + score_kernel",
                     "device_ms_per_query_set": dev_s * 1000.0, "algorithmic_bytes": alg_bytes},
    }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS, help="rows per GPU (default = the BASELINE workload)")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--ref-sample-rows", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-postings", action="store_true")
    ap.add_argument("--posting-docs", type=int, default=50_000_000)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from redisearch_b200 import vecsim as vs
    from redisearch_b200._lib import load_library

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    rows, nq = args.rows, args.batch

    L = vs.lib()
    S = load_library("libsynth_b200.so")
    S.Synth_FillRows.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
    S.Synth_NormalizeRowsF32.argtypes = [C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32, C.c_void_p]
    # an explicit stream: a NULL handle would mean CUDA's legacy default stream
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sp = C.c_void_p(stream.cuda_stream)
    assert sp.value

    # ---- build the shard in HBM: generate + normalise on device, ingest device-to-device
    index = vs.VecSimIndex(vs.VecSimType_FLOAT32, DIM, vs.VecSimMetric_Cosine)
    assert L.VecSimB200_Reserve(index.h, rows) == 0, "cannot reserve HBM for the corpus"
    chunk = min(rows, 1_000_000)
    buf = torch.empty((chunk, DIM), dtype=torch.float32, device=dev)
    row0_global = rank * rows
    t_build = time.perf_counter()
    done = 0
    while done < rows:
        n = min(chunk, rows - done)
        assert S.Synth_FillRows(buf.data_ptr(), DIM * 4, 0, SEED_ROWS, row0_global + done, n, DIM, sp) == 0
        assert S.Synth_NormalizeRowsF32(buf.data_ptr(), DIM * 4, n, DIM, sp) == 0
        torch.cuda.synchronize()
        assert L.VecSimB200_AddVectorsDevice(index.h, buf.data_ptr(), n, row0_global + done + 1) == n
        done += n
    del buf
    build_s = time.perf_counter() - t_build

    # ---- queries: same generator, normalised like VecSimIndex_TopKQuery would (preprocessors.h:121-131)
    qdev = torch.empty((nq, DIM), dtype=torch.float32, device=dev)
    assert S.Synth_FillRows(qdev.data_ptr(), DIM * 4, 0, SEED_QUERIES, 0, nq, DIM, sp) == 0
    q_host_raw = qdev.cpu().numpy().copy()  # raw (un-normalised) host blobs for the e2e arm
    assert S.Synth_NormalizeRowsF32(qdev.data_ptr(), DIM * 4, nq, DIM, sp) == 0
    out_labels = torch.empty((nq, K), dtype=torch.int64, device=dev)
    out_scores = torch.empty((nq, K), dtype=torch.float32, device=dev)
    if world > 1:
        gath_s = torch.empty((world, nq, K), dtype=torch.float32, device=dev)
        gath_l = torch.empty((world, nq, K), dtype=torch.int64, device=dev)
        fin_s = torch.empty((nq, K), dtype=torch.float32, device=dev)
        fin_l = torch.empty((nq, K), dtype=torch.int64, device=dev)

    def step_device():
        rc = L.VecSimB200_TopKQueryBatchDevice(index.h, qdev.data_ptr(), nq, K, out_labels.data_ptr(),
                                               out_scores.data_ptr(), sp)
        assert rc == 0
        if world > 1:  # the one exchange step: per-shard top-k -> all ranks -> G-way merge on device
            dist.all_gather_into_tensor(gath_s, out_scores)
            dist.all_gather_into_tensor(gath_l, out_labels)
            assert L.VecSimB200_MergeShardTopK(gath_s.data_ptr(), gath_l.data_ptr(), world, nq, K, fin_s.data_ptr(),
                                               fin_l.data_ptr(), sp) == 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step_device()
    barrier()
    index.stats(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step_device()
        ev1.record(stream)
        barrier()
    ms_total = ev0.elapsed_time(ev1)
    st = index.stats(reset=True)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * nq / (ms_step / 1000.0)

    # ---- e2e through the host-facing C-ABI: host blobs in, host results out
    h_labels = np.empty((nq, K), dtype=np.uint64)
    h_scores = np.empty((nq, K), dtype=np.float64)
    qh = np.ascontiguousarray(q_host_raw)
    for _ in range(2):
        assert L.VecSimB200_TopKQueryBatch(index.h, qh.ctypes.data, qh.strides[0], nq, K, None, h_labels.ctypes.data,
                                           h_scores.ctypes.data) == 0
    e2e_steps = max(3, min(args.steps, 10))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert L.VecSimB200_TopKQueryBatch(index.h, qh.ctypes.data, qh.strides[0], nq, K, None, h_labels.ctypes.data,
                                           h_scores.ctypes.data) == 0
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * nq / e2e_s
    # host-path and device-path answers must agree (same kernels)
    dl = out_labels.cpu().numpy()
    agree = bool((dl == h_labels.astype(np.int64)).all())

    # ---- B=1 through the stock VecSimIndex_TopKQuery (what hybrid_reader.c:374 calls).  Two legs: as served (the fp16
    # shadow built by the batches above is current, so a single query rides the tensor-core route too) and the exact
    # HBM-bound scan alone (coarse mode off) — the north_star's ">= 10x CPU at >= 70% of HBM roofline" figure
    def single_query_leg():
        index.stats(reset=True)
        q1 = np.ascontiguousarray(q_host_raw[0])
        for _ in range(3):
            index.topk(q1, K)
        index.stats(reset=True)
        n1 = 20
        barrier()
        t0 = time.perf_counter()
        for i in range(n1):
            index.topk(np.ascontiguousarray(q_host_raw[i % nq]), K)
        dt = (time.perf_counter() - t0) / n1
        st1 = index.stats(reset=True)
        return dt, st1.scan_device_us / max(1, st1.scan_launches), int(L.VecSimB200_LastBatchPath(index.h))

    b1s_s, b1s_scan_us, b1s_path = single_query_leg()
    L.VecSimB200_SetCoarseMode(0)
    b1_s, b1_scan_us, _ = single_query_leg()
    L.VecSimB200_SetCoarseMode(-1)
    b1_bytes = rows * DIM * 4 + DIM * 4 + K * 12

    # ---- roofline of the dominant kernel: the library brackets it with CUDA events on the launch stream
    # (VecSimB200_GetStats); one launch per synchronised call so that every interval is one kernel
    peak, peak_src = load_peaks()
    index.stats(reset=True)
    per_launch = []
    for _ in range(max(3, min(args.steps, 10))):
        step_device()
        torch.cuda.synchronize()
        s_i = index.stats(reset=True)
        if s_i.scan_launches:
            per_launch.append(s_i.scan_device_us / s_i.scan_launches)
    scan_us = float(np.mean(per_launch)) if per_launch else None
    coarse_mode = int(os.environ.get("VECSIM_B200_COARSE", "1"))
    dom_kernel = {0: "scan_topk_kernel<f32,IP,4,8>",
                  1: "coarse_qtmem_kernel (tcgen05 kind::f16 TS-mode, queries in TMEM, fp16 shadow rows)",
                  2: "coarse_kernel<CfgTF32> (tcgen05 kind::tf32)"}.get(coarse_mode, "?")
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f).get(dom_kernel.split(" ")[0].split("<")[0])
        if t and (t["rows"], t["dim"], t["batch"]) == (rows, DIM, nq):
            traffic = t["dram_bytes_per_launch"]
    except (OSError, ValueError, KeyError):
        pass
    alg_bytes = rows * DIM * 4 + nq * DIM * 4 + nq * K * 12  # N*D*s per corpus pass
    read_bytes = rows * DIM * (2 if coarse_mode == 1 else 4)   # what this kernel has to pull from HBM once
    achieved = alg_bytes / (scan_us * 1e-6) / 1e9 if scan_us else None

    hbm_view = {"achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "algorithmic_bytes_per_launch": alg_bytes, "hbm_bytes_read_per_launch": read_bytes,
                "frac_on_bytes_read": (read_bytes / (scan_us * 1e-6) / 1e9 / peak) if scan_us else None}
    common = {"traffic": traffic, "kernel": dom_kernel, "avg_launch_us": scan_us, "launch_us_samples": per_launch,
              "share_of_step": (scan_us / 1000.0 / ms_step) if scan_us else None}
    if coarse_mode in (1, 2) and scan_us:
        # batch of 256: 128 flop per fp32 corpus byte — the pass is tensor-bound, HBM view kept beside it
        flops = 2.0 * nq * rows * DIM
        tpeak, tsust, tsrc = load_tensor_peak()
        if coarse_mode == 2:
            tpeak, tsust = tpeak / 2, tsust / 2  # TF32 runs at half the 16-bit rate
        tf = flops / (scan_us * 1e-6) / 1e12
        roofline = dict(common, bound="tensor", achieved=tf, peak=tpeak, unit="TFLOP/s", frac=tf / tpeak,
                        flops_per_launch=flops, peak_source=tsrc + " — burst dense bf16 (cuBLAS), kernel timed alone",
                        peak_sustained=tsust, hbm=dict(hbm_view, peak_source=peak_src),
                        note="sustained runs sit at the 1000 W board power cap (clocks.reasons)")
    else:
        roofline = dict(common, bound="hbm", peak_source=peak_src, **hbm_view)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"FLAT {rows} x {DIM} fp32 cosine k={K} batch={nq} per GPU"
                                   + (f"; corpus = {world} shards, NCCL all-gather of per-shard top-k + device merge" if world > 1 else ""),
                       "rows_per_gpu": rows, "dim": DIM, "k": K, "batch": nq,
                       "l2_policy": "corpus (30.7 GB) >> 126 MB L2, no flush needed",
                       "value_unit_note": "queries x 10M-row shards per second" if world > 1 else "queries per second",
                       "build_seconds": round(build_s, 2), "host_device_results_agree": agree},
            "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(nq * DIM * 4),
                    "d2h_bytes_per_step": int(nq * K * 8), "ms_per_step": e2e_s * 1000.0},
            "gpu_launches": int(st.kernel_launches),
            "roofline": roofline,
            "clocks": clocks.summary(),
            "single_query_as_served": {"api": "VecSimIndex_TopKQuery with the fp16 shadow current", "value": 1.0 / b1s_s,
                                       "unit": "queries/s", "ms_per_query": b1s_s * 1000.0, "route": b1s_path,
                                       "dominant_kernel_us": b1s_scan_us},
            "single_query": {"api": "VecSimIndex_TopKQuery (host blob in, reply out), exact scan only", "value": 1.0 / b1_s,
                             "unit": "queries/s", "ms_per_query": b1_s * 1000.0,
                             "roofline": {"bound": "hbm", "achieved": b1_bytes / (b1_scan_us * 1e-6) / 1e9,
                                          "peak": peak, "unit": "GB/s",
                                          "frac": b1_bytes / (b1_scan_us * 1e-6) / 1e9 / peak,
                                          "kernel": "scan_topk_kernel<f32,IP,4,1>", "avg_launch_us": b1_scan_us,
                                          "algorithmic_bytes_per_launch": b1_bytes}},
        }
        if world == 1 and not args.no_postings:
            del index
            torch.cuda.empty_cache()
            try:
                line["bm25_intersect"] = bench_postings(torch, dev, sp, args.posting_docs, max(1, min(args.steps, 5)))
            except Exception as e:
                line["bm25_intersect"] = {"value": None, "error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                _, info = cpu_reference_qps(args.ref_sample_rows, 2, os.cpu_count() or 1)
                line["cpu_baseline"] = info
                if isinstance(line.get("bm25_intersect"), dict) and line["bm25_intersect"].get("value"):
                    line["bm25_intersect"]["cpu_baseline"] = cpu_postings_baseline(10_000_000, os.cpu_count() or 1)
            except Exception as e:  # the baseline is a reported side number; never fail the bench on it
                line["cpu_baseline"] = {"value": None, "error": repr(e)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
