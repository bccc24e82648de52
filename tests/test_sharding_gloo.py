"""N>1 path on CPU: two gloo ranks, each owning a row range of the corpus, exchange per-shard top-k lists
with one all-gather; the merged answer must equal the single-index answer of the oracle.  The per-shard
scan is done by the oracle here (no GPU in this tier); what is under test is the sharding arithmetic, the
label bookkeeping, the exchange and the (score,label) merge rule the CUDA merge kernel implements
(the kernel itself is checked in test_vecsim_sharded.py on the GPU)."""
import os
import socket
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import oracle_lib as ol
from redisearch_b200 import sharding

def merge_reference(gs, gl, k):
    """(score asc, label asc) over the union of the per-shard lists; label -1 = empty."""
    G, B, _ = gs.shape
    out_s = np.full((B, k), np.nan, dtype=np.float32); out_l = np.full((B, k), -1, dtype=np.int64)
    for b in range(B):
        items = [(float(gs[g, b, i]), int(gl[g, b, i])) for g in range(G) for i in range(gs.shape[2]) if gl[g, b, i] >= 0]
        items.sort()
        for i, (s, l) in enumerate(items[:k]):
            out_s[b, i], out_l[b, i] = s, l
    return out_s, out_l

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, DIM, K, B = 5000, 48, 10, 7
lo, hi = sharding.shard_range(N, world, rank)
rows = ol.synth_rows(ol.F32, 42, lo, hi - lo, DIM)           # this rank's rows only
shard = ol.PortIndex(ol.F32, DIM, ol.COS)
shard.add_many(rows, lo + 1)                                  # labels travel with the rows
qs = ol.synth_rows(ol.F32, 43, 0, B, DIM)
ls = np.full((B, K), np.nan, dtype=np.float32); ll = np.full((B, K), -1, dtype=np.int64)
for b in range(B):
    ids, sc = shard.topk(qs[b], K)
    ls[b, :len(ids)], ll[b, :len(ids)] = sc, ids
gs, gl = sharding.allgather_topk(torch.from_numpy(ls), torch.from_numpy(ll))
ms, ml = merge_reference(gs.numpy(), gl.numpy(), K)
if rank == 0:
    full = ol.PortIndex(ol.F32, DIM, ol.COS)
    full.add_many(ol.synth_rows(ol.F32, 42, 0, N, DIM), 1)
    for b in range(B):
        ids, sc = full.topk(qs[b], K)
        assert ml[b].tolist() == ids.tolist(), (b, ml[b], ids)
        assert ms[b].tobytes() == sc.astype(np.float32).tobytes()
    assert sum(sharding.shard_range(N, world, r)[1] - sharding.shard_range(N, world, r)[0] for r in range(world)) == N
    print("SHARDING-OK")
dist.destroy_process_group()
'''


def test_two_rank_gloo_exchange_and_merge(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARDING-OK" in r.stdout


POSTINGS_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import oracle_lib as ol
from redisearch_b200 import sharding

def local_topn(lists, freqs, terms, doc_len, n_docs, avg, n):
    """AND + BM25STD + top-n of this shard's slices, on the CPU oracle (the GPU does this step in production)."""
    idx = [ol.InvIndex(ol.CODEC_FREQS_ONLY, l, f) for l, f in zip(lists, freqs)]
    hits = ol.run_intersect(idx) if all(len(l) for l in lists) else []
    scored = []
    for doc, ch in hits:
        s = ol.oracle_score(ol.SCORER_BM25STD, [f for _, f in ch], [terms[c][1] for c, _ in ch], [terms[c][2] for c, _ in ch],
                            [1.0] * len(ch), 1.0, int(doc_len[doc]), 1, 1.0, n_docs, avg)
        scored.append((-s, doc))
    scored.sort()
    return [(d, -s) for s, d in scored[:n]], len(hits)

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(3)                      # same stream on every rank: the GLOBAL index
n_docs, TOP = 60_000, 10
lists = [np.unique(rng.integers(1, n_docs + 1, m)).astype(np.uint64) for m in (30_000, 12_000, 20_000)]
freqs = [rng.integers(1, 9, len(l)).astype(np.uint32) for l in lists]
doc_len = rng.integers(50, 500, n_docs + 1).astype(np.uint32)
avg = float(doc_len[1:].mean())
# statistics are global: idf from the full lists, avgDocLen over all documents
terms = [(1.0, ol.postings().orc_idf(n_docs, len(l)), ol.postings().orc_idf_bm25(n_docs, len(l))) for l in lists]
lo, hi = sharding.doc_range(n_docs, world, rank)
mine = [sharding.split_posting_list(l, f, lo, hi) for l, f in zip(lists, freqs)]
assert all((len(l) == 0 or (l[0] > lo and l[-1] <= hi)) for l, _ in mine)
top, nhits = local_topn([l for l, _ in mine], [f for _, f in mine], terms, doc_len, n_docs, avg, TOP)
sc = torch.full((TOP,), float("nan"), dtype=torch.float64); ids = torch.full((TOP,), -1, dtype=torch.int64)
for i, (d, s) in enumerate(top):
    sc[i], ids[i] = s, d
gs, gi, gc = sharding.allgather_topn(sc, ids, len(top))
m_ids, m_sc = sharding.merge_topn(gs, gi, gc, TOP)
tot = torch.tensor([nhits]); dist.all_reduce(tot)
if rank == 0:
    full, full_hits = local_topn(lists, freqs, terms, doc_len, n_docs, avg, TOP)
    assert int(tot.item()) == full_hits
    assert m_ids.tolist() == [d for d, _ in full], (m_ids, full)
    assert m_sc.tobytes() == np.array([s for _, s in full], dtype=np.float64).tobytes()
    print("POSTINGS-SHARDING-OK")
dist.destroy_process_group()
'''


def test_two_rank_gloo_posting_shards(tmp_path):
    """Postings sharded by docId range: per-shard AND + BM25STD + top-N with global statistics,
    one all-gather, II_MergeShardTopN — must equal the unsharded answer (ids and score bits)."""
    script = tmp_path / "pworker.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + POSTINGS_WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "POSTINGS-SHARDING-OK" in r.stdout


def test_shard_ranges_partition_exactly():
    sys.path.insert(0, ROOT)
    from redisearch_b200 import sharding

    for n in (0, 1, 7, 1000, 10_000_000, 50_000_001):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_posting_lists_split_exactly_at_the_docid_boundaries():
    sys.path.insert(0, ROOT)
    from redisearch_b200 import sharding

    rng = np.random.default_rng(0)
    n_docs = 100_003
    ids = np.unique(rng.integers(1, n_docs + 1, 40_000)).astype(np.uint64)
    fr = rng.integers(1, 9, len(ids)).astype(np.uint32)
    for world in (1, 2, 3, 8):
        parts = []
        for r in range(world):
            lo, hi = sharding.doc_range(n_docs, world, r)
            a, f = sharding.split_posting_list(ids, fr, lo, hi)
            assert len(a) == len(f) and (len(a) == 0 or (a[0] > lo and a[-1] <= hi))
            parts.append((a, f))
        assert np.concatenate([a for a, _ in parts]).tolist() == ids.tolist()   # disjoint, ordered, complete
        assert np.concatenate([f for _, f in parts]).tolist() == fr.tolist()
    # a docId equal to a boundary belongs to the lower shard: ranges are (lo, hi]
    lo, hi = sharding.doc_range(10, 2, 0)
    a, _ = sharding.split_posting_list(np.array([hi, hi + 1], dtype=np.uint64), None, lo, hi)
    assert a.tolist() == [hi]
