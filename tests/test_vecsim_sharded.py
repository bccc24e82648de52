"""GPU: the shard-merge kernel (VecSimB200_MergeShardTopK) and a 2-shard KNN on one device."""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _merge_reference(gs, gl, k):
    G, B, _ = gs.shape
    out_s = np.full((B, k), np.nan, dtype=np.float32)
    out_l = np.full((B, k), -1, dtype=np.int64)
    for b in range(B):
        items = sorted((float(gs[g, b, i]), int(gl[g, b, i])) for g in range(G) for i in range(gs.shape[2]) if gl[g, b, i] >= 0)
        for i, (s, l) in enumerate(items[:k]):
            out_s[b, i], out_l[b, i] = s, l
    return out_s, out_l


@pytest.mark.parametrize("G,B,k", [(2, 5, 10), (8, 33, 10), (4, 3, 100), (8, 2, 128), (3, 1, 1)])
def test_merge_kernel_matches_reference(G, B, k):
    import torch

    from redisearch_b200 import sharding

    rng = np.random.default_rng(G * 100 + k)
    gs = np.sort(rng.choice(np.array([0.1, 0.25, 0.5, 0.75, 1.5], dtype=np.float32), (G, B, k)) + rng.integers(0, 3, (G, B, k)).astype(np.float32), axis=2)
    gl = rng.permutation(G * B * k).reshape(G, B, k).astype(np.int64)
    gl[rng.random((G, B, k)) < 0.15] = -1  # short lists
    ms, ml = sharding.merge_topk_device(torch.from_numpy(gs).cuda(), torch.from_numpy(gl).cuda())
    torch.cuda.synchronize()
    rs, rl = _merge_reference(gs, gl, k)
    assert ml.cpu().numpy().tolist() == rl.tolist()
    a, b = ms.cpu().numpy(), rs
    assert ((a == b) | (np.isnan(a) & np.isnan(b))).all()


def test_two_shards_on_one_gpu_equal_single_index():
    """Row-range shards + merge == one index (exactness of the exchange)."""
    import torch

    from redisearch_b200 import sharding
    from redisearch_b200 import vecsim as vs

    N, DIM, K, B, G = 20000, 64, 10, 9, 3
    rows = ol.synth_rows(ol.F32, 42, 0, N, DIM)
    qs = ol.synth_rows(ol.F32, 43, 0, B, DIM)
    full = ol.PortIndex(ol.F32, DIM, ol.COS)
    full.add_many(rows, 1)
    gs = np.zeros((G, B, K), dtype=np.float32)
    gl = np.zeros((G, B, K), dtype=np.int64)
    for g in range(G):
        lo, hi = sharding.shard_range(N, G, g)
        ix = vs.VecSimIndex(vs.VecSimType_FLOAT32, DIM, vs.VecSimMetric_Cosine)
        ix.add_many(rows[lo:hi], label0=lo + 1)
        labels, scores, rc = ix.topk_batch(qs, K)
        assert rc == 0
        gs[g], gl[g] = scores.astype(np.float32), labels.astype(np.int64)
    ms, ml = sharding.merge_topk_device(torch.from_numpy(gs).cuda(), torch.from_numpy(gl).cuda())
    torch.cuda.synchronize()
    for b in range(B):
        ids, sc = full.topk(qs[b], K)
        assert ml[b].cpu().numpy().tolist() == ids.tolist()
        assert ms[b].cpu().numpy().tobytes() == sc.astype(np.float32).tobytes()


def test_config3_shape_fp16_ip_k100_batch_sharded():
    """BASELINE.json configs[2] in miniature: fp16 rows, raw inner product, k=100, a query batch, row-range shards,
    device merge.  fp16 bar: positions within 1e-2, ids equal modulo candidates within that tolerance of the k-th."""
    import torch

    from redisearch_b200 import sharding
    from redisearch_b200 import vecsim as vs

    N, DIM, K, B, G = 24_000, 768, 100, 24, 4
    rows = ol.synth_rows(ol.F16, 42, 0, N, DIM)
    qs = ol.synth_rows(ol.F16, 43, 0, B, DIM)
    full = ol.PortIndex(ol.F16, DIM, ol.IP, tier=ol.TIER_AVX512)  # fp32-accumulate tier (SURVEY finding 5)
    full.add_many(rows, 1)
    gs = np.zeros((G, B, K), dtype=np.float32)
    gl = np.zeros((G, B, K), dtype=np.int64)
    for g in range(G):
        lo, hi = sharding.shard_range(N, G, g)
        ix = vs.VecSimIndex(vs.VecSimType_FLOAT16, DIM, vs.VecSimMetric_IP)
        ix.add_many(rows[lo:hi], label0=lo + 1)
        labels, scores, rc = ix.topk_batch(qs, K)
        assert rc == 0
        gs[g], gl[g] = scores.astype(np.float32), labels.astype(np.int64)
    ms, ml = sharding.merge_topk_device(torch.from_numpy(gs).cuda(), torch.from_numpy(gl).cuda())
    torch.cuda.synchronize()
    ms, ml = ms.cpu().numpy(), ml.cpu().numpy()
    for b in range(B):
        ids, sc = full.topk(qs[b], K)
        assert len(ids) == K
        for a, e in zip(ms[b], sc):
            assert abs(a - e) <= 1e-2 * max(abs(e), 1.0) + 1e-6
        kth = sc[-1]
        slack = 1e-2 * max(abs(kth), 1.0) + 1e-6
        sure = {i for i, s in zip(ids.tolist(), sc.tolist()) if s < kth - slack}
        assert sure <= set(ml[b].tolist())


def test_config5_shape_hybrid_filter_sharded():
    """BASELINE.json configs[4] in miniature: rows AND the filter's docIds are cut at the same shard boundaries
    (no cross-GPU gather), every shard answers VecSimB200_TopKFiltered on its slice, one merge."""
    import torch

    from redisearch_b200 import sharding
    from redisearch_b200 import vecsim as vs

    N, DIM, K, G = 40_000, 128, 10, 4
    rows = ol.synth_rows(ol.F32, 42, 0, N, DIM)
    q = ol.synth_rows(ol.F32, 43, 0, 1, DIM)[0]
    qn = q.copy()
    ol.port().orc_normalize(ol._p(qn), DIM, ol.F32)
    full = ol.PortIndex(ol.F32, DIM, ol.COS, tier=ol.TIER_AVX512)
    full.add_many(rows, 1)
    rng = np.random.default_rng(8)
    a = np.unique(rng.integers(1, N + 1, 25_000)).astype(np.uint64)
    b = np.unique(rng.integers(1, N + 1, 18_000)).astype(np.uint64)
    filt = np.intersect1d(a, b).astype(np.uint32)  # the 2-term AND (its device evaluation is checked elsewhere)
    gs = np.full((G, 1, K), np.nan, dtype=np.float32)
    gl = np.full((G, 1, K), -1, dtype=np.int64)
    for g in range(G):
        lo, hi = sharding.shard_range(N, G, g)       # rows [lo, hi) carry labels lo+1 .. hi
        ix = vs.VecSimIndex(vs.VecSimType_FLOAT32, DIM, vs.VecSimMetric_Cosine)
        ix.add_many(rows[lo:hi], label0=lo + 1)
        mine, _ = sharding.split_posting_list(filt, None, lo, hi)
        labels, scores, rc = ix.topk_filtered(q, K, mine)
        assert rc == 0
        gs[g, 0, :len(labels)], gl[g, 0, :len(labels)] = scores, labels
    ms, ml = sharding.merge_topk_device(torch.from_numpy(gs).cuda(), torch.from_numpy(gl).cuda())
    torch.cuda.synchronize()
    exp = sorted((np.float32(full.distance_from(int(d), qn)), int(d)) for d in filt.tolist())[:K]
    assert ml[0].cpu().numpy().tolist() == [d for _, d in exp]
    assert ms[0].cpu().numpy().tobytes() == np.array([s for s, _ in exp], dtype=np.float32).tobytes()


@pytest.mark.parametrize("G,B,k", [(2, 5, 10), (8, 33, 10), (4, 3, 100), (3, 7, 1)])
def test_packed_exchange_blocks_merge_like_the_dense_arrays(G, B, k):
    """The shard group's exchange format — one block [labels int64 x B*k][scores float x B*k] per rank, padded to 16
    bytes, rank-major after the single all-gather — merges to the same answer as the dense [G][B][k] arrays."""
    import ctypes as C

    import torch

    from redisearch_b200 import vecsim as vs

    L = vs.lib()
    rng = np.random.default_rng(G * 7 + k)
    gs = np.sort(rng.random((G, B, k)).astype(np.float32), axis=2)
    gl = rng.permutation(G * B * k).reshape(G, B, k).astype(np.int64)
    gl[rng.random((G, B, k)) < 0.2] = -1
    block = int(L.VecSimB200_ShardBlockBytes(B, k))
    assert block % 16 == 0 and block >= B * k * 12
    buf = np.zeros(G * block, dtype=np.uint8)
    for g in range(G):
        buf[g * block:g * block + B * k * 8] = gl[g].reshape(-1).view(np.uint8)
        buf[g * block + B * k * 8:g * block + B * k * 12] = gs[g].reshape(-1).view(np.uint8)
    d = torch.from_numpy(buf).cuda()
    out_s = torch.empty((B, k), dtype=torch.float32, device="cuda")
    out_l = torch.empty((B, k), dtype=torch.int64, device="cuda")
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.VecSimB200_MergeShardBlocks(d.data_ptr(), G, B, k, out_s.data_ptr(), out_l.data_ptr(), sp) == 0
    torch.cuda.synchronize()
    rs, rl = _merge_reference(gs, gl, k)
    assert out_l.cpu().numpy().tolist() == rl.tolist()
    a = out_s.cpu().numpy()
    assert ((a == rs) | (np.isnan(a) & np.isnan(rs))).all()


def test_shard_group_of_one_answers_like_the_index():
    """world = 1: the collective entry points degenerate to the local scan (no NCCL is loaded)."""
    from redisearch_b200 import vecsim as vs

    L = vs.lib()
    n, dim, k, nq = 30_000, 64, 10, 12
    rows = ol.synth_rows(ol.F32, 42, 0, n, dim)
    qs = ol.synth_rows(ol.F32, 43, 0, nq, dim)
    ix = vs.VecSimIndex(vs.VecSimType_FLOAT32, dim, vs.VecSimMetric_Cosine)
    ix.add_many(rows, label0=1)
    g = L.VecSimB200_ShardGroup_New(None, 0, 1)
    assert g and L.VecSimB200_ShardGroup_Size(g) == 1 and L.VecSimB200_ShardGroup_Rank(g) == 0
    labels = np.zeros((nq, k), dtype=np.uint64)
    scores = np.zeros((nq, k), dtype=np.float64)
    assert L.VecSimB200_ShardGroup_TopKBatch(g, ix.h, qs.ctypes.data, qs.strides[0], nq, k, labels.ctypes.data, scores.ctypes.data) == 0
    el, es, rc = ix.topk_batch(qs, k)
    assert rc == 0 and (labels == el).all() and (scores.astype(np.float32) == es.astype(np.float32)).all()
    L.VecSimB200_ShardGroup_Free(g)


NCCL_SCRIPT = r'''
import os, sys, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol
from redisearch_b200 import vecsim as vs, sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("gloo")          # control plane only: ships the 128-byte NCCL id
L = vs.lib()
idbuf = np.zeros(128, dtype=np.uint8)
if rank == 0:
    assert L.VecSimB200_ShardGroup_UniqueId(idbuf.ctypes.data) == 0
t = torch.from_numpy(idbuf); dist.broadcast(t, 0)
g = L.VecSimB200_ShardGroup_New(idbuf.ctypes.data, rank, world)
assert g, "ncclCommInitRank failed"
N, DIM, K, B = 200_000, 96, 10, 40
lo, hi = sharding.shard_range(N, world, rank)
rows = ol.synth_rows(ol.F32, 42, lo, hi - lo, DIM)
qs = ol.synth_rows(ol.F32, 43, 0, B, DIM)
ix = vs.VecSimIndex(vs.VecSimType_FLOAT32, DIM, vs.VecSimMetric_Cosine)
ix.add_many(rows, label0=lo + 1)
labels = np.zeros((B, K), dtype=np.uint64); scores = np.zeros((B, K), dtype=np.float64)
for rep in range(3):
    assert L.VecSimB200_ShardGroup_TopKBatch(g, ix.h, qs.ctypes.data, qs.strides[0], B, K, labels.ctypes.data, scores.ctypes.data) == 0
# every rank checks the merged answer against the oracle over the whole corpus
full = ol.PortIndex(ol.F32, DIM, ol.COS, tier=ol.TIER_AVX512)
full.add_many(ol.synth_rows(ol.F32, 42, 0, N, DIM), 1)
for i in range(B):
    pi, ps = full.topk(qs[i], K)
    assert labels[i].astype(np.int64).tolist() == pi.tolist(), (rank, i)
    assert scores[i].astype(np.float32).tobytes() == ps.astype(np.float32).tobytes()
L.VecSimB200_ShardGroup_Free(g)
dist.barrier()
if rank == 0: print("SHARDGROUP-NCCL-OK")
'''


def test_shard_group_over_nccl_two_ranks(tmp_path):
    """Two processes, two GPUs: local scans, ONE ncclAllGather of the packed blocks inside the library, device merge;
    every rank's merged answer equals the oracle's over the unsharded corpus.  Skipped on a single-GPU box."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "sg.py"
    script.write_text(f"ROOT = {root!r}\n" + NCCL_SCRIPT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "SHARDGROUP-NCCL-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
