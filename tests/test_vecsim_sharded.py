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
