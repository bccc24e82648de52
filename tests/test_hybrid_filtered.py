"""GPU: the fused hybrid query VecSimB200_TopKFiltered against the reference's ad-hoc hybrid loop.

HybridIterator in HYBRID_ADHOC_BF mode (src/iterators/hybrid_reader.c:289-335) reads the filter's docIds in
ascending order, asks VecSimIndex_GetDistanceFrom_Unsafe for each, skips NaN (deleted) and keeps the k best in a
heap with strict `<` admission — i.e. the k smallest by (distance, docId).  The oracle restates that loop with the
port's / the reference's distance_from; ids and fp32 distance bits must match.  The filter itself comes from
libii_b200 (a 2-term AND on the device: BASELINE config C5) in the device-pointer variant.
"""
import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu


def _oracle_adhoc(p, q, doc_ids, k):
    best = []
    for d in doc_ids.tolist():
        s = p.distance_from(int(d), q)
        if s != s:
            continue
        best.append((np.float32(s), d))
    best.sort(key=lambda t: (t[0], t[1]))
    return best[:k]


@pytest.mark.parametrize("vtype,metric", [(ol.F32, ol.COS), (ol.F32, ol.L2), (ol.I8, ol.COS), (ol.F16, ol.IP)])
def test_topk_filtered_matches_adhoc_loop(vtype, metric):
    from redisearch_b200 import vecsim as vs

    n, dim, k = 20_000, 96, 10
    rows = ol.synth_rows(vtype, 42, 0, n, dim)
    g = vs.VecSimIndex({ol.F32: vs.VecSimType_FLOAT32, ol.I8: vs.VecSimType_INT8, ol.F16: vs.VecSimType_FLOAT16}[vtype], dim,
                       {ol.COS: vs.VecSimMetric_Cosine, ol.L2: vs.VecSimMetric_L2, ol.IP: vs.VecSimMetric_IP}[metric])
    p = ol.PortIndex(vtype, dim, metric, tier=ol.TIER_AVX512)
    g.add_many(rows, label0=1)
    p.add_many(rows, 1)
    rng = np.random.default_rng(vtype * 7 + metric)
    for lab in rng.choice(np.arange(1, n + 1), 200, replace=False).tolist():  # deleted docs stay in the filter: NaN
        g.delete(int(lab))
        p.delete(int(lab))
    q = ol.synth_rows(vtype, 43, 0, 1, dim)[0]
    # the reference's loop hands GetDistanceFrom_Unsafe a query it normalised itself (hybrid_reader.c:296-305)
    vtype_code = {ol.F32: vs.VecSimType_FLOAT32, ol.I8: vs.VecSimType_INT8, ol.F16: vs.VecSimType_FLOAT16}[vtype]
    mcode = {ol.COS: vs.VecSimMetric_Cosine, ol.L2: vs.VecSimMetric_L2, ol.IP: vs.VecSimMetric_IP}[metric]
    qb = np.zeros(g.L.VecSimParams_GetQueryBlobSize(vtype_code, dim, mcode), dtype=np.uint8)
    qb[: q.nbytes] = q.view(np.uint8)
    if metric == ol.COS:
        vs.normalize(qb, dim, vtype_code)
    for m in (3, 500, 6000):
        doc_ids = np.sort(rng.choice(np.arange(1, n + 400), m, replace=False)).astype(np.uint32)  # some ids beyond the index
        labels, scores, rc = g.topk_filtered(q, k, doc_ids)
        assert rc == 0
        exp = _oracle_adhoc(p, qb, doc_ids, k)
        assert labels.tolist() == [d for _, d in exp], (m, labels, exp)
        tol = 0 if vtype in (ol.F32, ol.I8) else 1e-2
        for s, (es, _) in zip(scores, exp):
            assert (np.float32(s).tobytes() == np.float32(es).tobytes()) if tol == 0 else abs(s - es) <= tol * max(1.0, abs(es))
    # empty filter and k larger than the filter
    labels, scores, rc = g.topk_filtered(q, k, np.zeros(0, dtype=np.uint32))
    assert rc == 0 and len(labels) == 0


def test_filter_from_device_intersection_feeds_the_knn():
    """C5 shape: 2-term AND on the device -> its docIds (still on the device) are the filter of the KNN."""
    from redisearch_b200 import postings as ps
    from redisearch_b200 import vecsim as vs

    n, dim, k = 50_000, 64, 10
    rows = ol.synth_rows(ol.F32, 42, 0, n, dim)
    g = vs.VecSimIndex(vs.VecSimType_FLOAT32, dim, vs.VecSimMetric_Cosine)
    p = ol.PortIndex(ol.F32, dim, ol.COS, tier=ol.TIER_AVX512)
    g.add_many(rows, label0=1)
    p.add_many(rows, 1)
    rng = np.random.default_rng(2)
    a = np.unique(rng.integers(1, n + 1, 30_000)).astype(np.uint64)
    b = np.unique(rng.integers(1, n + 1, 20_000)).astype(np.uint64)
    rs = ps.intersect([ps.PostingList.from_arrays(a), ps.PostingList.from_arrays(b)])
    filt = np.intersect1d(a, b)
    assert len(rs) == len(filt)
    q = ol.synth_rows(ol.F32, 43, 0, 1, dim)[0]
    qn = q.copy()
    ol.port().orc_normalize(ol._p(qn), dim, ol.F32)
    d_ptr = ps.lib().II_ResultSet_DeviceDocIds(rs.h)
    labels, scores, rc = g.topk_filtered(q, k, d_ptr, n=len(filt))
    assert rc == 0
    exp = _oracle_adhoc(p, qn, filt, k)
    assert labels.tolist() == [d for _, d in exp]
    assert np.asarray(scores, dtype=np.float32).tobytes() == np.array([s for s, _ in exp], dtype=np.float32).tobytes()


@pytest.mark.gpu
def test_batched_filters_and_batched_filtered_knn_equal_the_single_calls():
    """configs[4] in one call each: II_IntersectBatch (the filters of many queries) + VecSimB200_TopKFilteredBatch (their filtered
    KNN): ids and score bits equal to the per-query calls and to the oracle's ad-hoc loop; empty filters and filters smaller than k
    included."""
    import ctypes as C

    from redisearch_b200 import postings as ps
    from redisearch_b200 import vecsim as vs

    n, dim, k, nq = 60_000, 64, 10, 20
    rows = ol.synth_rows(ol.F32, 42, 0, n, dim)
    g = vs.VecSimIndex(vs.VecSimType_FLOAT32, dim, vs.VecSimMetric_Cosine)
    p = ol.PortIndex(ol.F32, dim, ol.COS, tier=ol.TIER_AVX512)
    g.add_many(rows, label0=1)
    p.add_many(rows, 1)
    rng = np.random.default_rng(12)
    sizes = [30_000, 20_000, 9_000, 400, 60, 5, 0]
    pool = [np.unique(rng.integers(1, n + 1, s)).astype(np.uint64) for s in sizes]
    pls = [ps.PostingList.from_arrays(x) for x in pool]
    pairs = [(int(rng.integers(0, len(pool))), int(rng.integers(0, len(pool)))) for _ in range(nq)]
    pairs[0], pairs[1] = (0, 6), (4, 5)  # an empty filter, a filter smaller than k
    P, L = ps.lib(), vs.lib()
    arrays = [(C.c_void_p * 2)(pls[a].h, pls[b].h) for a, b in pairs]
    lists_pp = (C.c_void_p * nq)(*[C.cast(a, C.c_void_p) for a in arrays])
    n_lists = (C.c_size_t * nq)(*([2] * nq))
    rs_out = (C.c_void_p * nq)()
    built = P.II_IntersectBatch(nq, lists_pp, n_lists, rs_out)
    assert built == nq
    qs = ol.synth_rows(ol.F32, 43, 0, nq, dim)
    q_ptrs = (C.c_void_p * nq)(*[qs[i].ctypes.data for i in range(nq)])
    id_ptrs, counts = (C.c_void_p * nq)(), (C.c_size_t * nq)()
    filters = []
    for i, (a, b) in enumerate(pairs):
        filt = np.intersect1d(pool[a], pool[b])
        filters.append(filt)
        assert P.II_ResultSet_Len(rs_out[i]) == len(filt)
        counts[i] = len(filt)
        id_ptrs[i] = P.II_ResultSet_DeviceDocIds(rs_out[i]) if len(filt) else None
    out_l = np.zeros((nq, k), dtype=np.uint64)
    out_s = np.zeros((nq, k), dtype=np.float64)
    out_c = (C.c_size_t * nq)()
    assert L.VecSimB200_TopKFilteredBatch(g.h, q_ptrs, nq, k, id_ptrs, counts, out_l.ctypes.data, out_s.ctypes.data, out_c) == 0
    for i in range(nq):
        qn = qs[i].copy()
        ol.port().orc_normalize(ol._p(qn), dim, ol.F32)
        exp = _oracle_adhoc(p, qn, filters[i], k)
        assert out_c[i] == len(exp), (i, pairs[i])
        assert out_l[i, :out_c[i]].tolist() == [d for _, d in exp]
        assert out_s[i, :out_c[i]].astype(np.float32).tobytes() == np.array([s for s, _ in exp], dtype=np.float32).tobytes()
        if len(filters[i]):
            labels, scores, rc = g.topk_filtered(qs[i], k, id_ptrs[i], n=len(filters[i]))
            assert rc == 0 and labels.tolist() == out_l[i, :out_c[i]].tolist()
        P.II_ResultSet_Free(rs_out[i])
