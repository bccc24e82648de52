"""GPU parity tests of the FLAT KNN path, through the C-ABI (libvecsim_b200.so) against the oracle.

Structure follows the reference's own unit tests (deps/VectorSimilarity/tests/unit/test_bruteforce.cpp,
test_spaces.cpp, test_fp16/bf16/int8/uint8.cpp): known-answer cases, dim-residual sweeps where every
kernel must equal the baseline, edge cases (empty index, k=0, k>n, inf scores, ties), batch iterator,
range, ad-hoc, swap-delete.  The checker is oracle/liboracle.so (our CPU restatement, itself pinned to
the reference by tests/test_oracle_vecsim.py) and, when present, oracle/_ref (the reference's code).

Bars: fp32 and int8/uint8 bit-exact ids AND scores; fp16/bf16 |d| <= 1e-2 * max(|ref|, scale)
(BASELINE.md §3.4), ids equal modulo candidates within that tolerance of the k-th score.
"""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import BF16, COS, F16, F32, I8, IP, L2, TIER_AVX512, U8

pytestmark = pytest.mark.gpu

VS_TYPE = {F32: 0, BF16: 2, F16: 3, I8: 4, U8: 5}  # identical numbering by construction
_KEEPALIVE = []


@pytest.fixture(scope="module")
def vs():
    from redisearch_b200 import vecsim

    return vecsim


def make_pair(vs, vtype, dim, metric, blobs, label0=1, multi=False):
    g = vs.VecSimIndex(vtype, dim, metric, multi=multi)
    p = ol.PortIndex(vtype, dim, metric, multi=multi, tier=TIER_AVX512)
    if len(blobs):
        assert g.add_many(blobs, label0=label0) == len(blobs)
        p.add_many(blobs, label0)
    return g, p


def rand_blobs(rng, vtype, n, dim):
    return ol.to_type(rng.uniform(-1, 1, (n, dim)).astype(np.float32), vtype)


def assert_same(gi, gs, pi, ps, exact, metric, kth_tol=1e-2):
    if exact:
        assert gi.tolist() == pi.tolist()
        assert gs.astype(np.float32).tobytes() == ps.astype(np.float32).tobytes()
        return
    assert len(gi) == len(pi)
    scale = 1.0 if metric != L2 else 0.0
    # scores, position by position, within tolerance
    for a, b in zip(gs, ps):
        assert abs(a - b) <= kth_tol * max(abs(b), scale) + 1e-6, (a, b)
    # ids equal modulo candidates within tolerance of the k-th score
    if len(pi):
        kth = ps[-1]
        slack = kth_tol * max(abs(kth), scale) + 1e-6
        sure = {i for i, s in zip(pi.tolist(), ps.tolist()) if s < kth - slack}
        assert sure <= set(gi.tolist())


# ------------------------------------------------------------------------------------------------
# known answers transcribed from the reference's unit tests
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vtype", [F32, F16, BF16])
@pytest.mark.parametrize("block_size", [1, 12, 1024])
def test_bruteforce_vector_search_l2_known_answers(vs, vtype, block_size):
    """test_bruteforce.cpp:781-812"""
    dim, n, k = 4, 100, 11
    g = vs.VecSimIndex(vtype, dim, L2, block_size=block_size)
    assert g.basic_info().blockSize == block_size and g.basic_info().algo == vs.VecSimAlgo_BF
    for i in range(n):
        assert g.add(ol.to_type(np.full(dim, float(i), dtype=np.float32), vtype), i) == 1
    assert g.size() == n
    q = ol.to_type(np.full(dim, 50.0, dtype=np.float32), vtype)
    ids, scores, code = g.topk(q, k)
    assert code == vs.VecSim_QueryReply_OK and len(ids) == k
    for idx, (i, s) in enumerate(zip(ids.tolist(), scores.tolist())):
        assert abs(i - 50) == (idx + 1) // 2
        assert s == 4 * ((idx + 1) // 2) ** 2
    assert len(g.topk(q, 0)[0]) == 0  # "search for nothing"


def test_bruteforce_vector_search_ip_known_answers(vs):
    """test_bruteforce.cpp:747-779"""
    dim, n, k = 4, 100, 11
    g = vs.VecSimIndex(F32, dim, IP)
    for i in range(n):
        g.add(np.full(dim, float(i), dtype=np.float32), i)
    ids, _, _ = g.topk(np.full(dim, 50.0, dtype=np.float32), k)
    assert set(ids.tolist()) == set(range(n - k, n))


def test_search_empty_index(vs):
    """test_bruteforce.cpp:814-862"""
    dim, n, k = 4, 100, 11
    g = vs.VecSimIndex(F32, dim, L2)
    q = np.full(dim, 50.0, dtype=np.float32)
    assert g.size() == 0
    assert len(g.topk(q, k)[0]) == 0
    assert len(g.range(q, 1.0)[0]) == 0
    for i in range(n):
        g.add(np.full(dim, float(i), dtype=np.float32), i)
    assert g.size() == n
    for i in range(n):
        assert g.delete(i) == 1
    assert g.size() == 0
    assert len(g.topk(q, k)[0]) == 0
    assert len(g.range(q, 1.0)[0]) == 0
    rep = g.L.VecSimIndex_TopKQuery(g.h, q.ctypes.data_as(C.c_void_p), k, None, 0)
    it = g.L.VecSimQueryReply_GetIterator(rep)
    assert not g.L.VecSimQueryReply_IteratorNext(it)
    assert g.L.VecSimQueryResult_GetId(None) == 0xFFFFFFFF and np.isnan(g.L.VecSimQueryResult_GetScore(None))
    g.L.VecSimQueryReply_IteratorFree(it)
    g.L.VecSimQueryReply_Free(rep)


def test_inf_score(vs):
    """test_bruteforce.cpp:864-905: +inf distances are valid results and sort last."""
    dim = 2
    g = vs.VecSimIndex(F32, dim, L2)
    inf_v = np.array([np.finfo(np.float32).max, np.finfo(np.float32).max], dtype=np.float32)
    g.add(np.array([1, 1], dtype=np.float32), 1)
    g.add(np.array([1, 1], dtype=np.float32), 2)
    g.add(inf_v, 3)
    g.add(np.array([1, 2], dtype=np.float32), 4)
    ids, scores, _ = g.topk(np.array([1, 1], dtype=np.float32), 4)
    assert ids.tolist()[:2] == [1, 2] and scores[0] == 0 and scores[1] == 0
    assert ids.tolist()[2] == 4 and scores[2] == 1
    assert ids.tolist()[3] == 3 and np.isinf(scores[3])


# ------------------------------------------------------------------------------------------------
# every kernel == baseline over dimension residuals (test_spaces.cpp:683-736, INSTANTIATE :877)
# ------------------------------------------------------------------------------------------------
DIM_SWEEP = [1, 3, 7, 8, 9, 15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 96, 100, 127, 128, 129, 255, 256, 257, 771]


@pytest.mark.parametrize("metric", [L2, IP, COS])
def test_fp32_bit_exact_over_dim_residuals(vs, metric):
    rng = np.random.default_rng(100 + metric)
    ref = ol.ref_vecsim()
    for dim in DIM_SWEEP:
        n, k = 400, 10
        blobs = rand_blobs(rng, F32, n, dim)
        g, p = make_pair(vs, F32, dim, metric, blobs)
        r = None
        if ref is not None and ol.host_has_avx512f():
            r = ol.RefIndex(F32, dim, metric)
            r.add_many(blobs, 1)
        for _ in range(3):
            q = rand_blobs(rng, F32, 1, dim)[0]
            gi, gs, _ = g.topk(q, k)
            pi, ps = p.topk(q, k)
            assert_same(gi, gs, pi, ps, True, metric)
            if r is not None:  # and against the reference's own code, bit for bit
                ri, rs = r.topk(q, k)
                assert_same(gi, gs, ri, rs, True, metric)


@pytest.mark.parametrize("vtype", [I8, U8])
@pytest.mark.parametrize("metric", [L2, IP, COS])
def test_int8_bit_exact_over_dim_residuals(vs, vtype, metric):
    rng = np.random.default_rng(200 + vtype * 3 + metric)
    for dim in DIM_SWEEP:
        n, k = 300, 10
        blobs = rand_blobs(rng, vtype, n, dim)
        g, p = make_pair(vs, vtype, dim, metric, blobs)
        for _ in range(2):
            q = rand_blobs(rng, vtype, 1, dim)[0]
            gi, gs, _ = g.topk(q, k)
            pi, ps = p.topk(q, k)
            # integer scores tie often: compare score lists exactly and ids as (score,label)-sorted sets
            assert gs.astype(np.float32).tobytes() == ps.astype(np.float32).tobytes(), dim
            kth = ps[-1]
            assert {i for i, s in zip(gi.tolist(), gs.tolist()) if s < kth} == {i for i, s in zip(pi.tolist(), ps.tolist()) if s < kth}


@pytest.mark.parametrize("vtype", [F16, BF16])
@pytest.mark.parametrize("metric", [L2, IP, COS])
def test_half_types_within_tolerance_over_dim_residuals(vs, vtype, metric):
    rng = np.random.default_rng(300 + vtype * 3 + metric)
    for dim in DIM_SWEEP:
        n, k = 300, 10
        blobs = rand_blobs(rng, vtype, n, dim)
        g, p = make_pair(vs, vtype, dim, metric, blobs)
        q = rand_blobs(rng, vtype, 1, dim)[0]
        gi, gs, _ = g.topk(q, k)
        pi, ps = p.topk(q, k)
        assert_same(gi, gs, pi, ps, False, metric)
        # in practice fp32 accumulation of exact products agrees far tighter than the 1e-2 bar
        np.testing.assert_allclose(gs, ps, rtol=2e-4, atol=2e-5)


# ------------------------------------------------------------------------------------------------
# sizes and k edge cases
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("metric", [L2, COS])
def test_config1_shape_100k_x128(vs, metric):
    """BASELINE.json configs[0]: FLAT 100K x 128 fp32, k=10, single query."""
    n, dim, k = 100_000, 128, 10
    blobs = ol.synth_rows(F32, 42, 0, n, dim)
    g, p = make_pair(vs, F32, dim, metric, blobs)
    qs = ol.synth_rows(F32, 43, 0, 8, dim)
    for q in qs:
        gi, gs, _ = g.topk(q, k)
        pi, ps = p.topk(q, k)
        assert_same(gi, gs, pi, ps, True, metric)
        gi, gs, _ = g.topk(q, k, order=vs.BY_ID)
        pi, ps = p.topk(q, k, order=1)
        assert_same(gi, gs, pi, ps, True, metric)


@pytest.mark.parametrize("k", [1, 2, 31, 32, 33, 100, 128, 129, 300, 1000])
def test_k_sweep_including_unfused_path(vs, k):
    rng = np.random.default_rng(k)
    n, dim = 5000, 64
    blobs = rand_blobs(rng, F32, n, dim)
    g, p = make_pair(vs, F32, dim, L2, blobs)
    q = rand_blobs(rng, F32, 1, dim)[0]
    gi, gs, _ = g.topk(q, k)
    pi, ps = p.topk(q, k)
    assert_same(gi, gs, pi, ps, True, L2)


def test_k_larger_than_index(vs):
    rng = np.random.default_rng(5)
    for n in [1, 2, 7, 33, 150]:
        blobs = rand_blobs(rng, F32, n, 24)
        g, p = make_pair(vs, F32, 24, IP, blobs)
        q = rand_blobs(rng, F32, 1, 24)[0]
        for k in [n, n + 1, 200]:
            gi, gs, _ = g.topk(q, k)
            pi, ps = p.topk(q, k)
            assert len(gi) == n
            assert_same(gi, gs, pi, ps, True, IP)


def test_exact_ties_resolve_by_score_then_label(vs):
    """Tie groups are sets in the reference's tests (test_bruteforce.cpp:1003-1055); ours are
    deterministic: (score asc, label asc) like the reference heap's drain order."""
    dim = 8
    g = vs.VecSimIndex(F32, dim, L2)
    v = np.ones(dim, dtype=np.float32)
    for lab in [50, 10, 40, 20, 30]:
        g.add(v, lab)
    g.add(v * 2, 5)
    ids, scores, _ = g.topk(v, 4)
    assert scores.tolist() == [0, 0, 0, 0]
    assert ids.tolist() == sorted(ids.tolist())
    ids, scores, _ = g.topk(v, 6)
    assert ids.tolist() == [10, 20, 30, 40, 50, 5]


# ------------------------------------------------------------------------------------------------
# batched entry point == per-query entry point
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("vtype,metric", [(F32, COS), (F32, L2), (F16, IP), (BF16, COS), (I8, COS), (U8, L2)])
@pytest.mark.parametrize("nq", [1, 2, 8, 9, 33, 70])
def test_batched_queries_match_oracle(vs, vtype, metric, nq):
    rng = np.random.default_rng(nq * 17 + vtype + metric)
    n, dim, k = 3000, 96, 10
    blobs = rand_blobs(rng, vtype, n, dim)
    g, p = make_pair(vs, vtype, dim, metric, blobs)
    qs = rand_blobs(rng, vtype, nq, dim)
    labels, scores, rc = g.topk_batch(qs, k)
    assert rc == 0
    exact = vtype in (F32, I8, U8)
    for i in range(nq):
        pi, ps = p.topk(qs[i], k)
        if vtype in (I8, U8):
            assert scores[i].astype(np.float32).tobytes() == ps.astype(np.float32).tobytes()
        else:
            assert_same(labels[i].astype(np.int64), scores[i], pi, ps, exact, metric)


# ------------------------------------------------------------------------------------------------
# range query, batch iterator, ad-hoc distances
# ------------------------------------------------------------------------------------------------
def test_range_query(vs):
    """test_bruteforce.cpp:1567-1657 shape: all vectors within radius, BY_SCORE and BY_ID."""
    rng = np.random.default_rng(9)
    n, dim = 4000, 32
    blobs = rand_blobs(rng, F32, n, dim)
    g, p = make_pair(vs, F32, dim, L2, blobs)
    q = rand_blobs(rng, F32, 1, dim)[0]
    for radius in [0.0, 5.0, 8.0, 11.0, 1e9]:
        for order in (vs.BY_SCORE, vs.BY_ID):
            gi, gs, code = g.range(q, radius, order)
            pi, ps = p.range(q, radius, order)
            assert code == 0
            assert_same(gi, gs, pi, ps, True, L2)
    with pytest.raises(ValueError):
        g.range(q, -1.0)
    with pytest.raises(ValueError):
        g.range(q, 1.0, order=vs.BY_SCORE_THEN_ID)
    assert g.debug_info()["LAST_SEARCH_MODE"] == "RANGE_QUERY"


@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("batch", [1, 10, 128, 500])
def test_batch_iterator(vs, order, batch):
    """test_bruteforce.cpp:959-1165: successive batches are the next-best results, no repeats,
    iterator depletes exactly at the label count."""
    rng = np.random.default_rng(batch + order)
    n, dim = 2500, 48
    blobs = rand_blobs(rng, F32, n, dim)
    g, p = make_pair(vs, F32, dim, L2, blobs)
    q = rand_blobs(rng, F32, 1, dim)[0]
    all_ids, all_scores = p.all_sorted(q)
    it = g.batch_iterator(q)
    pos = 0
    rounds = 0
    while it.has_next() and rounds < 12:
        ids, scores, code = it.next(batch, order)
        assert code == 0
        exp_i, exp_s = all_ids[pos:pos + batch], all_scores[pos:pos + batch]
        if order == 1:
            o = np.argsort(exp_i, kind="stable")
            exp_i, exp_s = exp_i[o], exp_s[o]
        assert ids.tolist() == exp_i.tolist()
        assert scores.astype(np.float32).tobytes() == exp_s.astype(np.float32).tobytes()
        pos += len(ids)
        rounds += 1
    it.reset()
    ids, _, _ = it.next(5, 0)
    assert ids.tolist() == all_ids[:5].tolist()
    it.free()
    # drain a small index completely
    g2, p2 = make_pair(vs, F32, dim, L2, blobs[:37])
    it = g2.batch_iterator(q)
    got = []
    while it.has_next():
        ids, _, _ = it.next(10, 0)
        got += ids.tolist()
    assert got == p2.all_sorted(q)[0].tolist()
    assert len(it.next(10, 0)[0]) == 0


@pytest.mark.parametrize("vtype,metric", [(F32, COS), (I8, COS), (F16, L2)])
def test_adhoc_distances(vs, vtype, metric):
    """VecSimIndex_GetDistanceFrom_Unsafe (brute_force_single.h:200-212) and the batched ad-hoc ctx."""
    rng = np.random.default_rng(12 + vtype)
    n, dim = 1500, 40
    blobs = rand_blobs(rng, vtype, n, dim)
    g, p = make_pair(vs, vtype, dim, metric, blobs)
    q = rand_blobs(rng, vtype, 1, dim)[0]
    # the _Unsafe call expects a query the CALLER already normalised (hybrid_reader.c:296-305)
    qb = np.zeros(g.L.VecSimParams_GetQueryBlobSize(vtype, dim, metric), dtype=np.uint8)
    qb[: q.nbytes] = q.view(np.uint8)
    if metric == COS:
        vs.normalize(qb, dim, vtype)
    for lab in [1, 2, 700, n, n + 5]:
        a = g.distance_from(lab, qb)
        b = p.distance_from(lab, qb)
        if lab > n:
            assert np.isnan(a) and np.isnan(b)
        elif vtype == F16:
            assert abs(a - b) <= 1e-4 * max(abs(b), 1e-3)
        else:
            assert np.float32(a).tobytes() == np.float32(b).tobytes()
    labels = np.array([5, 9999999, 17, 1, n, 3], dtype=np.uint64)
    d = g.adhoc_distances(q, labels)  # the ctx normalises internally
    for lab, a in zip(labels.tolist(), d.tolist()):
        b = p.distance_from(lab, qb)
        assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= 1e-4 * max(abs(b), 1e-3)


def test_prefer_adhoc_and_modes(vs):
    """brute_force.h:380-451 thresholds + lastMode bookkeeping."""
    for dim, n in [(4, 100), (4, 6000), (128, 6000), (768, 6000)]:
        g = vs.VecSimIndex(F32, dim, L2)
        p = ol.PortIndex(F32, dim, L2)
        x = np.zeros((n, dim), dtype=np.float32)
        g.add_many(x, label0=0)
        p.add_many(x, 0)
        for frac in [0.0, 0.1, 0.15, 0.16, 0.35, 0.36, 0.55, 0.56, 0.75, 0.76, 1.0, 3.0]:
            s = int(frac * n)
            assert g.prefer_adhoc(s, 10, True) == p.prefer_adhoc(s, 10, True), (dim, n, frac)
        res = g.prefer_adhoc(0, 10, True)
        assert g.debug_info()["LAST_SEARCH_MODE"] == ("HYBRID_ADHOC_BF" if res else "HYBRID_BATCHES")
        if g.prefer_adhoc(0, 10, False):
            assert g.debug_info()["LAST_SEARCH_MODE"] == "HYBRID_BATCHES_TO_ADHOC_BF"


# ------------------------------------------------------------------------------------------------
# mutation: swap-delete, update in place, re-add
# ------------------------------------------------------------------------------------------------
def test_swap_delete_update_readd(vs):
    rng = np.random.default_rng(21)
    n, dim = 3000, 20
    blobs = rand_blobs(rng, F32, n, dim)
    g, p = make_pair(vs, F32, dim, COS, blobs, label0=0)
    for lab in rng.choice(n, 700, replace=False).tolist():
        assert g.delete(lab) == 1 and p.delete(lab) == 1
    assert g.delete(10**9) == 0
    for lab in [3, 4, 5, 6, 7, 8]:  # overwrite (or re-add if deleted above)
        v = rand_blobs(rng, F32, 1, dim)[0]
        assert g.add(v, lab) == p.add(v, lab)
    extra = rand_blobs(rng, F32, 500, dim)
    g.add_many(extra, label0=10_000)
    p.add_many(extra, 10_000)
    assert g.size() == p.size()
    for _ in range(4):
        q = rand_blobs(rng, F32, 1, dim)[0]
        gi, gs, _ = g.topk(q, 25)
        pi, ps = p.topk(q, 25)
        assert_same(gi, gs, pi, ps, True, COS)
    info = g.debug_info()
    assert info["INDEX_SIZE"] == g.size() and info["ALGORITHM"] == "FLAT" and info["METRIC"] == "COSINE"
    assert g.stats_info().memory > 0
    # the by-value struct of VecSimIndex_DebugInfo (BruteForceIndex::debugInfo, brute_force.h:318-325) says the same
    di = vs.lib().VecSimIndex_DebugInfo(g.h)
    assert C.sizeof(di) == 360 and di.commonInfo.indexSize == g.size() and di.commonInfo.indexLabelCount == info["INDEX_LABEL_COUNT"]
    assert di.commonInfo.basicInfo.metric == COS and di.commonInfo.basicInfo.dim == dim and di.commonInfo.memory == info["MEMORY"]
    assert di.commonInfo.lastMode == vs.STANDARD_KNN
    vs.lib().VecSim_SetTestLogContext(b"parity", b"unit")


def test_multi_value_index(vs):
    """brute_force_multi.h: several vectors per label, best score per label, delete removes all."""
    rng = np.random.default_rng(33)
    dim, n_labels, per = 16, 400, 3
    g = vs.VecSimIndex(F32, dim, L2, multi=True)
    p = ol.PortIndex(F32, dim, L2, multi=True)
    for lab in range(n_labels):
        for _ in range(per):
            v = rand_blobs(rng, F32, 1, dim)[0]
            assert g.add(v, lab) == 1 and p.add(v, lab) == 1
    assert g.size() == n_labels * per
    assert g.debug_info()["INDEX_LABEL_COUNT"] == n_labels
    q = rand_blobs(rng, F32, 1, dim)[0]
    for k in [1, 10, 150, 500]:
        gi, gs, _ = g.topk(q, k)
        pi, ps = p.topk(q, k)
        assert len(set(gi.tolist())) == len(gi)
        assert_same(gi, gs, pi, ps, True, L2)
    assert g.delete(7) == per and p.delete(7) == per
    gi, gs, _ = g.topk(q, 20)
    pi, ps = p.topk(q, 20)
    assert_same(gi, gs, pi, ps, True, L2)
    a, b = g.distance_from(9, q), p.distance_from(9, q)
    assert np.float32(a).tobytes() == np.float32(b).tobytes()
    it = g.batch_iterator(q)
    got = []
    while it.has_next():
        ids, _, _ = it.next(64, 0)
        if not len(ids):
            break
        got += ids.tolist()
    assert got == p.all_sorted(q)[0].tolist()


# ------------------------------------------------------------------------------------------------
# timeouts and parameter resolution
# ------------------------------------------------------------------------------------------------
def test_timeout_callback(vs):
    """test_bruteforce.cpp:1489-1565: a firing timeout callback yields VecSim_QueryReply_TimedOut."""
    L = vs.lib()
    g = vs.VecSimIndex(F32, 8, L2)
    g.add_many(np.zeros((100, 8), dtype=np.float32), label0=0)
    cb = vs.TIMEOUT_CB(lambda ctx: 1)
    cb_off = vs.TIMEOUT_CB(lambda ctx: 0)
    _KEEPALIVE.extend([cb, cb_off])  # ctypes trampolines must outlive their registration
    L.VecSim_SetTimeoutCallbackFunction(cb)
    try:
        q = np.zeros(8, dtype=np.float32)
        assert g.topk(q, 5)[2] == vs.VecSim_QueryReply_TimedOut
        assert g.range(q, 1.0)[2] == vs.VecSim_QueryReply_TimedOut
        it = g.batch_iterator(q)
        assert it.next(5)[2] == vs.VecSim_QueryReply_TimedOut
        it.free()
    finally:
        L.VecSim_SetTimeoutCallbackFunction(cb_off)
    assert g.topk(np.zeros(8, dtype=np.float32), 5)[2] == vs.VecSim_QueryReply_OK


def test_timeout_fires_while_the_scan_is_running(vs):
    """The reference polls the timeout per vector (brute_force.h:265-269).  A device pass cannot be interrupted, but the host
    polls the callback while kernels run: a deadline that passes DURING a long exact scan releases the caller at once with
    VecSim_QueryReply_TimedOut, and the abandoned scratch is drained before it is reused (the next query is correct)."""
    import time

    L = vs.lib()
    L.VecSimB200_SetCoarseMode(0)  # force the slow exact batched scan
    n, dim, nq, k = 400_000, 256, 256, 10
    rows = ol.synth_rows(ol.F32, 5, 0, n, dim)
    g = vs.VecSimIndex(F32, dim, L2)
    g.add_many(rows, label0=1)
    qs = ol.synth_rows(ol.F32, 6, 0, nq, dim)
    labels, scores, rc = g.topk_batch(qs, k)  # warm: allocations, first launch
    t0 = time.perf_counter()
    labels, scores, rc = g.topk_batch(qs, k)
    full_s = time.perf_counter() - t0
    assert rc == vs.VecSim_QueryReply_OK
    calls = {"n": 0}

    def fire_after_a_few_polls(ctx):
        calls["n"] += 1
        return 1 if calls["n"] > 4 else 0  # the pre-launch check and three polls pass, then the deadline is over

    cb = vs.TIMEOUT_CB(fire_after_a_few_polls)
    cb_off = vs.TIMEOUT_CB(lambda ctx: 0)
    _KEEPALIVE.extend([cb, cb_off])
    qp = vs.VecSimQueryParams()
    L.VecSim_SetTimeoutCallbackFunction(cb)
    try:
        out_l = np.zeros((nq, k), dtype=np.uint64)
        out_s = np.zeros((nq, k), dtype=np.float64)
        t0 = time.perf_counter()
        rc = L.VecSimB200_TopKQueryBatch(g.h, qs.ctypes.data, qs.strides[0], nq, k, C.byref(qp), out_l.ctypes.data, out_s.ctypes.data)
        early_s = time.perf_counter() - t0
    finally:
        L.VecSim_SetTimeoutCallbackFunction(cb_off)
    assert rc == vs.VecSim_QueryReply_TimedOut
    if full_s > 0.004:  # only meaningful when the scan is long compared with the polling period
        assert early_s < 0.6 * full_s, (early_s, full_s)
    l2, s2, rc2 = g.topk_batch(qs, k)
    assert rc2 == vs.VecSim_QueryReply_OK and (l2 == labels).all() and s2.tobytes() == scores.tobytes()
    L.VecSimB200_SetCoarseMode(-1)


def test_abandoned_scan_winds_down_on_the_device(vs):
    """A caller that timed out has left; the exact-scan kernel polls a host flag (mapped pinned memory) and stops, so the GPU does
    not finish a long pass nobody waits for: after the timed-out call the device drains in a fraction of the pass."""
    import time

    import torch

    L = vs.lib()
    L.VecSimB200_SetCoarseMode(0)
    try:
        n, dim, nq, k = 2_000_000, 256, 256, 10
        g = vs.VecSimIndex(F32, dim, L2)
        for r0 in range(0, n, 500_000):
            g.add_many(ol.synth_rows(ol.F32, 5, r0, 500_000, dim), label0=1 + r0)
        qs = ol.synth_rows(ol.F32, 6, 0, nq, dim)
        g.topk_batch(qs, k)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        labels, scores, rc = g.topk_batch(qs, k)
        full_s = time.perf_counter() - t0
        assert rc == vs.VecSim_QueryReply_OK and full_s > 0.01, full_s
        calls = {"n": 0}

        def fire_after_a_few_polls(ctx):
            calls["n"] += 1
            return 1 if calls["n"] > 4 else 0

        cb = vs.TIMEOUT_CB(fire_after_a_few_polls)
        cb_off = vs.TIMEOUT_CB(lambda ctx: 0)
        _KEEPALIVE.extend([cb, cb_off])
        qp = vs.VecSimQueryParams()
        out_l = np.zeros((nq, k), dtype=np.uint64)
        out_s = np.zeros((nq, k), dtype=np.float64)
        L.VecSim_SetTimeoutCallbackFunction(cb)
        try:
            t0 = time.perf_counter()
            rc = L.VecSimB200_TopKQueryBatch(g.h, qs.ctypes.data, qs.strides[0], nq, k, C.byref(qp), out_l.ctypes.data, out_s.ctypes.data)
            early_s = time.perf_counter() - t0
            torch.cuda.synchronize()
            drained_s = time.perf_counter() - t0
        finally:
            L.VecSim_SetTimeoutCallbackFunction(cb_off)
        assert rc == vs.VecSim_QueryReply_TimedOut
        assert early_s < 0.5 * full_s and drained_s < 0.5 * full_s, (early_s, drained_s, full_s)
        l2, s2, rc2 = g.topk_batch(qs, k)  # the flag is cleared when the scratch is reused
        assert rc2 == vs.VecSim_QueryReply_OK and (l2 == labels).all() and s2.tobytes() == scores.tobytes()
    finally:
        L.VecSimB200_SetCoarseMode(-1)


def test_resolve_params(vs):
    """vec_sim.cpp:270-343 for a FLAT index."""
    L = vs.lib()
    g = vs.VecSimIndex(F32, 8, L2)
    qp = vs.VecSimQueryParams()

    def resolve(pairs, qtype):
        arr = (vs.VecSimRawParam * max(1, len(pairs)))()
        for i, (n, v) in enumerate(pairs):
            arr[i] = vs.VecSimRawParam(n.encode(), len(n), v.encode(), len(v))
        return L.VecSimIndex_ResolveParams(g.h, arr, len(pairs), C.byref(qp), qtype)

    assert resolve([], vs.QUERY_TYPE_KNN) == 0
    assert resolve([("BATCH_SIZE", "100")], vs.QUERY_TYPE_HYBRID) == 0 and qp.batchSize == 100
    assert resolve([("batch_size", "100")], vs.QUERY_TYPE_KNN) == 6  # NHybrid
    assert resolve([("BATCH_SIZE", "0")], vs.QUERY_TYPE_HYBRID) == 4  # BadValue
    assert resolve([("BATCH_SIZE", "7"), ("BATCH_SIZE", "8")], vs.QUERY_TYPE_HYBRID) == 2  # AlreadySet
    assert resolve([("HYBRID_POLICY", "batches")], vs.QUERY_TYPE_HYBRID) == 0 and qp.searchMode == vs.HYBRID_BATCHES
    assert resolve([("HYBRID_POLICY", "ADHOC_BF")], vs.QUERY_TYPE_HYBRID) == 0 and qp.searchMode == vs.HYBRID_ADHOC_BF
    assert resolve([("HYBRID_POLICY", "nope")], vs.QUERY_TYPE_HYBRID) == 5  # NExits
    assert resolve([("HYBRID_POLICY", "adhoc_bf"), ("BATCH_SIZE", "5")], vs.QUERY_TYPE_HYBRID) == 8
    assert resolve([("EF_RUNTIME", "10")], vs.QUERY_TYPE_KNN) == 3  # UnknownParam for FLAT
    assert resolve([("EPSILON", "0.1")], vs.QUERY_TYPE_RANGE) == 3
    assert resolve([("WHATEVER", "1")], vs.QUERY_TYPE_KNN) == 3
    assert L.VecSimIndex_ResolveParams(g.h, None, 1, C.byref(qp), 1) == 1  # NullParam
