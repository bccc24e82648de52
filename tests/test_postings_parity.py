"""GPU parity tests of the posting-list path through the C-ABI (libii_b200.so) against the oracle.

Bars (BASELINE.json north_star): docID sequences of intersection / union bit-exact and in the same
order as the reference iterators; BM25 / TF-IDF scores within 1e-5 relative — in practice the
kernels reproduce the reference expression trees and are checked for BIT equality against
oracle/scorer_oracle.c (itself bit-equal to the reference's default.c).

Shapes follow the reference's tests: rqe_iterators/tests/integration/intersection.rs
(NUM_CHILDREN x RESULT_SET cases, read / skip_to / rewind), union_common.rs, the codec golden tests
and tests/cpptests/test_cpp_index.cpp:542-601.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import oracle_lib as ol

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "postings_golden.json")))


@pytest.fixture(scope="module")
def ps():
    from redisearch_b200 import postings

    return postings


def make_lists(ps, codec, id_lists, freq_lists=None, on_device=False):
    idx, pls = [], []
    for i, ids in enumerate(id_lists):
        fr = freq_lists[i] if freq_lists is not None else [1] * len(ids)
        ix = ol.InvIndex(codec, ids, fr, [1] * len(ids))
        idx.append(ix)
        pls.append(ps.PostingList.from_blocks(ix.blocks(), codec, on_device=on_device))
    return idx, pls


# ------------------------------------------------------------------------------------------------
# block decoding: every codec, host and device decoders, vs the oracle reader
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("codec", range(ol.N_CODECS))
@pytest.mark.parametrize("on_device", [False, True])
def test_decode_blocks_all_codecs(ps, codec, on_device):
    rng = np.random.default_rng(codec * 2 + on_device)
    ids = np.cumsum(rng.integers(1, 5000, 5321)).astype(np.uint64)
    freqs = rng.integers(1, 70000, len(ids))
    masks = rng.integers(1, 1 << 30, len(ids)).tolist()
    if codec in ol.CODECS_WIDE:
        masks = [m << int(s_) for m, s_ in zip(masks, rng.integers(0, 98, len(ids)))]
    offs = [bytes(rng.integers(0, 255, int(rng.integers(0, 6))).astype(np.uint8)) for _ in ids]
    ix = ol.InvIndex(codec)
    for d, f, m, o in zip(ids.tolist(), freqs.tolist(), masks, offs):
        ix.add(d, f, m, o if codec in ol.CODECS_WITH_OFFSETS else b"")
    pl = ps.PostingList.from_blocks(ix.blocks(), codec, on_device=on_device)
    assert len(pl) == len(ids) == pl.num_estimated()
    rs = ps.union([pl])
    got_ids, _, got_fr = rs.fetch()
    exp = ix.read_all()
    assert got_ids.tolist() == [e[0] for e in exp]
    assert got_fr[0].tolist() == [e[1] for e in exp]


@pytest.mark.parametrize("codec", ol.CODECS_WITH_MASK)
@pytest.mark.parametrize("on_device", [False, True])
def test_field_mask_filter(ps, codec, on_device):
    """FilterMaskReader: records whose fieldMask misses the query mask are dropped; estimate unchanged.  The *Wide codecs carry
    u128 masks (more than 32 fields) and take a 128-bit filter."""
    rng = np.random.default_rng(3)
    ids = np.cumsum(rng.integers(1, 9, 3000)).astype(np.uint64)
    masks = rng.integers(1, 16, len(ids)).tolist()
    filters = (1, 6, 8)
    if codec in ol.CODECS_WIDE:
        masks = [m << int(s_) for m, s_ in zip(masks, rng.choice([0, 30, 62, 100], len(ids)))]
        filters = (1, 6, 1 << 33, (1 << 64) | (1 << 101), (1 << 3) | (1 << 65))
    ix = ol.InvIndex(codec)
    for d, m in zip(ids.tolist(), masks):
        ix.add(d, 2, m, b"\5" if codec in ol.CODECS_WITH_OFFSETS else b"")
    for flt in filters:
        pl = ps.PostingList.from_blocks(ix.blocks(), codec, field_mask_filter=flt, on_device=on_device)
        exp = ix.read_all(flt)
        assert len(pl) == len(exp) and pl.num_estimated() == len(ids), (codec, flt)
        if len(exp):
            got, _, _ = ps.union([pl]).fetch()
            assert got.tolist() == [e[0] for e in exp]
    if codec not in ol.CODECS_WIDE:  # bits above 31 cannot be met by a 32-bit mask codec: refused, not silently truncated
        with pytest.raises(RuntimeError):
            ps.PostingList.from_blocks(ix.blocks(), codec, field_mask_filter=1 << 40)


def test_golden_codec_bytes_decode(ps):
    """The reference's golden byte vectors (codec/freqs_only.rs:26-49) decode to the values they encode."""
    base = 1 << 20
    for freq, delta, data in G["freqs_only"]:
        if delta > 0xFFFF0000:
            continue  # docIds must stay below 2^32 on the device
        blocks = [(base, base + delta, 1, bytes(data))]
        pl = ps.PostingList.from_blocks(blocks, ps.CODEC_FREQS_ONLY)
        ids, _, fr = ps.union([pl]).fetch()
        assert ids.tolist() == [base + delta] and fr[0].tolist() == [freq if freq else 0] or freq == 0


# ------------------------------------------------------------------------------------------------
# intersection
# ------------------------------------------------------------------------------------------------
def _children_for(result_set, num_children):
    nxt, out = 1, []
    for _ in range(num_children):
        ids = set(result_set)
        for _ in range(100):
            ids.add(nxt)
            nxt += 1
        out.append(sorted(ids))
    return out


@pytest.mark.parametrize("num_children", [2, 5, 16])
@pytest.mark.parametrize("case", range(3))
def test_intersection_reference_cases(ps, num_children, case):
    """intersection.rs:59-148 read_all_combinations (25 children exceed the device's 16-list limit)."""
    children = _children_for(G["intersection_result_sets"][case], num_children)
    idx, pls = make_lists(ps, ps.CODEC_FREQS_ONLY, children)
    rs = ps.intersect(pls)
    exp = ol.run_intersect(idx)
    ids, _, fr = rs.fetch()
    assert ids.tolist() == [e[0] for e in exp]
    assert rs.child_order().tolist() == [c for c, _ in exp[0][1]]
    assert (fr == 1).all()


def test_cpp_intersection_known_answer_and_iterator_contract(ps):
    """test_cpp_index.cpp:542-601 through the QueryIterator facade: 50000 hits, docId (count*2+2)*2,
    freq 2; Rewind; SkipTo(8)=OK; Read -> 12; SkipTo(200000)=OK; Read=EOF (atEOF set, current NULL)."""
    g = G["cpp_intersection"]
    a = np.arange(1, g["size"] + 1) * g["steps"][0]
    b = np.arange(1, g["size"] + 1) * g["steps"][1]
    _, pls = make_lists(ps, ps.CODEC_FULL, [a, b])
    it = ps.intersect(pls).into_iterator()
    q = it.contents
    assert q.lastDocId == 0 and not q.atEOF and not q.current
    assert q.NumEstimated(it) == g["hits"]
    count = 0
    while q.Read(it) != ps.ITERATOR_EOF:
        assert q.lastDocId == (count * 2 + 2) * 2
        assert q.current.contents.docId == q.lastDocId and q.current.contents.freq == g["freq"]
        count += 1
    assert count == g["hits"] and q.atEOF and not q.current
    assert q.Read(it) == ps.ITERATOR_EOF
    q.Rewind(it)
    assert q.lastDocId == 0 and not q.atEOF
    assert q.SkipTo(it, 8) == ps.ITERATOR_OK and q.lastDocId == 8
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 12
    assert q.SkipTo(it, 13) == ps.ITERATOR_NOTFOUND and q.lastDocId == 16
    assert q.SkipTo(it, 200000) == ps.ITERATOR_OK and q.lastDocId == 200000
    assert q.Read(it) == ps.ITERATOR_EOF
    q.Rewind(it)
    assert q.SkipTo(it, 200001) == ps.ITERATOR_EOF and q.atEOF and not q.current
    assert q.Revalidate(it, None) == 0
    q.Free(it)


@pytest.mark.parametrize("sizes", [(50, 40000), (30000, 31000), (5, 7, 9), (1, 100000), (20000, 3000, 90000, 45000),
                                   (100000, 100000, 100000)])
def test_intersection_random_skewed_lists(ps, sizes):
    """Zipf-like skew: window staging (similar sizes) and global binary search (|B| >> |A|) paths."""
    rng = np.random.default_rng(sum(sizes))
    universe = 400_000
    lists = [np.unique(rng.integers(1, universe, m)) for m in sizes]
    freqs = [rng.integers(1, 12, len(l)) for l in lists]
    idx, pls = make_lists(ps, ps.CODEC_FREQS_ONLY, lists, freqs)
    rs = ps.intersect(pls)
    exp = ol.run_intersect(idx)
    ids, _, fr = rs.fetch()
    assert ids.tolist() == [e[0] for e in exp]
    if exp:
        order = rs.child_order().tolist()
        assert order == [c for c, _ in exp[0][1]]
        exp_fr = np.array([[f for _, f in e[1]] for e in exp], dtype=np.uint32).T
        assert (fr == exp_fr).all()


def test_intersection_edge_cases(ps):
    one = ps.PostingList.from_arrays([5, 9, 12], [1, 2, 3])
    empty = ps.PostingList.from_arrays([], [])
    assert len(ps.intersect([one, empty])) == 0
    ids, _, fr = ps.intersect([one]).fetch()
    assert ids.tolist() == [5, 9, 12] and fr[0].tolist() == [1, 2, 3]
    disjoint = ps.PostingList.from_arrays([6, 10, 13])
    assert len(ps.intersect([one, disjoint])) == 0
    with pytest.raises(RuntimeError):
        ps.PostingList.from_arrays([3, 3])  # not strictly ascending
    with pytest.raises(RuntimeError):
        ps.PostingList.from_arrays([1 << 33])  # not representable on the device


# ------------------------------------------------------------------------------------------------
# union
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("quick", [False, True])
@pytest.mark.parametrize("sizes", [(50, 400, 1500, 7), (30000, 30000), (1,), (100000, 3, 50000, 777, 12, 9000)])
def test_union_matches_oracle(ps, sizes, quick):
    rng = np.random.default_rng(len(sizes) * 7 + quick)
    lists = [np.unique(rng.integers(1, 300_000, m)) for m in sizes]
    freqs = [rng.integers(1, 9, len(l)) for l in lists]
    idx, pls = make_lists(ps, ps.CODEC_FREQS_ONLY, lists, freqs)
    rs = ps.union(pls, quick_exit=quick)
    exp = ol.run_intersect(idx, union=True, quick=quick)
    if quick:
        ids, _, _ = rs.fetch(want_freqs=False)
        assert ids.tolist() == [e[0] for e in exp]
        return
    ids, _, fr = rs.fetch()
    assert ids.tolist() == [e[0] for e in exp]
    for i in range(0, len(exp), max(1, len(exp) // 400)):
        present = {c: f for c, f in exp[i][1]}
        for c in range(len(sizes)):
            assert fr[c, i] == present.get(c, 0)


# ------------------------------------------------------------------------------------------------
# scorers
# ------------------------------------------------------------------------------------------------
def _score_setup(ps, rng, sizes, n_docs=150_000):
    lists = [np.unique(rng.integers(1, n_docs, m)) for m in sizes]
    freqs = [rng.integers(1, 40, len(l)) for l in lists]
    idx, pls = make_lists(ps, ps.CODEC_FREQS_ONLY, lists, freqs)
    doc_len = rng.integers(1, 900, n_docs + 1).astype(np.uint32)
    doc_score = rng.choice(np.array([1.0, 0.5, 0.1, 0.77, 0.0], dtype=np.float32), n_docs + 1)
    max_freq = rng.integers(1, 60, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len, doc_score, max_freq)
    P = ol.postings()
    weights = rng.choice([1.0, 0.5, 2.0, 0.3], len(sizes)).tolist()
    terms = [(w, P.orc_idf(n_docs, len(l)), P.orc_idf_bm25(n_docs, len(l))) for w, l in zip(weights, lists)]
    return idx, pls, dt, terms, doc_len, doc_score, max_freq, n_docs


@pytest.mark.parametrize("scorer", range(7))
@pytest.mark.parametrize("is_union", [False, True])
def test_scorers_bit_equal_to_oracle(ps, scorer, is_union):
    """src/ext/default.c scorers on device vs oracle/scorer_oracle.c (bit-equal to the reference's default.c)."""
    rng = np.random.default_rng(scorer * 2 + is_union)
    idx, pls, dt, terms, doc_len, doc_score, max_freq, n_docs = _score_setup(ps, rng, (30000, 50000, 20000))
    avg, aggw, min_score, tanh = 123.456, 0.7, 0.01, 7
    rs = ps.union(pls) if is_union else ps.intersect(pls)
    rs.score(scorer, terms, aggw, n_docs, avg, dt, min_score, tanh)
    ids, scores, fr = rs.fetch()
    exp = ol.run_intersect(idx, union=is_union)
    assert ids.tolist() == [e[0] for e in exp]
    # every 300th hit plus the tail, where the union's children run out one after the other and the reference's active array
    # (hence the order the leaf scores are summed in) is permuted by swap_remove_child (union_flat.rs:174-180)
    for i in sorted(set(range(0, len(exp), max(1, len(exp) // 300))) | set(range(max(0, len(exp) - 200), len(exp)))):
        doc, ch = exp[i]
        # FreqsOnly lists carry no term positions: GetSlop = children - 1 (1 for a single child), index_result.c:57-60,107
        slop = len(ch) - 1 if len(ch) > 1 else 1
        s = ol.oracle_score(scorer, [f for _, f in ch], [terms[c][1] for c, _ in ch], [terms[c][2] for c, _ in ch],
                            [terms[c][0] for c, _ in ch], aggw, int(doc_len[doc]), int(max_freq[doc]), float(doc_score[doc]),
                            n_docs, avg, slop, min_score, float(tanh))
        if scorer == ol.SCORER_DISMAX and is_union:
            s = aggw * max(terms[c][0] * f for c, f in ch)
        if scorer == ol.SCORER_BM25STD_TANH:
            assert abs(s - scores[i]) <= 1e-12 * max(1.0, abs(s))  # device tanh vs libm tanh
        else:
            assert np.float64(s).tobytes() == np.float64(scores[i]).tobytes(), (scorer, is_union, s, scores[i])


def test_bm25std_reference_golden_explainscore(ps):
    """tests/pytests/test_scorers.py:198-221: two terms with F=10 in 3 docs of length 23/35/45 -> 0.54/0.52/0.51."""
    g = G["bm25std_explain"]
    a = ps.PostingList.from_arrays([1, 2, 3], [g["freq"]] * 3)
    b = ps.PostingList.from_arrays([1, 2, 3], [g["freq"]] * 3)
    dt = ps.DocTable(3, [0] + [c[0] for c in g["cases"]])
    L = ps.lib()
    idf = L.II_CalculateIDF_BM25(g["num_docs"], g["term_docs"])
    rs = ps.intersect([a, b])
    rs.score(ps.SCORER_BM25STD, [(1.0, 0.0, idf)] * 2, 1.0, g["num_docs"], g["avg_doc_len"], dt)
    _, scores, _ = rs.fetch()
    assert [f"{s:.2f}" for s in scores] == [f"{c[1]:.2f}" for c in g["cases"]]
    for total, term, expected in G["idf"]:
        assert L.II_CalculateIDF(total, term) == expected


def test_topn_ranking(ps):
    """RPSorter order: higher score first, ties -> lower docId (src/result_processor.c:834-850)."""
    rng = np.random.default_rng(77)
    idx, pls, dt, terms, doc_len, doc_score, max_freq, n_docs = _score_setup(ps, rng, (60000, 80000))
    rs = ps.intersect(pls)
    rs.score(ps.SCORER_BM25STD, terms, 1.0, n_docs, 200.0, dt)
    ids, scores, _ = rs.fetch()
    order = sorted(range(len(ids)), key=lambda i: (-scores[i], ids[i]))
    for n in (1, 10, 100, 1000, 5000):
        ti, ts = rs.topn(n)
        k = min(n, len(ids))
        assert ti.tolist() == [int(ids[i]) for i in order[:k]]
        assert ts.tobytes() == np.array([scores[i] for i in order[:k]]).tobytes()
    # the fused one-call entry point gives the same answer
    ti, ts, total = ps.search_topn(pls, False, ps.SCORER_BM25STD, terms, 1.0, n_docs, 200.0, dt, 10)
    assert total == len(ids) and ti.tolist() == [int(ids[i]) for i in order[:10]]
    # ties: DOCSCORE gives few distinct values
    rs.score(ps.SCORER_DOCSCORE, terms, 1.0, n_docs, 200.0, dt)
    ids, scores, _ = rs.fetch()
    order = sorted(range(len(ids)), key=lambda i: (-scores[i], ids[i]))
    ti, _ = rs.topn(50)
    assert ti.tolist() == [int(ids[i]) for i in order[:50]]


def test_large_intersection_properties(ps):
    """BASELINE-scale property checks (no oracle at this size): AND is a subset of every input, ascending,
    idempotent (A AND A == A), and equals numpy's intersect1d."""
    rng = np.random.default_rng(4)
    n_docs = 20_000_000
    a = np.unique(rng.integers(1, n_docs, 4_000_000)).astype(np.uint64)
    b = np.unique(rng.integers(1, n_docs, 2_000_000)).astype(np.uint64)
    c = np.unique(rng.integers(1, n_docs, 1_300_000)).astype(np.uint64)
    pa, pb, pc = (ps.PostingList.from_arrays(x) for x in (a, b, c))
    ids, _, _ = ps.intersect([pa, pb, pc]).fetch()
    exp = np.intersect1d(np.intersect1d(a, b), c)
    assert ids.tobytes() == exp.astype(np.uint64).tobytes()
    ids2, _, _ = ps.intersect([pa, pa]).fetch(want_freqs=False)
    assert ids2.tobytes() == a.tobytes()
    u, _, _ = ps.union([pb, pc], quick_exit=True).fetch(want_freqs=False)
    assert u.tobytes() == np.union1d(b, c).astype(np.uint64).tobytes()


def test_concurrent_worker_threads_get_their_own_stream(ps):
    """RediSearch runs one iterator tree per worker thread; libii_b200 gives every calling thread its own
    stream and staging.  Eight threads searching shared posting lists concurrently must each get the
    single-threaded answer (ids, score bits, hit count)."""
    from concurrent.futures import ThreadPoolExecutor

    rng = np.random.default_rng(21)
    n_docs = 3_000_000
    sizes = [900_000, 400_000, 150_000, 60_000, 20_000, 5_000]
    ids = [np.unique(rng.integers(1, n_docs, s)).astype(np.uint64) for s in sizes]
    freqs = [rng.integers(1, 20, len(x)).astype(np.uint32) for x in ids]
    pls = [ps.PostingList.from_arrays(x, f) for x, f in zip(ids, freqs)]
    doc_len = rng.integers(50, 500, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len)
    avg = float(doc_len[1:].mean())
    queries = [(0, 1, 2), (0, 3), (1, 2, 4), (0, 1), (2, 3, 5), (0, 5), (1, 4), (0, 2, 3)]

    def run(q):
        terms = [(1.0, ps.lib().II_CalculateIDF(n_docs, len(ids[i])), ps.lib().II_CalculateIDF_BM25(n_docs, len(ids[i]))) for i in q]
        return ps.search_topn([pls[i] for i in q], False, ps.SCORER_BM25STD, terms, 1.0, n_docs, avg, dt, 10)

    expect = [run(q) for q in queries]
    with ThreadPoolExecutor(max_workers=len(queries)) as pool:
        for _ in range(5):
            got = list(pool.map(run, queries))
            for (gi, gs, gt), (ei, es, et) in zip(got, expect):
                assert gi.tolist() == ei.tolist() and gs.tobytes() == es.tobytes() and gt == et
    # the batch entry point (pool of streams inside the library) returns the same rows
    batch = ps.SearchBatch([([pls[i] for i in q],
                             [(1.0, ps.lib().II_CalculateIDF(n_docs, len(ids[i])), ps.lib().II_CalculateIDF_BM25(n_docs, len(ids[i]))) for i in q])
                            for q in queries * 3], 10)
    for _ in range(3):
        got = batch.run(False, ps.SCORER_BM25STD, 1.0, n_docs, avg, dt)
        for (gi, gs, gt), (ei, es, et) in zip(got, expect * 3):
            assert gi.tolist() == ei.tolist() and gs.tobytes() == es.tobytes() and gt == et
    # and the sequential answers are the oracle's: hit counts equal numpy's intersection
    for q, (_, _, tot) in zip(queries, expect):
        ref = ids[q[0]]
        for i in q[1:]:
            ref = np.intersect1d(ref, ids[i])
        assert tot == len(ref)


def test_docid_range_shards_merge_to_the_unsharded_topn(ps):
    """SURVEY.md §8e on one GPU: cut every list at the shard boundaries, search each slice with the global
    statistics, merge with II_MergeShardTopN — identical to searching the whole lists."""
    from redisearch_b200 import sharding

    rng = np.random.default_rng(31)
    n_docs, top = 2_000_000, 25
    lists = [np.unique(rng.integers(1, n_docs + 1, m)).astype(np.uint64) for m in (700_000, 250_000, 1_100_000)]
    freqs = [rng.integers(1, 12, len(l)).astype(np.uint32) for l in lists]
    doc_len = rng.integers(50, 500, n_docs + 1).astype(np.uint32)
    avg = float(doc_len[1:].mean())
    dt = ps.DocTable(n_docs, doc_len)
    terms = [(1.0, ps.lib().II_CalculateIDF(n_docs, len(l)), ps.lib().II_CalculateIDF_BM25(n_docs, len(l))) for l in lists]
    whole = [ps.PostingList.from_arrays(l, f) for l, f in zip(lists, freqs)]
    e_ids, e_sc, e_tot = ps.search_topn(whole, False, ps.SCORER_BM25STD, terms, 1.0, n_docs, avg, dt, top)
    for world in (2, 3, 8):
        sc = np.full((world, top), np.nan)
        ids = np.zeros((world, top), dtype=np.uint64)
        cnt = np.zeros(world, dtype=np.uint64)
        total = 0
        for g in range(world):
            lo, hi = sharding.doc_range(n_docs, world, g)
            sl = [sharding.split_posting_list(l, f, lo, hi) for l, f in zip(lists, freqs)]
            pls = [ps.PostingList.from_arrays(l, f) for l, f in sl]
            gi, gs, gt = ps.search_topn(pls, False, ps.SCORER_BM25STD, terms, 1.0, n_docs, avg, dt, top)
            ids[g, :len(gi)], sc[g, :len(gi)], cnt[g] = gi, gs, len(gi)
            total += gt
        out_i = np.zeros(top, dtype=np.uint64)
        out_s = np.zeros(top, dtype=np.float64)
        got = ps.lib().II_MergeShardTopN(sc.ctypes.data, ids.ctypes.data, cnt.ctypes.data, world, top, top, out_i.ctypes.data,
                                         out_s.ctypes.data)
        assert total == e_tot and got == len(e_ids)
        assert out_i[:got].tolist() == e_ids.tolist() and out_s[:got].tobytes() == e_sc.tobytes()


def test_union_reference_edge_cases_through_the_iterator(ps):
    """rqe_iterators/tests/integration/union_common.rs:243-452 — the reference's known answers for Union (disjoint,
    overlapping, empty children, skip_to exact / not found / past EOF + rewind, interleaved read and skip_to), through
    II_Union and the QueryIterator facade."""
    E = G["union_edge_cases"]

    def lists_of(children):
        return [ps.PostingList.from_arrays(np.array(c, dtype=np.uint64)) for c in children]

    for name in ("disjoint", "overlapping", "empty_mixed", "all_empty"):
        for quick in (False, True):
            ids, _, _ = ps.union(lists_of(E[name]["children"]), quick_exit=quick).fetch(want_freqs=False)
            assert ids.tolist() == E[name]["expected"], (name, quick, ids)
    it = ps.union(lists_of(E["skip_exact"]["children"])).into_iterator()
    q = it.contents
    assert q.SkipTo(it, 30) == ps.ITERATOR_OK and q.lastDocId == 30
    q.Rewind(it)
    assert q.SkipTo(it, 22) == ps.ITERATOR_NOTFOUND and q.lastDocId == 25
    q.Free(it)
    it = ps.union(lists_of(E["skip_past_eof"]["children"])).into_iterator()
    q = it.contents
    assert q.SkipTo(it, 100) == ps.ITERATOR_EOF and q.atEOF
    q.Rewind(it)
    assert not q.atEOF and q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 10
    q.Free(it)
    # union_common.rs:331-352
    it = ps.union(lists_of([[10, 20, 30, 40, 50, 60, 70, 80], [15, 25, 35, 45, 55, 65, 75, 85]])).into_iterator()
    q = it.contents
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 10
    assert q.SkipTo(it, 35) == ps.ITERATOR_OK and q.lastDocId == 35
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 40
    assert q.SkipTo(it, 70) == ps.ITERATOR_OK and q.lastDocId == 70
    q.Free(it)


def test_intersection_reference_edge_cases_through_the_iterator(ps):
    """rqe_iterators/tests/integration/intersection.rs:325-521, 876-931 through II_Intersect and the QueryIterator
    facade: empty / single-element / single-child / overlapping result sets, 10^7 docId gaps, skip_to exact / not found /
    past EOF (then Read and SkipTo stay at EOF until Rewind), sequential and interleaved skip_to."""
    E = G["intersection_edge_cases"]

    def lists_of(children):
        return [ps.PostingList.from_arrays(np.array(c, dtype=np.uint64)) for c in children]

    status = {0: ps.ITERATOR_OK, 1: ps.ITERATOR_NOTFOUND, 2: ps.ITERATOR_EOF}
    for name, case in E.items():
        if name.startswith("_"):
            continue
        if "expected" in case:
            ids, _, _ = ps.intersect(lists_of(case["children"])).fetch(want_freqs=False)
            assert ids.tolist() == case["expected"], name
        if "skips" in case:
            it = ps.intersect(lists_of(case["children"])).into_iterator()
            q = it.contents
            for target, st, landed in case["skips"]:
                q.Rewind(it)
                assert q.SkipTo(it, target) == status[st], (name, target)
                if st != 2:
                    assert q.lastDocId == landed
                else:
                    assert q.atEOF and q.Read(it) == ps.ITERATOR_EOF and q.SkipTo(it, 10) == ps.ITERATOR_EOF
            q.Free(it)
    ids10 = list(range(10, 101, 10))
    it = ps.intersect(lists_of([ids10, ids10])).into_iterator()
    q = it.contents
    for i in ids10[:5]:  # skip_to_sequential
        assert q.SkipTo(it, i) == ps.ITERATOR_OK and q.lastDocId == i
    q.Rewind(it)         # interleaved_read_and_skip_to
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 10
    assert q.SkipTo(it, 40) == ps.ITERATOR_OK and q.lastDocId == 40
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 50
    assert q.SkipTo(it, 80) == ps.ITERATOR_OK and q.lastDocId == 80
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 90
    assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 100
    assert q.Read(it) == ps.ITERATOR_EOF
    q.Free(it)


@pytest.mark.parametrize("scorer", [0, 1, 2, 3, 4, 5, 6])
def test_fused_batch_search_equals_the_per_query_chains(ps, scorer):
    """II_SearchTopNBatch runs AND queries with <= 8 terms and top_n <= 128 FUSED (two launches for the whole batch:
    membership + scorer + per-chunk top-N, then per-query top-N).  Every query's rows must be the rows of II_SearchTopN
    (the per-query kernel chain, itself pinned to the oracle by the tests above): docIds, score bits, hit counts — for
    every scorer, 1..8 terms, ties in the score (DOCSCORE: all equal -> docId order), empty children and top_n 1 / 10 / 128;
    a 9-term query in the same batch takes the chain path."""
    rng = np.random.default_rng(100 + scorer)
    n_docs = 400_000
    sizes = [150_000, 90_000, 60_000, 30_000, 12_000, 5_000, 2_500, 900, 300, 1]
    ids = [np.unique(rng.integers(1, n_docs, s)).astype(np.uint64) for s in sizes]
    ids[3] = np.union1d(ids[3], ids[4][:4000])  # make the deeper ANDs non-empty
    for j in range(5, 9):
        ids[j] = np.union1d(ids[j], ids[4][:300])
    freqs = [rng.integers(1, 30, len(x)).astype(np.uint32) for x in ids]
    pls = [ps.PostingList.from_arrays(x, f) for x, f in zip(ids, freqs)]
    empty = ps.PostingList.from_arrays(np.zeros(0, dtype=np.uint64))
    doc_len = rng.integers(50, 500, n_docs + 1).astype(np.uint32)
    doc_score = (rng.integers(1, 4, n_docs + 1) / 2.0).astype(np.float32)
    max_freq = rng.integers(1, 40, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len, doc_score, max_freq)
    avg = float(doc_len[1:].mean())
    L = ps.lib()

    def terms_of(q):
        return [(1.0 + 0.25 * (i % 3), L.II_CalculateIDF(n_docs, max(1, len(ids[i]) if i >= 0 else 1)),
                 L.II_CalculateIDF_BM25(n_docs, max(1, len(ids[i]) if i >= 0 else 1))) for i in q]

    queries = [(0, 1), (0, 1, 2), (2, 1, 0), (3, 4), (0,), (9,), (4, 3, 2, 1, 0), (0, 1, 2, 3, 4, 5, 6, 7), (8, 7, 6, 5), (0, 9),
               (1, 2, 3, 4, 5, 6, 7, 8, 0), (5, 6), (0, 1, 4), (2, 3)]
    for top_n in (1, 10, 128):
        qs = []
        for q in queries:
            qs.append(([pls[i] for i in q], terms_of(q)))
        qs.append(([pls[0], empty, pls[1]], terms_of((0, -1, 1))))  # an empty child: the AND is empty
        batch = ps.SearchBatch(qs, top_n)
        got = batch.run(False, scorer, 1.5, n_docs, avg, dt)
        for (lists, terms), (gi, gs, gt) in zip(qs, got):
            ei, es, et = ps.search_topn(lists, False, scorer, terms, 1.5, n_docs, avg, dt, top_n)
            assert gt == et, (scorer, top_n, gt, et)
            assert gi.tolist() == ei.tolist(), (scorer, top_n, len(lists))
            assert gs.tobytes() == es.tobytes()
    # batches whose largest child count is 2 and 3: the kernel is instantiated per child count (2 / 3 / 4 / 8)
    for kmax in (2, 3):
        small = []
        for _ in range(40):
            q = tuple(int(x) for x in rng.choice(6, int(rng.integers(1, kmax + 1)), replace=False))
            small.append(([pls[i] for i in q], terms_of(q)))
        small.append(([pls[i] for i in range(kmax)], terms_of(tuple(range(kmax)))))
        got = ps.SearchBatch(small, 10).run(False, scorer, 1.0, n_docs, avg, dt)
        for (lists, terms), (gi, gs, gt) in zip(small, got):
            ei, es, et = ps.search_topn(lists, False, scorer, terms, 1.0, n_docs, avg, dt, 10)
            assert gt == et and gi.tolist() == ei.tolist() and gs.tobytes() == es.tobytes(), (kmax, len(lists))
    # a few hundred random queries in one call
    many = []
    for _ in range(300):
        k = int(rng.integers(1, 5))
        q = tuple(int(x) for x in rng.choice(9, k, replace=False))
        many.append(([pls[i] for i in q], terms_of(q)))
    got = ps.SearchBatch(many, 10).run(False, scorer, 1.0, n_docs, avg, dt)
    for (lists, terms), (gi, gs, gt) in list(zip(many, got))[::7]:
        ei, es, et = ps.search_topn(lists, False, scorer, terms, 1.0, n_docs, avg, dt, 10)
        assert gt == et and gi.tolist() == ei.tolist() and gs.tobytes() == es.tobytes()


def _block_views(ps, blocks):
    """list of (first, last, n, bytes) -> (II_BlockView array, keep-alive buffers)"""
    arr = (ps.II_BlockView * max(1, len(blocks)))()
    keep = []
    for i, (first, last, n, data) in enumerate(blocks):
        buf = (C.c_uint8 * max(1, len(data))).from_buffer_copy(data or b"\0")
        keep.append(buf)
        arr[i] = ps.II_BlockView(first, last, n, C.cast(buf, C.POINTER(C.c_uint8)), len(data))
    return arr, keep


@pytest.mark.parametrize("codec", range(ol.N_CODECS))
def test_batch_decode_of_many_lists_matches_the_oracle_reader(ps, codec):
    """II_PostingList_FromBlocksBatch: the blocks of MANY lists in one gather / one copy / one decode launch (bytes staged in
    shared memory, one thread per block).  Every list must read back exactly like the oracle reader: long lists, a one-entry
    list, an empty list, wide deltas (4-byte qint fields) and, for the Full codec, records with offset payloads."""
    rng = np.random.default_rng(40 + codec)
    specs = [(7000, 50), (1, 10), (0, 1), (333, 3_000_000), (12_345, 9), (100, 70_000), (2501, 300)]
    idx, views, keep, nbl = [], [], [], []
    for n, gap in specs:
        ix = ol.InvIndex(codec)
        doc = 0
        for _ in range(n):
            doc += int(rng.integers(1, gap + 1))
            off = bytes(rng.integers(0, 255, int(rng.integers(0, 9))).astype(np.uint8)) if codec in ol.CODECS_WITH_OFFSETS else b""
            mask = int(rng.integers(1, 1 << 30)) << (int(rng.integers(0, 98)) if codec in ol.CODECS_WIDE else 0)
            ix.add(doc, int(rng.integers(1, 1 << int(rng.integers(1, 31)))), mask, off)
        bl = ix.blocks()
        arr, k = _block_views(ps, bl)
        idx.append(ix)
        views.append(arr)
        keep.append(k)
        nbl.append(len(bl))
    L = ps.lib()
    ptrs = (C.c_void_p * len(specs))(*[C.cast(v, C.c_void_p) for v in views])
    ns = (C.c_size_t * len(specs))(*nbl)
    out = (C.c_void_p * len(specs))()
    assert L.II_PostingList_FromBlocksBatch(len(specs), ptrs, ns, codec, out) == len(specs)
    for ix, h in zip(idx, out):
        pl = ps.PostingList(h)
        exp = ix.read_all()
        assert len(pl) == len(exp)
        if not exp:
            continue
        got_ids, _, got_fr = ps.union([pl]).fetch()
        assert got_ids.tolist() == [e[0] for e in exp]
        assert got_fr[0].tolist() == [e[1] for e in exp]


def test_term_cache_hits_versions_pins_and_eviction(ps):
    """II_TermCache: a (key, version) hit returns the resident list without touching the blocks; a new version (the index was
    written / collected: gc_marker, RS/inverted_index/src/reader/core.rs:372-374) rebuilds it; pinned lists survive
    replacement until released; LRU eviction keeps the resident bytes under the budget."""
    L = ps.lib()
    rng = np.random.default_rng(77)

    def make(n, seed):
        r = np.random.default_rng(seed)
        ids = np.cumsum(r.integers(1, 50, n)).astype(np.uint64)
        ix = ol.InvIndex(ol.CODEC_FREQS_ONLY, ids, r.integers(1, 9, n).tolist())
        arr, keep = _block_views(ps, ix.blocks())
        return ix, arr, keep, len(ix.blocks())

    terms = [make(5000 + 1000 * i, i) for i in range(6)]
    cache = L.II_TermCache_New(8 * (5000 + 6000 + 7000) + 64)  # room for about three of the lists

    def acquire(which, versions):
        n = len(which)
        keys = (C.c_uint64 * n)(*[1000 + w for w in which])
        vers = (C.c_uint64 * n)(*versions)
        bl = (C.c_void_p * n)(*[C.cast(terms[w][1], C.c_void_p) for w in which])
        nb = (C.c_size_t * n)(*[terms[w][3] for w in which])
        out = (C.c_void_p * n)()
        assert L.II_TermCache_Acquire(cache, n, keys, vers, bl, nb, ol.CODEC_FREQS_ONLY, out) == n
        return out

    def check(handle, w):
        h = C.c_void_p(handle)
        arr = (C.c_void_p * 1)(h)
        rs = L.II_Union(arr, 1, 0)
        m = L.II_ResultSet_Len(rs)
        ids = np.zeros(m, dtype=np.uint64)
        assert L.II_ResultSet_Fetch(rs, ids.ctypes.data, None, None) == 0
        L.II_ResultSet_Free(rs)
        assert ids.tolist() == [e[0] for e in terms[w][0].read_all()]

    a = acquire([0, 1, 0], [1, 1, 1])  # the same term twice in one batch: decoded once
    assert a[0] == a[2]
    st = L.II_TermCache_GetStats(cache)
    assert (st.misses, st.hits, st.resident_lists) == (2, 0, 2)
    check(a[0], 0)
    check(a[1], 1)
    L.II_TermCache_Release(cache, 3, a)
    b = acquire([0, 1], [1, 1])
    st = L.II_TermCache_GetStats(cache)
    assert (st.misses, st.hits) == (2, 2) and b[0] == a[0] and b[1] == a[1]
    # version bump of term 0 while it is still pinned by `b`: a fresh list is built, the old handle stays valid until released
    c2 = acquire([0], [2])
    assert c2[0] != b[0]
    check(b[0], 0)
    check(c2[0], 0)
    L.II_TermCache_Release(cache, 2, b)
    L.II_TermCache_Release(cache, 1, c2)
    # fill past the budget: least recently used unpinned lists go
    for w in (2, 3, 4, 5):
        h = acquire([w], [1])
        check(h[0], w)
        L.II_TermCache_Release(cache, 1, h)
    st = L.II_TermCache_GetStats(cache)
    assert st.evictions >= 2 and st.resident_bytes <= 8 * (5000 + 6000 + 7000) + 64
    L.II_TermCache_Invalidate(cache, 1005)
    assert L.II_TermCache_GetStats(cache).resident_lists == st.resident_lists - 1
    L.II_TermCache_Free(cache)


def test_and_with_not_and_optional_children(ps):
    """"a b -c ~d": required terms intersect, NOT children exclude (not.rs as a child of an intersection), OPTIONAL children never
    reject (optional.rs) and contribute only where present.  Excluded / absent children are the reference's virtual results:
    freq 0 and nothing in any scorer (src/ext/default.c:289-297).  DocIds vs numpy set algebra, freqs per child in
    aggregate order, BM25STD / TFIDF scores bit-equal to the scorer oracle fed with the non-virtual children."""
    rng = np.random.default_rng(55)
    n_docs = 500_000
    ids = [np.unique(rng.integers(1, n_docs, s)).astype(np.uint64) for s in (200_000, 90_000, 150_000, 60_000)]
    freqs = [rng.integers(1, 30, len(x)).astype(np.uint32) for x in ids]
    pls = [ps.PostingList.from_arrays(x, f) for x, f in zip(ids, freqs)]
    doc_len = rng.integers(50, 500, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len)
    avg = float(doc_len[1:].mean())
    L = ps.lib()
    P = ol.postings()
    maps = [dict(zip(x.tolist(), f.tolist())) for x, f in zip(ids, freqs)]
    for modes in ([0, 0, 1, 2], [0, 1, 0, 0], [2, 0, 0, 1], [0, 2, 2, 2], [0, 1, 1, 1]):
        arr = (C.c_void_p * 4)(*[p.h for p in pls])
        marr = (C.c_int * 4)(*modes)
        rs = ps.ResultSet(L.II_IntersectEx(arr, marr, 4))
        exp = None
        for i, m in enumerate(modes):
            if m == 0:
                exp = ids[i] if exp is None else np.intersect1d(exp, ids[i])
        for i, m in enumerate(modes):
            if m == 1:
                exp = np.setdiff1d(exp, ids[i])
        terms = [(1.0 + 0.5 * i, P.orc_idf(n_docs, len(ids[i])), P.orc_idf_bm25(n_docs, len(ids[i]))) for i in range(4)]
        for scorer in (ps.SCORER_BM25STD, ps.SCORER_TFIDF):
            rs.score(scorer, terms, 1.25, n_docs, avg, dt)
            got_ids, got_sc, got_fr = rs.fetch()
            assert got_ids.tolist() == exp.tolist(), modes
            order = rs.child_order().tolist()
            # required children first (ascending estimate), NOT / OPTIONAL children behind in their given order
            req = sorted([i for i in range(4) if modes[i] == 0], key=lambda i: len(ids[i]))
            assert order == req + [i for i in range(4) if modes[i] != 0], (modes, order)
            for j in range(0, len(exp), max(1, len(exp) // 60)):
                d = int(exp[j])
                fr, idf, bidf, w = [], [], [], []
                for slot, c in enumerate(order):
                    present = modes[c] != 1 and d in maps[c]
                    assert got_fr[slot][j] == (maps[c][d] if present else 0)
                    if present:
                        fr.append(maps[c][d]); idf.append(terms[c][1]); bidf.append(terms[c][2]); w.append(terms[c][0])
                # 4 aggregate children (the virtual ones included), no term positions: GetSlop = 4 - 1 (index_result.c:107)
                s = ol.oracle_score(scorer, fr, idf, bidf, w, 1.25, int(doc_len[d]), 1, 1.0, n_docs, avg, 3)
                assert np.float64(s).tobytes() == np.float64(got_sc[j]).tobytes(), (modes, scorer, d)
    # no required child: refused
    arr = (C.c_void_p * 2)(pls[0].h, pls[1].h)
    assert not L.II_IntersectEx(arr, (C.c_int * 2)(1, 2), 2)


def _positions_to_offsets(positions):
    """term positions (ascending) -> the offsets payload of a Full-codec record: varint deltas"""
    out, last = b"", 0
    for p in positions:
        buf = (C.c_uint8 * 16)()
        n = ol.postings().orc_varint_encode(int(p) - last, buf)
        out += bytes(buf[:n])
        last = int(p)
    return out


def _phrase_corpus(rng, n_docs, n_terms, density, doc_words, codec=ol.CODEC_FULL):
    """indexes of n_terms terms with random term positions (a codec that stores them); returns (indexes, per term {doc: offsets bytes})"""
    idx, offs = [], []
    for t in range(n_terms):
        ix = ol.InvIndex(codec)
        docs = np.flatnonzero(rng.random(n_docs) < density[t]) + 1
        m = {}
        for d in docs.tolist():
            k = int(rng.integers(1, 6))
            pos = np.unique(rng.integers(1, doc_words, k))
            if rng.random() < 0.02:
                pos = pos[:0]  # a record without positions: left out of the check by the reference
            if rng.random() < 0.03:
                pos = np.unique(np.concatenate([pos, rng.integers(200, 70_000, 2)]))  # multi-byte varints
            ob = _positions_to_offsets(pos.tolist())
            m[d] = ob
            ix.add(d, max(1, len(pos)), int(rng.integers(1, 1 << 20)), ob)
        idx.append(ix)
        offs.append(m)
    return idx, offs


@pytest.mark.parametrize("in_order", [False, True])
@pytest.mark.parametrize("n_terms", [2, 3, 5])
def test_phrase_intersection_slop_and_order(ps, in_order, n_terms):
    """Exact phrases / "within N words": II_IntersectPhrase against the oracle's restatement of the reference's proximity check
    (RS/index_result/src/core/proximity.rs within_range_in_order / within_range_unordered, pinned on the reference's own
    known answers in test_oracle_postings.py) applied to every hit of the oracle intersection, for several slops incl. none.
    DocIds, per-child freqs and the aggregate child order (given order when in_order) must be identical."""
    rng = np.random.default_rng(900 + n_terms + 10 * in_order)
    n_docs = 60_000
    density = [0.5, 0.35, 0.6, 0.45, 0.7][:n_terms]
    idx, offs = _phrase_corpus(rng, n_docs, n_terms, density, 40)
    pls = ps.postings_with_offsets([ix.blocks() for ix in idx], ol.CODEC_FULL)
    assert all(ps.lib().II_PostingList_HasOffsets(p.h) for p in pls)
    hits = ol.run_intersect(idx)
    freq_of = [dict((d, f) for d, f, _ in ix.read_all()) for ix in idx]
    sorted_order = [c for c, _ in hits[0][1]] if hits else list(range(n_terms))
    for slop in (0, 1, 3, 10, None):
        if slop is None and not in_order:
            continue
        rs = ps.intersect_phrase(pls, slop, in_order)
        order = rs.child_order().tolist()
        assert order == (list(range(n_terms)) if in_order else sorted_order)
        exp = [d for d, _ in hits if ol.within_range([offs[c][d] for c in order], slop, in_order)]
        got_ids, _, got_fr = rs.fetch()
        assert got_ids.tolist() == exp, (slop, in_order, len(got_ids), len(exp))
        assert 0 < len(exp) < len(hits) or slop is None or slop >= 10
        for slot, c in enumerate(order):
            assert got_fr[slot].tolist() == [freq_of[c][d] for d in exp]
    # no constraint at all = the plain intersection
    rs = ps.intersect_phrase(pls, None, False)
    assert rs.fetch()[0].tolist() == [d for d, _ in hits]


def _decode_offsets(ob):
    pos, last, i = [], 0, 0
    while i < len(ob):
        b = ob[i]
        i += 1
        v = b & 0x7F
        while b & 0x80:
            b = ob[i]
            i += 1
            v = ((v + 1) << 7) | (b & 0x7F)
        last += v
        pos.append(last)
    return pos


@pytest.mark.parametrize("scorer", [ol.SCORER_BM25, ol.SCORER_TFIDF, ol.SCORER_TFIDF_DOCNORM])
@pytest.mark.parametrize("shape", ["and3", "and5", "or3", "and_not_optional", "phrase"])
def test_legacy_scorers_divide_by_the_slop_of_the_hit(ps, scorer, shape):
    """BM25 / TFIDF / TFIDF.DOCNORM divide by GetSlop = IndexResult_MinOffsetDelta (src/index_result/index_result.c:51-108) of the
    hit.  With term positions on the device the kernel walks them pair by pair like the reference; the expected value is the
    oracle restatement (pinned on the reference's own compiled index_result.c in test_oracle_postings.py) over the decoded
    positions of the hit's children in aggregate order.  Scores bit-equal."""
    n_terms = {"and3": 3, "and5": 5, "or3": 3, "and_not_optional": 4, "phrase": 3}[shape]
    rng = np.random.default_rng(1200 + scorer * 10 + n_terms + len(shape))
    n_docs = 40_000
    density = [0.5, 0.35, 0.6, 0.45, 0.7][:n_terms]
    if shape == "or3":
        density = [0.02, 0.03, 0.015]
    idx, offs = _phrase_corpus(rng, n_docs, n_terms, density, 40)
    pls = ps.postings_with_offsets([ix.blocks() for ix in idx], ol.CODEC_FULL)
    doc_len = rng.integers(1, 900, n_docs + 1).astype(np.uint32)
    doc_score = rng.choice(np.array([1.0, 0.5, 0.77], dtype=np.float32), n_docs + 1)
    max_freq = rng.integers(1, 60, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len, doc_score, max_freq)
    P = ol.postings()
    weights = rng.choice([1.0, 0.5, 2.0], n_terms).tolist()
    terms = [(w, P.orc_idf(n_docs, ix.num_docs()), P.orc_idf_bm25(n_docs, ix.num_docs())) for w, ix in zip(weights, idx)]
    avg, aggw = 222.5, 0.7
    freq_of = [dict((d, f) for d, f, _ in ix.read_all()) for ix in idx]
    if shape == "or3":
        rs = ps.union(pls)
        exp = ol.run_intersect(idx, union=True)
        rows = [(d, [(c, f, False) for c, f in ch]) for d, ch in exp]
    elif shape == "and_not_optional":
        modes = [0, 0, 1, 2]  # t0 AND t1 AND NOT t2 AND OPTIONAL t3
        rs = ps.intersect_ex(pls, modes)
        order = rs.child_order().tolist()
        members = [set(freq_of[c]) for c in range(4)]
        docs = sorted((members[0] & members[1]) - members[2])
        rows = [(d, [(c, freq_of[c].get(d, 0) if modes[c] != 1 else 0, modes[c] == 1 or (modes[c] == 2 and d not in members[c]))
                     for c in order]) for d in docs]
    elif shape == "phrase":
        rs = ps.intersect_phrase(pls, 4, False)
        order = rs.child_order().tolist()
        rows = [(d, [(c, freq_of[c][d], False) for c in order]) for d, _ in ol.run_intersect(idx)
                if ol.within_range([offs[c][d] for c in order], 4, False)]
    else:
        rs = ps.intersect(pls)
        rows = [(d, [(c, f, False) for c, f in ch]) for d, ch in ol.run_intersect(idx)]
    rs.score(scorer, terms, aggw, n_docs, avg, dt, 0.0, 4)
    ids, scores, _ = rs.fetch()
    assert ids.tolist() == [d for d, _ in rows] and len(rows) > 50
    slops = set()
    for i in sorted(set(range(0, len(rows), max(1, len(rows) // 400))) | set(range(max(0, len(rows) - 100), len(rows)))):
        doc, ch = rows[i]
        positions = [[] if virt else _decode_offsets(offs[c][doc]) for c, _, virt in ch]
        slop = ol.min_offset_delta(positions, [virt for _, _, virt in ch])
        slops.add(slop)
        real = [(c, f) for c, f, virt in ch if f]
        s = ol.oracle_score(scorer, [f for _, f in real], [terms[c][1] for c, _ in real], [terms[c][2] for c, _ in real],
                            [terms[c][0] for c, _ in real], aggw, int(doc_len[doc]), int(max_freq[doc]), float(doc_score[doc]),
                            n_docs, avg, slop, 0.0, 4.0)
        assert np.float64(s).tobytes() == np.float64(scores[i]).tobytes(), (shape, scorer, doc, slop, s, scores[i])
    assert len(slops) >= 3, slops  # the data exercises several different slop values


def test_hamming_scorer_matches_the_reference(ps):
    """HAMMING (src/ext/default.c:475-497): 1 / (bit distance of the payloads + 1), 0 without a payload or with another length;
    checked against the reference's own default.c when oracle/_ref is built, else against the formula."""
    rng = np.random.default_rng(4711)
    n_docs = 30_000
    ids = [np.unique(rng.integers(1, n_docs, s)).astype(np.uint64) for s in (20_000, 9_000)]
    pls = [ps.PostingList.from_arrays(x, np.ones(len(x), dtype=np.uint32)) for x in ids]
    payloads = [None] * (n_docs + 1)
    for d in range(1, n_docs + 1):
        r = rng.random()
        if r < 0.7:
            payloads[d] = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
        elif r < 0.8:
            payloads[d] = rng.integers(0, 256, int(rng.integers(1, 40)), dtype=np.uint8).tobytes()
    dt = ps.DocTable(n_docs)
    dt.set_payloads(payloads)
    rs = ps.intersect(pls)
    q = rng.integers(0, 256, 16, dtype=np.uint8).tobytes()
    rs.score_hamming(dt, q)
    got_ids, got, _ = rs.fetch()
    assert got_ids.tolist() == np.intersect1d(ids[0], ids[1]).tolist() and len(got_ids) > 1000
    seen = set()
    for i, d in enumerate(got_ids.tolist()):
        p = payloads[d]
        if not p or len(p) != len(q):
            exp = 0.0
        else:
            bits = int(np.unpackbits(np.frombuffer(p, np.uint8) ^ np.frombuffer(q, np.uint8)).sum())
            exp = 1.0 / (bits + 1)
        if ol.ref_scorers() is not None and i % 7 == 0:
            assert ol.reference_hamming(p, q) == exp
        assert got[i] == exp, (d, got[i], exp)
        seen.add(exp == 0.0)
    assert seen == {True, False}
    rs.score_hamming(dt, b"")  # an empty query payload never matches (payload length 0 is "no payload")
    assert not rs.fetch()[1].any()


@pytest.mark.parametrize("codec", [ol.CODEC_FREQS_OFFSETS, ol.CODEC_OFFSETS_ONLY, ol.CODEC_FIELDS_OFFSETS, ol.CODEC_FULL_WIDE,
                                   ol.CODEC_FIELDS_OFFSETS_WIDE])
def test_phrase_intersection_over_the_other_offset_codecs(ps, codec):
    """every codec that stores term positions keeps them on the device: same phrase answers as the Full codec path"""
    rng = np.random.default_rng(1500 + codec)
    idx, offs = _phrase_corpus(rng, 30_000, 3, [0.5, 0.4, 0.6], 40, codec)
    pls = ps.postings_with_offsets([ix.blocks() for ix in idx], codec)
    assert all(ps.lib().II_PostingList_HasOffsets(p.h) for p in pls)
    hits = ol.run_intersect(idx)
    freq_of = [dict((d, f) for d, f, _ in ix.read_all()) for ix in idx]
    for slop, in_order in ((0, True), (2, False), (5, True)):
        rs = ps.intersect_phrase(pls, slop, in_order)
        order = rs.child_order().tolist()
        exp = [d for d, _ in hits if ol.within_range([offs[c][d] for c in order], slop, in_order)]
        got_ids, _, got_fr = rs.fetch()
        assert got_ids.tolist() == exp and 0 < len(exp) < len(hits)
        for slot, c in enumerate(order):
            assert got_fr[slot].tolist() == [freq_of[c][d] for d in exp]


def test_phrase_constructor_takes_slop_and_in_order(ps):
    """NewIntersectionIterator(max_slop >= 0 / in_order) over B200 term leaves that carry positions is evaluated on the device;
    leaves without positions are refused with nothing consumed (the caller keeps the reference's iterator)."""
    rng = np.random.default_rng(931)
    idx, offs = _phrase_corpus(rng, 20_000, 3, [0.5, 0.4, 0.6], 30)
    L = ps.lib()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def its_of(pls):
        arr = libc.malloc(8 * len(pls))
        view = (C.c_void_p * len(pls)).from_address(arr)
        leaves = []
        for i, p in enumerate(pls):
            leaf = L.II_NewTermIterator(p.h, 0, 1.0, 1.0, 1.0)
            view[i] = C.cast(leaf, C.c_void_p).value
            leaves.append(leaf)
        return arr, leaves

    hits = ol.run_intersect(idx)
    pls = ps.postings_with_offsets([ix.blocks() for ix in idx], ol.CODEC_FULL)
    for slop, in_order in ((0, True), (2, False), (-1, True)):
        arr, _ = its_of(pls)
        qi = L.NewIntersectionIterator(arr, 3, slop, in_order, 1.0)
        assert qi
        got = []
        while qi.contents.Read(qi) == 0:
            got.append(qi.contents.lastDocId)
        order = list(range(3)) if in_order else [c for c, _ in hits[0][1]]
        exp = [d for d, _ in hits if ol.within_range([offs[c][d] for c in order], None if slop < 0 else slop, in_order)]
        assert got == exp and 0 < len(exp) < len(hits)
        qi.contents.Free(qi)
    # leaves decoded WITHOUT positions: refused, children untouched
    plain = [ps.PostingList.from_blocks(ix.blocks(), ol.CODEC_FULL, on_device=True) for ix in idx]
    arr, leaves = its_of(plain)
    assert not L.NewIntersectionIterator(arr, 3, 0, True, 1.0)
    for lf in leaves:
        lf.contents.Free(lf)
    libc.free.argtypes = [C.c_void_p]
    libc.free(arr)


def test_wildcard_children_of_the_constructors(ps):
    """A wildcard child (rqe_iterators/src/wildcard.rs) under the constructors: stripped by an AND (intersection.rs:363-417), it
    takes over a quick OR (union_reducer.rs:41-53), and inside a full OR it contributes every docId 1..top_id."""
    rng = np.random.default_rng(77)
    top = 5000
    a = np.unique(rng.integers(1, top, 700)).astype(np.uint64)
    b = np.unique(rng.integers(1, top, 900)).astype(np.uint64)
    pa, pb = ps.PostingList.from_arrays(a), ps.PostingList.from_arrays(b)
    L = ps.lib()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def its_of(children):
        arr = libc.malloc(8 * len(children))
        view = (C.c_void_p * len(children)).from_address(arr)
        for i, c in enumerate(children):
            view[i] = C.cast(c, C.c_void_p).value
        return arr

    def drain(qi):
        got = []
        while qi.contents.Read(qi) == 0:
            got.append(qi.contents.lastDocId)
        qi.contents.Free(qi)
        return got

    leaf = lambda p: L.II_NewTermIterator(p.h, 0, 1.0, 1.0, 1.0)
    # AND(a, *, b) == AND(a, b)
    qi = L.NewIntersectionIterator(its_of([leaf(pa), L.II_NewWildcardIterator(top, 1.0), leaf(pb)]), 3, -1, False, 1.0)
    assert drain(qi) == np.intersect1d(a, b).tolist()
    # quick OR with a wildcard child IS the wildcard
    qi = L.NewUnionIterator(its_of([leaf(pa), L.II_NewWildcardIterator(top, 1.0)]), 2, True, 1.0, 0, None, None)
    assert qi.contents.type == 12 and drain(qi) == list(range(1, top + 1))
    # full OR: every document, the term children's freqs where present
    qi = L.NewUnionIterator(its_of([leaf(pa), L.II_NewWildcardIterator(top, 1.0), leaf(pb)]), 3, False, 1.0, 0, None, None)
    assert qi.contents.type == 6 and drain(qi) == list(range(1, top + 1))


@pytest.mark.parametrize("compress", [False, True])
def test_numeric_index_decode_and_range_filter(ps, compress):
    """Numeric index leaves (RS/inverted_index/src/codec/numeric.rs) decoded on the device and filtered by range
    (NumericFilter::value_in_range, reader/numeric.rs:80-85): values bit-equal to the oracle decoder for every value class (tiny,
    +/- integers, f32, f64, infinities), multi-value documents (repeated docIds) yield one hit, and the filtered list intersects
    with a term list like any leaf (the hybrid pre-filter shape)."""
    rng = np.random.default_rng(2024 + compress)
    n = 40_000
    ids = np.cumsum(rng.integers(0, 4, n)) + 1  # steps of 0: multi-value documents
    pool = np.concatenate([rng.integers(0, 8, n // 4).astype(np.float64), rng.integers(-5000, 70_000, n // 4).astype(np.float64),
                           rng.integers(0, 200, n // 4) * 0.125, rng.normal(0, 50, n - 3 * (n // 4))])
    rng.shuffle(pool)
    pool[:6] = [np.inf, -np.inf, -0.0, 2.0**60, -(2.0**40), 1e-9]
    blocks = ol.numeric_blocks(ids.tolist(), pool.tolist(), compress)
    exp_vals = []
    for first, last, cnt, data in blocks:  # what the reference's decoder returns for these bytes
        pos = 0
        for _ in range(cnt):
            used, _, v = ol.numeric_decode(data[pos:])
            pos += used
            exp_vals.append(v)
        assert pos == len(data)
    nl = ps.NumericList(blocks)
    got_ids, got_vals = nl.fetch()
    assert got_ids.tolist() == ids.tolist()
    assert got_vals.tobytes() == np.array(exp_vals, dtype=np.float64).tobytes()
    P = ol.postings()
    term = np.unique(rng.integers(1, int(ids[-1]), 9000)).astype(np.uint64)
    tl = ps.PostingList.from_arrays(term)
    for lo, hi, li, hi_i in ((0.0, 7.0, True, True), (0.0, 7.0, False, False), (-np.inf, np.inf, True, True), (-100.5, 12.125, True, False),
                             (3.0, 3.0, True, True), (5.0, 1.0, True, True), (2.0**60, np.inf, True, False)):
        pl = nl.filter(lo, hi, li, hi_i)
        seen, exp = set(), []
        for d, v in zip(ids.tolist(), exp_vals):
            if d not in seen and P.orc_numeric_in_range(v, lo, hi, int(li), int(hi_i)):
                seen.add(d)
                exp.append(d)
        assert len(pl) == len(exp), (lo, hi, li, hi_i)
        if exp:
            got, _, fr = ps.union([pl]).fetch()
            assert got.tolist() == exp and (fr[0] == 1).all()
            both, _, _ = ps.intersect([pl, tl]).fetch()
            assert both.tolist() == np.intersect1d(np.array(exp, dtype=np.uint64), term).tolist()


def test_union_of_an_expansion_sized_child_set(ps):
    """A prefix / fuzzy expansion is a union of up to MAXEXPANSIONS (200) terms: docIds = the set union, per-child freqs exact,
    BM25STD = the scorer oracle over the present children (the reference runs UnionHeap above 20 children, whose aggregate
    order follows its heap array: the sum may differ from list order in the last bits, hence 1e-13 relative here)."""
    rng = np.random.default_rng(404)
    n_docs, n_children = 200_000, 200
    lists = [np.unique(rng.integers(1, n_docs, int(rng.integers(1, 4000)))).astype(np.uint64) for _ in range(n_children)]
    freqs = [rng.integers(1, 30, len(l)).astype(np.uint32) for l in lists]
    pls = [ps.PostingList.from_arrays(l, f) for l, f in zip(lists, freqs)]
    doc_len = rng.integers(1, 900, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len)
    P = ol.postings()
    terms = [(float(rng.choice([1.0, 0.5, 2.0])), P.orc_idf(n_docs, len(l)), P.orc_idf_bm25(n_docs, len(l))) for l in lists]
    rs = ps.union(pls)
    rs.score(ps.SCORER_BM25STD, terms, 0.9, n_docs, 300.0, dt)
    ids, scores, fr = rs.fetch()
    assert ids.tolist() == np.unique(np.concatenate(lists)).tolist()
    maps = [dict(zip(l.tolist(), f.tolist())) for l, f in zip(lists, freqs)]
    for i in range(0, len(ids), max(1, len(ids) // 150)):
        d = int(ids[i])
        present = [c for c in range(n_children) if d in maps[c]]
        assert [int(fr[c, i]) for c in present] == [maps[c][d] for c in present] and int(fr[:, i].astype(bool).sum()) == len(present)
        s = ol.oracle_score(ol.SCORER_BM25STD, [maps[c][d] for c in present], [terms[c][1] for c in present], [terms[c][2] for c in present],
                            [terms[c][0] for c in present], 0.9, int(doc_len[d]), 1, 1.0, n_docs, 300.0)
        assert abs(s - scores[i]) <= 1e-13 * max(1.0, abs(s)), (d, s, scores[i])
    quick, _, _ = ps.union(pls, quick_exit=True).fetch(want_freqs=False)
    assert quick.tolist() == ids.tolist()
    top_ids, top_scores, total = ps.search_topn(pls, True, ps.SCORER_BM25STD, terms, 0.9, n_docs, 300.0, dt, 10)
    order = np.lexsort((ids, -scores))[:10]
    assert total == len(ids) and top_ids.tolist() == ids[order].tolist()
