"""NESTED aggregates on the device — `(a|b) c`, `(a b)|c`, `((a|b) c)|d`, `(a|b) (c|d)`: an evaluated AND / OR as ONE child of
another aggregate (II_ResultSet_IntoChild).  What the reference does with such a result tree:
  * the scorers recurse (src/ext/default.c tfidfRecursive / bm25Recursive / bm25StdRecursive / dismaxRecursive),
  * GetSlop and the proximity check merge the children's term positions (src/offset_vector.c, proximity.rs OffsetIter::Merge),
is restated in oracle/tree_oracle.c and pinned on the reference's own compiled code in tests/test_oracle_trees.py; here every hit of
the device's answer is rebuilt as a tree and compared: docIds equal, scores BIT-equal, phrase survivors equal."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as ol
from oracle_lib import KIND_AND, KIND_OR, KIND_TERM, ResultTree

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ps():
    from redisearch_b200 import postings

    return postings


# ---- a tiny query algebra evaluated twice: on the device, and document by document into oracle trees ------------------
class T:  # term leaf i
    def __init__(self, i):
        self.i = i


class A:  # AND
    def __init__(self, kids, weight=1.0):
        self.kids, self.weight = kids, weight


class O:  # OR
    def __init__(self, kids, weight=1.0):
        self.kids, self.weight = kids, weight


class Corpus:
    def __init__(self, ps, rng, n_docs, densities, doc_words=40, codec=ol.CODEC_FULL):
        self.ps, self.n_docs = ps, n_docs
        self.idx, self.offs, self.freq = [], [], []
        for dens in densities:
            ix = ol.InvIndex(codec)
            docs = (np.flatnonzero(rng.random(n_docs - 1) < dens) + 1).tolist() + [n_docs]  # every term ends on the last document:
            m, fr = {}, {}                                                                  # no union child is exhausted early
            for d in docs:
                k = int(rng.integers(1, 6))
                pos = np.unique(rng.integers(1, doc_words, k))
                if rng.random() < 0.03:
                    pos = pos[:0]
                ob = ol.varint_deltas(pos.tolist())
                f = max(1, len(pos))
                m[d], fr[d] = (pos.tolist(), ob), f
                ix.add(d, f, 1, ob)
            self.idx.append(ix)
            self.offs.append(m)
            self.freq.append(fr)
        self.pls = ps.postings_with_offsets([ix.blocks() for ix in self.idx], codec)
        P = ol.postings()
        w = rng.choice([1.0, 0.5, 2.0], len(densities)).tolist()
        self.terms = [(w[i], P.orc_idf(n_docs, ix.num_docs()), P.orc_idf_bm25(n_docs, ix.num_docs())) for i, ix in enumerate(self.idx)]

    # --- host model -------------------------------------------------------------------------------------------------------
    def docs(self, q):
        if isinstance(q, T):
            return set(self.freq[q.i])
        sets = [self.docs(k) for k in q.kids]
        return set.intersection(*sets) if isinstance(q, A) else set.union(*sets)

    def est(self, q):  # num_estimated: leaf = unique docs, AND = min, OR = sum (intersection.rs:146, union_flat.rs:102)
        if isinstance(q, T):
            return len(self.freq[q.i])
        e = [self.est(k) for k in q.kids]
        return min(e) if isinstance(q, A) else sum(e)

    def sort_weight(self, q):  # intersection_sort_weight (intersection.rs:580, union_flat.rs:817 with the default configuration)
        return 1.0 / len(q.kids) if isinstance(q, A) else 1.0

    def ordered_kids(self, q, in_order=False):
        if isinstance(q, A) and not in_order:
            return sorted(q.kids, key=lambda k: self.est(k) * self.sort_weight(k))  # stable
        return list(q.kids)

    def tree(self, q, d, in_order=False):
        """the reference's result for document d as a nested dict (None: d does not match q)"""
        if isinstance(q, T):
            if d not in self.freq[q.i]:
                return None
            w, idf, bidf = self.terms[q.i]
            return {"kind": KIND_TERM, "freq": self.freq[q.i][d], "weight": w, "idf": idf, "bm25_idf": bidf, "positions": self.offs[q.i][d][0]}
        kids = [self.tree(k, d) for k in self.ordered_kids(q, in_order)]
        if isinstance(q, A):
            if any(k is None for k in kids):
                return None
            return {"kind": KIND_AND, "weight": q.weight, "children": kids}
        kids = [k for k in kids if k is not None]
        return {"kind": KIND_OR, "weight": q.weight, "children": kids} if kids else None

    # --- device -----------------------------------------------------------------------------------------------------------
    def child_list(self, q, with_positions=False):
        """(list view, term params entry) of q as a child"""
        if isinstance(q, T):
            return self.pls[q.i], self.terms[q.i]
        rs, terms = self.evaluate(q)
        return rs.into_child(terms, q.weight, with_positions), (q.weight, 1.0, 1.0)

    def evaluate(self, q, phrase=None):
        views = [self.child_list(k, with_positions=phrase is not None) for k in q.kids]
        lists, terms = [v[0] for v in views], [v[1] for v in views]
        if isinstance(q, O):
            return self.ps.union(lists), terms
        if phrase is not None:
            return self.ps.intersect_phrase(lists, phrase[0], phrase[1]), terms
        return self.ps.intersect(lists), terms


SHAPES = {
    "(a|b) c": lambda: A([O([T(0), T(1)], 0.8), T(2)], 0.7),
    "(a b)|c": lambda: O([A([T(0), T(1)], 1.5), T(2)], 0.9),
    "((a|b) c)|d": lambda: O([A([O([T(0), T(1)], 0.5), T(2)], 2.0), T(3)], 1.1),
    "(a|b) (c|d)": lambda: A([O([T(0), T(1)]), O([T(2), T(3)], 0.6)], 1.0),
    "a (b|c) d": lambda: A([T(0), O([T(1), T(2)], 1.3), T(3)], 0.4),
}


@pytest.mark.parametrize("scorer", [ol.SCORER_BM25STD, ol.SCORER_BM25, ol.SCORER_TFIDF, ol.SCORER_TFIDF_DOCNORM, ol.SCORER_BM25STD_TANH,
                                    ol.SCORER_DISMAX, ol.SCORER_DOCSCORE])
@pytest.mark.parametrize("shape", list(SHAPES))
def test_nested_aggregates_score_like_the_reference_recursion(ps, shape, scorer):
    rng = np.random.default_rng(2000 + 17 * list(SHAPES).index(shape))
    n_docs = 30_000
    cp = Corpus(ps, rng, n_docs, [0.30, 0.22, 0.45, 0.12])
    q = SHAPES[shape]()
    doc_len = rng.integers(1, 900, n_docs + 1).astype(np.uint32)
    doc_score = rng.choice(np.array([1.0, 0.5, 0.77], dtype=np.float32), n_docs + 1)
    max_freq = rng.integers(1, 60, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len, doc_score, max_freq)
    avg = 222.5
    rs, terms = cp.evaluate(q)
    rs.score(scorer, terms, q.weight, n_docs, avg, dt, 0.0, 4)
    ids, scores, _ = rs.fetch()
    exp_docs = sorted(cp.docs(q))
    assert ids.tolist() == exp_docs and len(exp_docs) > 500
    slops = set()
    step = max(1, len(exp_docs) // 300)
    for i in sorted(set(range(0, len(exp_docs), step)) | set(range(max(0, len(exp_docs) - 50), len(exp_docs)))):
        d = exp_docs[i]
        t = ResultTree(cp.tree(q, d))
        slops.add(t.min_offset_delta())
        s = t.score(scorer, int(doc_len[d]), int(max_freq[d]), float(doc_score[d]), n_docs, avg, slop=-1, min_score=0.0, tanh_factor=4.0)
        if scorer == ol.SCORER_BM25STD_TANH:  # libm tanh on the host, the device's on the GPU
            assert abs(s - scores[i]) <= 1e-12, (shape, d, s, scores[i])
        else:
            assert np.float64(s).tobytes() == np.float64(scores[i]).tobytes(), (shape, scorer, d, s, scores[i])
    assert len(slops) >= 2, slops


@pytest.mark.parametrize("in_order", [False, True])
@pytest.mark.parametrize("shape", ["(a|b) c", "a (b|c) d", "(a|b) (c|d)"])
def test_phrase_over_nested_unions(ps, shape, in_order):
    """slop / in-order over children that are unions of expansions: the merged positions of the union take part like a term's"""
    rng = np.random.default_rng(3100 + len(shape) + in_order)
    n_docs = 30_000
    cp = Corpus(ps, rng, n_docs, [0.35, 0.3, 0.5, 0.4], doc_words=30)
    q = SHAPES[shape]()
    cand = sorted(cp.docs(q))
    seen = set()
    for slop in (0, 2, 6, None):
        if slop is None and not in_order:
            continue
        rs, _ = cp.evaluate(q, phrase=(slop, in_order))
        got = rs.fetch()[0].tolist()
        exp = [d for d in cand if ResultTree(cp.tree(q, d, in_order)).within_range(slop, in_order)]
        assert got == exp, (shape, slop, in_order, len(got), len(exp))
        seen.add(0 < len(exp) < len(cand))
    assert True in seen


def test_union_of_terms_without_positions_counts_as_having_offsets(ps):
    """has-offsets of an aggregate goes by the kind mask of its children (index_result.c:23-35): a union of terms that were indexed
    WITHOUT positions still takes part in the proximity check and, yielding nothing, rejects every document; the same terms
    directly under the AND are skipped instead"""
    rng = np.random.default_rng(5)
    n_docs = 5_000
    with_pos = Corpus(ps, rng, n_docs, [0.6, 0.6], doc_words=12)
    ids = [np.unique(rng.integers(1, n_docs, 3000)).astype(np.uint64) for _ in range(2)]
    plain = [ps.PostingList.from_arrays(x, np.ones(len(x), dtype=np.uint32)) for x in ids]
    inner = ps.union(plain).into_child([(1.0, 1.0, 1.0)] * 2, 1.0, with_positions=True)
    rs = ps.intersect_phrase([inner, with_pos.pls[0], with_pos.pls[1]], 100, False)
    assert len(rs) == 0
    flat = ps.intersect_phrase([with_pos.pls[0], with_pos.pls[1]], 100, False)
    assert len(flat) > 100


def test_nested_constructors_stay_on_the_device(ps):
    """NewUnionIterator inside NewIntersectionIterator: the nested node's result set moves into the parent (no host round trip of
    its hits), NumEstimated follows the reference's rule (AND: min over the children, OR: their sum), and a phrase constraint over a
    nested union is evaluated on the device."""
    rng = np.random.default_rng(77)
    n_docs = 20_000
    cp = Corpus(ps, rng, n_docs, [0.3, 0.25, 0.5])
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

    def leaf(i):
        w, idf, bidf = cp.terms[i]
        return L.II_NewTermIterator(cp.pls[i].h, 0, w, idf, bidf)

    def drain(qi):
        got = []
        while qi.contents.Read(qi) == 0:
            got.append(qi.contents.lastDocId)
        qi.contents.Free(qi)
        return got

    q = SHAPES["(a|b) c"]()
    un = L.NewUnionIterator(its_of([leaf(0), leaf(1)]), 2, False, 0.8, 0, None, None)
    assert un.contents.NumEstimated(un) == cp.est(q.kids[0])
    qi = L.NewIntersectionIterator(its_of([un, leaf(2)]), 2, -1, False, 0.7)
    assert qi.contents.NumEstimated(qi) == cp.est(q)
    assert drain(qi) == sorted(cp.docs(q))
    for slop, in_order in ((1, True), (3, False)):
        un = L.NewUnionIterator(its_of([leaf(0), leaf(1)]), 2, False, 0.8, 0, None, None)
        qi = L.NewIntersectionIterator(its_of([un, leaf(2)]), 2, slop, in_order, 0.7)
        assert qi
        exp = [d for d in sorted(cp.docs(q)) if ResultTree(cp.tree(q, d, in_order)).within_range(slop, in_order)]
        got = drain(qi)
        assert got == exp and 0 < len(exp) < len(cp.docs(q))


@pytest.mark.parametrize("scorer", [ol.SCORER_BM25STD, ol.SCORER_TFIDF, ol.SCORER_BM25])
def test_optional_and_not_over_nested_sets(ps, scorer):
    """`a ~(b c) -(d|e)`: OPTIONAL over a nested AND contributes the set's recursive score where it matches (with the OPTIONAL's
    weight as the aggregate's own, optional.rs:260) and a virtual result elsewhere; NOT over a nested OR excludes its documents."""
    rng = np.random.default_rng(4200 + scorer)
    n_docs = 30_000
    cp = Corpus(ps, rng, n_docs, [0.5, 0.45, 0.5, 0.1, 0.08])
    doc_len = rng.integers(1, 900, n_docs + 1).astype(np.uint32)
    doc_score = rng.choice(np.array([1.0, 0.5], dtype=np.float32), n_docs + 1)
    max_freq = rng.integers(1, 60, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len, doc_score, max_freq)
    avg, opt_w, aggw = 180.0, 2.5, 0.9
    q_opt, q_not = A([T(1), T(2)]), O([T(3), T(4)])
    opt_rs, opt_terms = cp.evaluate(q_opt)
    not_rs, not_terms = cp.evaluate(q_not)
    lists = [cp.pls[0], opt_rs.into_child(opt_terms, opt_w), not_rs.into_child(not_terms, 1.0)]
    rs = ps.intersect_ex(lists, [0, 2, 1])
    terms = [cp.terms[0], (opt_w, 1.0, 1.0), (0.0, 1.0, 1.0)]
    rs.score(scorer, terms, aggw, n_docs, avg, dt, 0.0, 4)
    ids, scores, fr = rs.fetch()
    exp_docs = sorted(cp.docs(T(0)) - cp.docs(q_not))
    assert ids.tolist() == exp_docs and len(exp_docs) > 1000
    assert rs.child_order().tolist() == [0, 1, 2]  # the required child first, NOT / OPTIONAL behind in their given order
    virtual = {"kind": ol.KIND_VIRTUAL, "freq": 0, "weight": 0.0}
    n_opt = 0
    for i in range(0, len(exp_docs), max(1, len(exp_docs) // 300)):
        d = exp_docs[i]
        opt = cp.tree(A(q_opt.kids, opt_w), d)
        n_opt += opt is not None
        t = ResultTree({"kind": KIND_AND, "weight": aggw, "children": [cp.tree(T(0), d), opt if opt else virtual, virtual]})
        s = t.score(scorer, int(doc_len[d]), int(max_freq[d]), float(doc_score[d]), n_docs, avg, slop=-1)
        assert np.float64(s).tobytes() == np.float64(scores[i]).tobytes(), (scorer, d, s, scores[i])
        assert fr[:, i].tolist() == [cp.freq[0][d], (cp.freq[1][d] + cp.freq[2][d]) if opt else 0, 0]
    assert 20 < n_opt < 280
    # the same through the constructors: NOT / OPTIONAL wrap the nested nodes
    L = ps.lib()
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]

    def its_of(children):
        arr = libc.malloc(8 * len(children))
        view = (C.c_void_p * len(children)).from_address(arr)
        for k, c in enumerate(children):
            view[k] = C.cast(c, C.c_void_p).value
        return arr

    def leaf(i):
        w, idf, bidf = cp.terms[i]
        return L.II_NewTermIterator(cp.pls[i].h, 0, w, idf, bidf)

    inner_and = L.NewIntersectionIterator(its_of([leaf(1), leaf(2)]), 2, -1, False, 1.0)
    inner_or = L.NewUnionIterator(its_of([leaf(3), leaf(4)]), 2, False, 1.0, 0, None, None)
    opt_it = L.II_NewOptionalIterator(inner_and, n_docs, opt_w)
    not_it = L.II_NewNotIterator(inner_or, n_docs, 1.0)
    assert opt_it and not_it
    qi = L.NewIntersectionIterator(its_of([leaf(0), opt_it, not_it]), 3, -1, False, aggw)
    assert qi
    got, freqs = [], []
    while qi.contents.Read(qi) == 0:
        got.append(qi.contents.lastDocId)
        freqs.append(qi.contents.current.contents.freq)
    qi.contents.Free(qi)
    assert got == exp_docs
    both = cp.docs(q_opt)
    assert freqs == [cp.freq[0][d] + ((cp.freq[1][d] + cp.freq[2][d]) if d in both else 0) for d in exp_docs]


@pytest.mark.parametrize("scorer", [ol.SCORER_BM25STD, ol.SCORER_TFIDF])
def test_search_topn_over_a_nested_child(ps, scorer):
    """II_SearchTopN / II_SearchTopNBatch (AND -> scorer -> top-N in one call) with a nested set among the lists: the per-query chain
    recurses into it; same rows as scoring the result set and sorting on the host (score desc, docId asc: RPSorter's order)."""
    rng = np.random.default_rng(5100 + scorer)
    n_docs = 30_000
    cp = Corpus(ps, rng, n_docs, [0.3, 0.25, 0.5])
    doc_len = rng.integers(1, 900, n_docs + 1).astype(np.uint32)
    dt = ps.DocTable(n_docs, doc_len, None, rng.integers(1, 60, n_docs + 1).astype(np.uint32))
    q = SHAPES["(a|b) c"]()
    avg = 200.0
    rs, terms = cp.evaluate(q)
    rs.score(scorer, terms, q.weight, n_docs, avg, dt, 0.0, 4)
    ids, scores, _ = rs.fetch()
    order = np.lexsort((ids, -scores))[:10]
    views = [cp.child_list(k) for k in q.kids]
    got_ids, got_scores, total = ps.search_topn([v[0] for v in views], False, scorer, [v[1] for v in views], q.weight, n_docs, avg, dt, 10)
    assert total == len(ids)
    assert got_ids.tolist() == ids[order].tolist()
    assert got_scores.tobytes() == scores[order].tobytes()
    views2 = [cp.child_list(k) for k in q.kids]
    batch = ps.SearchBatch([([v[0] for v in views2], [v[1] for v in views2])], 10)
    b_ids, b_scores, b_total = batch.run(False, scorer, q.weight, n_docs, avg, dt)[0]
    assert b_total == len(ids) and b_ids.tolist() == ids[order].tolist() and b_scores.tobytes() == scores[order].tobytes()
