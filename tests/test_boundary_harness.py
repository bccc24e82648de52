"""The drop-in boundary of libii_b200.so driven the way RediSearch would drive it, from a C host (tests/cpp/ext_harness.c):
dlopen + RS_ExtensionInit with a capturing RSExtensionCtx (src/extension.c:121-145), term leaves straight from the host's
InvertedIndex blocks (accessors resolved in the host with dlsym), NewIntersectionIterator / NewUnionIterator with the
reference's signatures and ownership rules (RS/headers/iterators_ffi.h:309,594) over B200 leaves, NOT / OPTIONAL wrappers
and a FOREIGN iterator, Read / SkipTo / Rewind / Free through the vtable, and the registered RSScoringFunction per result.
The GPU test compares every printed row with numpy set algebra + the scorer oracle (bit-equal scores)."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "redisearch_b200", "lib", "libii_b200.so")


def _build(tmp_path):
    exe = tmp_path / "ext_harness"
    subprocess.run(["gcc", "-std=gnu11", "-O1", "-Wall", "-rdynamic", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "ext_harness.c"), "-ldl"],
                   check=True)
    return exe


def test_extension_init_registers_the_b200_scorers(tmp_path):
    """CPU: the entry symbol exists and registers seven uniquely named scoring functions (no device work at load time)."""
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), LIB, str(tmp_path), "10", "1.0", "register-only"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip() == ("registered 8: BM25STD.B200 BM25.B200 TFIDF.B200 TFIDF.DOCNORM.B200 DOCSCORE.B200 BM25STD.TANH.B200 "
                                "DISMAX.B200 HAMMING.B200")


@pytest.mark.gpu
def test_constructors_and_scorer_extension_from_a_c_host(tmp_path):
    exe = _build(tmp_path)
    rng = np.random.default_rng(12)
    n_docs = 300_000
    sizes = {"a": 120_000, "b": 60_000, "c": 90_000, "d": 50_000, "e": 40_000}
    lists = {}
    for name, m in sizes.items():
        ids = np.unique(rng.integers(1, n_docs + 1, m)).astype(np.uint32)
        fr = rng.integers(1, 25, len(ids)).astype(np.uint32)
        lists[name] = (ids, fr)
        with open(tmp_path / f"{name}.bin", "wb") as f:
            f.write(np.uint32(len(ids)).tobytes())
            f.write(np.stack([ids, fr], axis=1).astype(np.uint32).tobytes())
    doc_len = rng.integers(50, 500, n_docs + 1).astype(np.uint32)
    doc_len.tofile(tmp_path / "doclen.bin")
    avg = float(doc_len[1:].mean())
    r = subprocess.run([str(exe), LIB, str(tmp_path), str(n_docs), repr(avg)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "HARNESS-OK" in r.stdout, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    lines = r.stdout.splitlines()
    P = ol.postings()

    def lookup(name):
        ids, fr = lists[name]
        return dict(zip(ids.tolist(), fr.tolist()))

    def params(name, weight=1.0, foreign=False):
        n = len(lists[name][0])
        # the host's numeric iterator yields Numeric results: an "irrelevant token" for BM25STD (default.c:296-300 -> contributes 0),
        # weight * freq for TFIDF (:104)
        return (weight, 1.0, 0.0) if foreign else (weight, P.orc_idf(n_docs, n), P.orc_idf_bm25(n_docs, n))

    def expect(required, excluded=(), optional=(), foreign=(), opt_weight=2.0):
        docs = None
        for name in required:
            docs = lists[name][0] if docs is None else np.intersect1d(docs, lists[name][0])
        for name in excluded:
            docs = np.setdiff1d(docs, lists[name][0])
        # aggregate child order: required children ascending by num_estimated (stable), NOT / OPTIONAL children behind
        order = sorted(range(len(required)), key=lambda i: len(lists[required[i]][0]))
        rows = []
        maps = {name: lookup(name) for name in list(required) + list(optional)}
        for d in docs.tolist():
            fr, idf, bidf, w = [], [], [], []
            for i in order:
                name = required[i]
                pw, pi, pb = params(name, foreign=name in foreign)
                fr.append(maps[name][d]); idf.append(pi); bidf.append(pb); w.append(pw)
            for name in optional:
                if d in maps[name]:
                    pw, pi, pb = params(name, opt_weight)
                    fr.append(maps[name][d]); idf.append(pi); bidf.append(pb); w.append(pw)
            s = ol.oracle_score(ol.SCORER_BM25STD, fr, idf, bidf, w, 1.0, int(doc_len[d]), 1, 1.0, n_docs, avg)
            rows.append((d, s, sum(fr)))
        return rows

    def block(variant):
        i = lines.index(next(l for l in lines if l.startswith(f"variant {variant} ")))
        out = []
        for l in lines[i + 1:]:
            if l.startswith("cache"):
                return out, l
            d, s, f = l.split()
            out.append((int(d), float.fromhex(s), int(f)))
        raise AssertionError("no cache line")

    # variant 4: the leaves came from NewInvIndIterator_TermQuery (the reference's signature): same rows as a plain A & B
    for variant, exp in ((0, expect(["a", "b", "c"], foreign=("c",))), (1, expect(["a", "b", "c"], foreign=("c",))),
                         (2, expect(["a", "b"], excluded=("d",))), (3, expect(["a", "b"], optional=("e",))), (4, expect(["a", "b"]))):
        got, cache_line = block(variant)
        assert len(got) == len(exp) and len(exp) > 100, (variant, len(got), len(exp))
        assert [g[0] for g in got] == [e[0] for e in exp], variant
        for g, e in zip(got, exp):
            assert np.float64(g[1]).tobytes() == np.float64(e[1]).tobytes(), (variant, g, e)
            assert g[2] == e[2]
    # variant 5: (a | e) & b with the union NESTED under the intersection: each hit rebuilt as the reference's result tree and scored
    # by the tree oracle (pinned on the reference's default.c recursion in test_oracle_trees.py); BM25STD and TFIDF bits
    i5 = lines.index(next(l for l in lines if l.startswith("variant 5 ")))
    na, ne, nb = (len(lists[x][0]) for x in "aeb")
    assert lines[i5] == f"variant 5 estimated {min(na + ne, nb)}"  # AND: min over the children; OR: their sum
    rows5 = []
    explain_doc, explain_lines, in_explain = None, [], False
    for l in lines[i5 + 1:]:
        if l.startswith("cache"):
            break
        if l.startswith("explain "):
            explain_doc, in_explain = int(l.split()[1]), True
            continue
        if l == "explain-end":
            in_explain = False
            continue
        if in_explain:
            explain_lines.append(l)
            continue
        d, s1, f, s2 = l.split()
        rows5.append((int(d), float.fromhex(s1), int(f), float.fromhex(s2)))
    ma, me, mb = lookup("a"), lookup("e"), lookup("b")
    docs5 = sorted((set(ma) | set(me)) & set(mb))
    assert [r[0] for r in rows5] == docs5 and len(docs5) > 100

    def term_node(name, m, d, weight):
        w, idf, bidf = params(name, weight)
        return {"kind": ol.KIND_TERM, "freq": m[d], "weight": w, "idf": idf, "bm25_idf": bidf}

    for d, s1, f, s2 in rows5:
        un = {"kind": ol.KIND_OR, "weight": 0.5,
              "children": [term_node(x, m, d, w) for x, m, w in (("a", ma, 1.0), ("e", me, 2.0)) if d in m]}
        kids = [un, term_node("b", mb, d, 1.0)]
        if nb < na + ne:  # children ascending by num_estimated * sort weight (stable)
            kids.reverse()
        t = ol.ResultTree({"kind": ol.KIND_AND, "weight": 1.5, "children": kids})
        e1 = t.score(ol.SCORER_BM25STD, int(doc_len[d]), 1, 1.0, n_docs, avg)
        e2 = t.score(ol.SCORER_TFIDF, int(doc_len[d]), 1, 1.0, n_docs, avg)
        assert np.float64(s1).tobytes() == np.float64(e1).tobytes(), (d, s1, e1)
        assert np.float64(s2).tobytes() == np.float64(e2).tobytes(), (d, s2, e2)
        assert f == sum(m[d] for m in (ma, me, mb) if d in m)
        if d == explain_doc:
            # EXPLAINSCORE through TFIDF.B200 (scrExp set): the tree read back from the device, explained with the reference's strings
            want = "\n".join(explain_lines) + "\n"
            if ol.ref_scorers() is not None:
                assert t.ref_explain(ol.SCORER_TFIDF, int(doc_len[d]), 1, 1.0, n_docs, avg, slop=t.min_offset_delta())[1] == want
            assert explain_lines[0].startswith("0 Final TFIDF : words TFIDF ") and explain_lines[1].startswith("1 (Weight 1.50 * total children TFIDF")
            assert sum(1 for l in explain_lines if "= Weight" in l and "TF " in l) == sum(1 for m in (ma, me, mb) if d in m)
    assert explain_doc == docs5[0] and len(explain_lines) >= 4
    # the term cache decoded a and b once: later constructions hit
    assert block(0)[1] == "cache hits 0 misses 2" and block(1)[1] == "cache hits 2 misses 2"
    assert f"union {len(np.union1d(lists['a'][0], lists['c'][0]))}" in lines
