"""II_IndexWriter — the ingest side of posting storage (InvertedIndex::add_record, RS/inverted_index/src/index/core.rs:235-358):
host code, so it is checked without a GPU.  Byte-identity with the oracle restatement (which is pinned on the reference's golden
vectors in test_oracle_postings.py) for every term codec and the numeric codec, block splitting and duplicate rules included."""
import math

import numpy as np
import pytest

import oracle_lib as ol


@pytest.fixture(scope="module")
def ps():
    from redisearch_b200 import postings

    return postings


@pytest.mark.parametrize("codec", range(ol.N_CODECS))
def test_writer_blocks_are_byte_identical_to_the_oracle(ps, codec):
    rng = np.random.default_rng(700 + codec)
    ids = np.cumsum(rng.integers(1, 60, 2600)).astype(np.uint64)
    ids[900:] += np.uint64(1 << 33)  # a delta that does not fit u32: a fresh block (index/core.rs:272-285)
    w, o = ps.IndexWriter(codec), ol.InvIndex(codec)
    grown = []
    for i, d in enumerate(ids.tolist()):
        freq = int(rng.integers(1, 1 << int(rng.integers(1, 31))))
        mask = int(rng.integers(1, 1 << 30)) << (int(rng.integers(0, 98)) if codec in ol.CODECS_WIDE else 0)
        off = bytes(rng.integers(0, 255, int(rng.integers(0, 7)), dtype=np.uint8).tolist()) if codec in ol.CODECS_WITH_OFFSETS else b""
        a, b = w.add(d, freq, mask, off), o.add(d, freq, mask, off)
        assert a == b, (codec, i, a, b)
        grown.append(a)
        if i % 500 == 17:  # a repeated docId is dropped by the term codecs
            assert w.add(d, freq, mask, off) == 0 and o.add(d, freq, mask, off) == 0
    assert w.blocks() == o.blocks() and w.num_docs() == o.num_docs() == len(ids)
    per = 1000 if codec in (ol.CODEC_DOCIDS_ONLY, ol.CODEC_RAW_DOCIDS_ONLY) else 100
    assert all(b[2] <= per for b in w.blocks()) and len(w.blocks()) >= len(ids) // per


def test_writer_reproduces_the_reference_golden_bytes(ps):
    """the first record of a block has delta 0; a second record carries the golden delta (tests/golden/postings_golden.json)"""
    import json
    import os

    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "postings_golden.json")))
    base = 1 << 32

    def single(codec, delta, freq, mask, offs=b""):
        w = ps.IndexWriter(codec)
        if delta == 0:
            w.add(base, freq, mask, offs)
            return list(w.blocks()[0][3])
        first = w.add(base, 7, 1, b"\1")
        w.add(base + delta, freq, mask, offs)
        return list(w.blocks()[0][3][first:])

    for delta, freq, mask, offs, expected in G["full"]:
        assert single(ol.CODEC_FULL, delta, freq, mask, bytes(offs)) == expected
    for delta, freq, offs, expected in G["freqs_offsets"]["cases"]:
        assert single(ol.CODEC_FREQS_OFFSETS, delta, freq, 1, bytes(offs)) == expected
    for delta, mask, offs, expected in G["fields_offsets_wide"]["cases"]:
        assert single(ol.CODEC_FIELDS_OFFSETS_WIDE, delta, 1, int(mask), bytes(offs)) == expected
    for delta, freq, mask, offs, expected in G["full_wide"]["cases"]:
        assert single(ol.CODEC_FULL_WIDE, delta, freq, int(mask), bytes(offs)) == expected
    for delta, mask, expected in G["fields_only_wide"]["cases"]:
        assert single(ol.CODEC_FIELDS_ONLY_WIDE, delta, 1, int(mask)) == expected
    for delta, expected in G["doc_ids_only"]:
        assert single(ol.CODEC_DOCIDS_ONLY, delta, 1, 1) == expected


@pytest.mark.parametrize("compress", [False, True])
def test_numeric_writer_matches_the_oracle_and_keeps_a_document_together(ps, compress):
    rng = np.random.default_rng(801 + compress)
    n = 5000
    ids = np.cumsum(rng.integers(0, 3, n)) + 1  # steps of 0: multi-value documents (ALLOW_DUPLICATES, numeric.rs:323)
    vals = np.concatenate([rng.integers(0, 8, n // 4).astype(np.float64), rng.integers(-5000, 70_000, n // 4).astype(np.float64),
                           rng.integers(0, 200, n // 4) * 0.125, rng.normal(0, 50, n - 3 * (n // 4))])
    rng.shuffle(vals)
    vals[:8] = [math.inf, -math.inf, -0.0, 2.0**60, -(2.0**40), 1e-9, 3.124, 100.500001]
    w = ps.IndexWriter(numeric=True, compress_floats=compress)
    for d, v in zip(ids.tolist(), vals.tolist()):
        assert w.add_numeric(d, v) > 0
    exp = ol.numeric_blocks(ids.tolist(), vals.tolist(), compress)
    got = w.blocks()
    assert got == exp
    assert w.num_docs() == len(np.unique(ids))
    assert any(b[2] > 100 for b in got) or all(ids[i] != ids[i - 1] for i in range(100, n, 100))  # a full block grows for the same document
    # a delta above 7 bytes opens a block
    w2 = ps.IndexWriter(numeric=True)
    w2.add_numeric(5, 1.0)
    w2.add_numeric(5 + (1 << 56), 2.0)
    assert [b[:3] for b in w2.blocks()] == [(5, 5, 1), (5 + (1 << 56), 5 + (1 << 56), 1)]
