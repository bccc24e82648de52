"""CPU-only checks of the C-ABI boundary: the library loads without a GPU, exports every symbol
include/vecsim_b200.h declares, and the header's struct layouts / enum values are identical to the
reference's (golden table taken from the reference headers)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b((?:VecSim|II_|RS_)\w*)\s*\(", src)
    # function-pointer typedefs and macros are not exports
    return sorted({n for n in names if not n.endswith("_t")})


def test_library_loads_and_exports_every_declared_symbol():
    from redisearch_b200 import vecsim

    L = vecsim.lib()
    declared = _declared_functions("vecsim_b200.h")
    assert len(declared) > 50
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    bound = {s[0] for s in vecsim.SIGNATURES} | set(vecsim.EXTRA_SYMBOLS)
    assert set(declared) <= bound, sorted(set(declared) - bound)
    assert b"sm_100a" in L.VecSimB200_Version()


def test_no_device_means_null_index_not_a_cpu_fallback():
    """Without a CUDA device VecSimIndex_New must fail (NULL), never fall back to host compute."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from redisearch_b200 import vecsim

    with pytest.raises(RuntimeError):
        vecsim.VecSimIndex(vecsim.VecSimType_FLOAT32, 8, vecsim.VecSimMetric_L2)


def test_blob_helpers_run_on_host():
    """VecSim_Normalize / GetQueryBlobSize are host arithmetic (normalize_naive.h) — usable without a GPU."""
    import numpy as np

    import oracle_lib as ol
    from redisearch_b200 import vecsim

    L = vecsim.lib()
    assert L.VecSimParams_GetQueryBlobSize(vecsim.VecSimType_FLOAT32, 128, vecsim.VecSimMetric_Cosine) == 512
    assert L.VecSimParams_GetQueryBlobSize(vecsim.VecSimType_INT8, 100, vecsim.VecSimMetric_Cosine) == 104
    assert L.VecSimParams_GetQueryBlobSize(vecsim.VecSimType_UINT8, 100, vecsim.VecSimMetric_L2) == 100
    assert L.VecSimParams_GetQueryBlobSize(vecsim.VecSimType_FLOAT16, 100, vecsim.VecSimMetric_IP) == 200
    rng = np.random.default_rng(3)
    for vtype in (ol.F32, ol.F16, ol.BF16, ol.I8, ol.U8):
        for dim in (3, 16, 129, 768):
            x = rng.uniform(-1, 1, dim).astype(np.float32)
            blob = ol.to_type(x, vtype)
            extra = 4 if vtype in (ol.I8, ol.U8) else 0
            a = np.zeros(blob.nbytes + extra, dtype=np.uint8)
            a[: blob.nbytes] = blob.view(np.uint8)
            b = a.copy()
            vecsim.normalize(a, dim, vtype)
            ol.port().orc_normalize(ol._p(b), dim, vtype)
            assert a.tobytes() == b.tobytes(), (vtype, dim)


def test_header_abi_layout_matches_reference_golden(tmp_path):
    """sizeof/offsetof/enum table of include/vecsim_b200.h == tests/golden/vecsim_abi_layout.txt
    (generated from the reference's vec_sim.h by tests/golden/make_fixtures.py)."""
    exe = tmp_path / "abi_probe"
    hdr = os.path.join(ROOT, "include", "vecsim_b200.h")
    subprocess.run(["gcc", f'-DHDR="{hdr}"', os.path.join(ROOT, "tests", "abi", "abi_probe.c"), "-o", str(exe)], check=True)
    mine = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    golden = open(os.path.join(ROOT, "tests", "golden", "vecsim_abi_layout.txt")).read()
    assert mine == golden
    ref_hdr = "/root/reference/deps/VectorSimilarity/src/VecSim/vec_sim.h"
    if os.path.exists(ref_hdr):  # in the build container also re-derive the golden from the reference itself
        exe2 = tmp_path / "abi_probe_ref"
        subprocess.run(["gcc", '-DHDR="VecSim/vec_sim.h"', "-I/root/reference/deps/VectorSimilarity/src",
                        os.path.join(ROOT, "tests", "abi", "abi_probe.c"), "-o", str(exe2)], check=True)
        assert subprocess.run([str(exe2)], check=True, capture_output=True, text=True).stdout == golden


def test_ii_library_loads_and_exports_every_declared_symbol():
    from redisearch_b200 import postings

    L = postings.lib()
    declared = [n for n in _declared_functions("ii_b200.h") if n.startswith("II_") and n[3].isupper() and "_" in n[3:] or n in ("II_Intersect", "II_Union", "II_Score", "II_Version", "II_GetStats", "II_SearchTopN", "II_NewResultIterator", "II_CalculateIDF", "II_CalculateIDF_BM25")]
    declared = [n for n in declared if not n.startswith("II_IteratorType") and not n.startswith("II_ResultData") and not n.startswith("II_CODEC") and not n.startswith("II_SCORER")]
    assert len(declared) >= 25, declared
    missing = [n for n in declared if not hasattr(L, n)]
    assert not missing, missing
    bound = {s[0] for s in postings.SIGNATURES}
    assert set(declared) <= bound, sorted(set(declared) - bound)
    assert L.II_CalculateIDF(100, 10) == 3.0 and abs(L.II_CalculateIDF_BM25(100, 10) - 2.2635) < 1e-3


def test_ii_header_abi_layout_matches_reference_golden(tmp_path):
    exe = tmp_path / "ii_probe"
    hdr = os.path.join(ROOT, "include", "ii_b200.h")
    subprocess.run(["gcc", f'-DHDR="{hdr}"', os.path.join(ROOT, "tests", "abi", "ii_abi_probe.c"), "-o", str(exe)], check=True)
    mine = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    assert mine == open(os.path.join(ROOT, "tests", "golden", "ii_abi_layout.txt")).read()


def test_merge_shard_topn_is_host_code_with_cmpbyscore_order():
    """II_MergeShardTopN is the coordinator-side reducer (no device needed): score desc, docId asc on ties
    (cmpByScore, src/result_processor.c:834-850), padding beyond counts[g] ignored."""
    import numpy as np

    from redisearch_b200 import postings

    L = postings.lib()
    G, per, n = 3, 4, 6
    scores = np.array([[5.0, 3.0, 3.0, 99.0], [5.0, 4.0, 0.0, 0.0], [3.0, 2.5, 1.0, 0.5]], dtype=np.float64)
    ids = np.array([[10, 7, 9, 1], [4, 20, 0, 0], [8, 30, 31, 32]], dtype=np.uint64)
    counts = np.array([3, 2, 4], dtype=np.uint64)  # shard 0's 4th and shard 1's 3rd/4th entries are padding
    out_i = np.zeros(n, dtype=np.uint64)
    out_s = np.zeros(n, dtype=np.float64)
    got = L.II_MergeShardTopN(scores.ctypes.data, ids.ctypes.data, counts.ctypes.data, G, per, n, out_i.ctypes.data, out_s.ctypes.data)
    assert got == 6
    assert out_i.tolist() == [4, 10, 20, 7, 8, 9] and out_s.tolist() == [5.0, 5.0, 4.0, 3.0, 3.0, 3.0]
    got = L.II_MergeShardTopN(scores.ctypes.data, ids.ctypes.data, counts.ctypes.data, G, per, 50, np.zeros(50, dtype=np.uint64).ctypes.data,
                              np.zeros(50).ctypes.data)
    assert got == 9


def test_wildcard_iterator_contract_known_answers():
    """rqe_iterators/tests/integration/wildcard.rs (initial_state, read_sequential, skip_to_valid_targets, skip_to_beyond_range,
    rewind): the wildcard iterator is host code (a counter), so its QueryIterator contract is checked without a GPU."""
    from redisearch_b200 import postings as ps

    L = ps.lib()
    it = L.NewWildcardIterator_NonOptimized(10, 5.0)
    q = it.contents
    assert q.type == 12 and q.lastDocId == 0 and not q.atEOF and q.NumEstimated(it) == 10  # initial_state
    q.Free(it)
    it = L.II_NewWildcardIterator(5, 0.5)
    q = it.contents
    for expected in range(1, 6):  # read_sequential
        assert q.Read(it) == ps.ITERATOR_OK and q.lastDocId == expected and not q.atEOF
        cur = q.current.contents
        assert cur.docId == expected and cur.weight == 0.5 and cur.freq == 1 and cur.data.tag == 8  # a virtual result
    assert q.Read(it) == ps.ITERATOR_EOF and q.atEOF and not q.current
    assert q.Read(it) == ps.ITERATOR_EOF
    q.Rewind(it)
    assert q.lastDocId == 0 and not q.atEOF and q.Read(it) == ps.ITERATOR_OK and q.lastDocId == 1
    q.Free(it)
    it = L.II_NewWildcardIterator(10, 5.0)
    q = it.contents
    assert q.SkipTo(it, 5) == ps.ITERATOR_OK and q.lastDocId == 5 and not q.atEOF  # skip_to_valid_targets
    assert q.SkipTo(it, 10) == ps.ITERATOR_OK and q.lastDocId == 10 and not q.atEOF
    assert q.Read(it) == ps.ITERATOR_EOF and q.atEOF and not q.current
    q.Rewind(it)
    assert q.SkipTo(it, 3) == ps.ITERATOR_OK
    assert q.SkipTo(it, 11) == ps.ITERATOR_EOF and q.atEOF and q.lastDocId == 3  # beyond the range: the position stays
    assert q.SkipTo(it, 4) == ps.ITERATOR_EOF
    q.Free(it)


def test_index_flags_to_codec_table():
    """II_CodecFromIndexFlags == the match of NewInvertedIndex_Ex (RS/c_entrypoint/inverted_index_ffi/src/lib.rs:49-165) over the
    storage flags of src/spec.h:171-181."""
    from redisearch_b200 import postings as ps

    L = ps.lib()
    OFF, FLD, FRQ, NUM, BYTEOFF, WIDE = 0x01, 0x02, 0x10, 0x20, 0x40, 0x80
    table = {FRQ | OFF | FLD: ps.CODEC_FULL, FRQ | OFF | FLD | WIDE: 9, FRQ | FLD: ps.CODEC_FREQS_FIELDS, FRQ | FLD | WIDE: 10,
             FRQ: ps.CODEC_FREQS_ONLY, FLD: ps.CODEC_FIELDS_ONLY, FLD | WIDE: 11, FLD | OFF: 8, FLD | OFF | WIDE: 12, OFF: 7, FRQ | OFF: 6}
    for flags, codec in table.items():
        assert L.II_CodecFromIndexFlags(flags, 0) == codec and L.II_CodecFromIndexFlags(flags | BYTEOFF | 0x100, 1) == codec
    assert L.II_CodecFromIndexFlags(0, 0) == ps.CODEC_DOCIDS_ONLY and L.II_CodecFromIndexFlags(0, 1) == ps.CODEC_RAW_DOCIDS_ONLY
    assert L.II_CodecFromIndexFlags(NUM, 0) == -1 and L.II_CodecFromIndexFlags(FRQ | WIDE, 0) == -1
