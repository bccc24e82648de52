/* oracle.h — CPU restatement (plain C) of the reference's query-time scoring hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked, loaded or executed by the
 * product (redisearch_b200/, include/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * Parity pinning:
 *   - vecsim_oracle.c is checked against oracle/_ref/libvecsim_ref.so — the reference's own VecSim
 *     sources compiled in place — by tests/test_oracle_vecsim.py (bit-equal on AVX-512F hosts for
 *     the fp32 tier it restates, exact for int8/uint8, tolerance for fp16/bf16), and against the
 *     known answers of deps/VectorSimilarity/tests/unit/test_bruteforce.cpp / test_spaces.cpp.
 *   - postings_oracle.c follows the Rust sources (no Rust toolchain here) and is pinned by the
 *     golden byte vectors / known answers of the reference's own tests, transcribed under
 *     tests/golden/ (tests/test_oracle_postings.py).
 *   - scorer_oracle.c is checked against oracle/_ref/libscorers_ref.so (the reference's
 *     src/ext/default.c) and tests/pytests/test_scorers.py golden scores.
 */
#ifndef RS_ORACLE_H
#define RS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as VecSimType / VecSimMetric (VS/vec_sim_common.h:60-87). */
enum { ORC_F32 = 0, ORC_BF16 = 2, ORC_F16 = 3, ORC_I8 = 4, ORC_U8 = 5 };
enum { ORC_L2 = 0, ORC_IP = 1, ORC_COS = 2 };
/* Which reference tier the fp32 arithmetic follows:
 *   ORC_TIER_SCALAR  VS/spaces/L2/L2.cpp:76-86, IP/IP.cpp:185-194 (sequential, unfused)
 *   ORC_TIER_AVX512  VS/spaces/L2/L2_AVX512F_FP32.h:21-59, IP/IP_AVX512F_FP32.h:19-56 emulated
 *                    lane by lane (what an AVX-512F host dispatches to for dim >= 8,
 *                    L2_space.cpp:213-223) */
enum { ORC_TIER_SCALAR = 0, ORC_TIER_AVX512 = 1 };

/* ---- vecsim_oracle.c ------------------------------------------------------------------------ */
float orc_distance(int type, int metric, size_t dim, const void *a, const void *b, int tier);
void orc_normalize(void *blob, size_t dim, int type);       /* normalize_naive.h:23-88 */
size_t orc_stored_size(int type, size_t dim, int metric);   /* vec_utils.cpp:296-302 */
float orc_half_to_float(uint16_t h);                        /* types/float16.h:33-52 */
uint16_t orc_float_to_half(float f);                        /* types/float16.h:62-117 */
uint16_t orc_float_to_bf16(float f);                        /* types/bfloat16.h:22-29 */

typedef struct OrcIndex OrcIndex;
OrcIndex *orc_index_new(int type, size_t dim, int metric, int multi, int tier);
void orc_index_free(OrcIndex *ix);
int orc_index_add(OrcIndex *ix, const void *blob, size_t label);      /* brute_force_single.h:135-148 */
void orc_index_add_bulk(OrcIndex *ix, const void *blobs, size_t stride, size_t n, size_t label0);
int orc_index_delete(OrcIndex *ix, size_t label);                     /* brute_force.h:196-224 */
size_t orc_index_size(const OrcIndex *ix);
/* brute_force.h:243-291; order 0 = BY_SCORE, 1 = BY_ID.  Returns the number of results. */
size_t orc_index_topk(const OrcIndex *ix, const void *query, size_t k, int order, size_t *labels, double *scores);
size_t orc_index_range(const OrcIndex *ix, const void *query, double radius, int order, size_t cap,
                       size_t *labels, double *scores);               /* brute_force.h:293-326 */
double orc_index_distance_from(const OrcIndex *ix, size_t label, const void *stored_form_query);
int orc_index_prefer_adhoc(const OrcIndex *ix, size_t subset, size_t k, int initial); /* brute_force.h:380-451 */
/* All (score,label) pairs of one query in batch-iterator order ((score,label) ascending, best score
 * per label for multi): what successive VecSimBatchIterator_Next calls hand out
 * (bf_batch_iterator.h:59-200).  Returns the number of pairs (= label count). */
size_t orc_index_all_sorted(const OrcIndex *ix, const void *query, size_t *labels, double *scores);
/* Wall seconds for nq top-k queries spread over nthreads threads (one query = one thread, like the
 * reference). */
double orc_index_time_topk(const OrcIndex *ix, const void *queries, size_t qstride, size_t nq, size_t k,
                           int nthreads, size_t *labels, double *scores);
/* streaming top-k over a flat chunk of stored-form rows (state carried between chunks); see vecsim_oracle.c */
void orc_scan_topk_chunk(int type, int metric, int tier, size_t dim, const void *rows, size_t stride, size_t n, size_t label0,
                         const void *queries, size_t qstride, size_t nq, size_t k, int nthreads, size_t *labels, float *scores,
                         size_t *counts);

/* Counter-based synthetic data shared with the CUDA generator: element (row,col)
 * of stream `seed` is U(-1,1) from a 64-bit mix; bit-identical on host and device. */
uint64_t orc_mix64(uint64_t seed, uint64_t a, uint64_t b);
float orc_synth_f32(uint64_t seed, uint64_t row, uint64_t col);
void orc_synth_rows(int type, uint64_t seed, uint64_t row0, size_t nrows, size_t dim, void *out);

/* ---- postings_oracle.c ---------------------------------------------------------------------- */
/* qint / varint codecs (RS/qint/src/lib.rs:149-286, RS/varint/src/lib.rs). */
size_t orc_qint_encode(const uint32_t *vals, int n, uint8_t *out);            /* n in 2..4 */
size_t orc_qint_decode(const uint8_t *in, int n, uint32_t *vals);
size_t orc_varint_encode(uint64_t v, uint8_t *out);
size_t orc_varint_decode(const uint8_t *in, uint64_t *v);

/* Inverted index with 100-entry (1000 for doc-id-only codecs) delta-coded blocks
 * (RS/inverted_index/src/index/core.rs:31-94,235-358). */
enum {
    ORC_CODEC_FULL = 0,       /* qint4[delta,freq,fieldMask,offsetsLen] + offsets   codec/full.rs:66-69 */
    ORC_CODEC_FREQS_ONLY,     /* qint2[delta,freq]                                  codec/freqs_only.rs:33 */
    ORC_CODEC_FREQS_FIELDS,   /* qint3[delta,freq,fieldMask]                        codec/freqs_fields.rs:43 */
    ORC_CODEC_FIELDS_ONLY,    /* qint2[delta,fieldMask]                             codec/fields_only.rs:43 */
    ORC_CODEC_DOCIDS_ONLY,    /* varint(delta)                                      codec/doc_ids_only.rs:33 */
    ORC_CODEC_RAW_DOCIDS_ONLY,/* u32 LE (docId - block.first_doc_id)                codec/raw_doc_ids_only.rs:31-37 */
    ORC_CODEC_FREQS_OFFSETS,  /* qint3[delta,freq,offsetsLen] + offsets             codec/freqs_offsets.rs:32-50 */
    ORC_CODEC_OFFSETS_ONLY,   /* qint2[delta,offsetsLen] + offsets (freq 1)         codec/offsets_only.rs:31-48 */
    ORC_CODEC_FIELDS_OFFSETS, /* qint3[delta,fieldMask,offsetsLen] + offsets        codec/fields_offsets.rs:36-60 */
    ORC_CODEC_FULL_WIDE,      /* qint3[delta,freq,offsetsLen] + varint(mask u128) + offsets   codec/full.rs:197-217 */
    ORC_CODEC_FREQS_FIELDS_WIDE, /* qint2[delta,freq] + varint(mask u128)           codec/freqs_fields.rs:114-126 */
    ORC_CODEC_FIELDS_ONLY_WIDE,  /* varint(delta) + varint(mask u128)               codec/fields_only.rs:109-121 */
    ORC_CODEC_FIELDS_OFFSETS_WIDE /* qint2[delta,offsetsLen] + varint(mask u128) + offsets    codec/fields_offsets.rs:138-160 */
};
typedef struct OrcInvIndex OrcInvIndex;
OrcInvIndex *orc_ii_new(int codec);
void orc_ii_free(OrcInvIndex *ii);
/* Returns bytes the index grew by (0 for a silently skipped duplicate docId). */
size_t orc_ii_add(OrcInvIndex *ii, uint64_t doc_id, uint32_t freq, uint32_t field_mask, const uint8_t *offsets,
                  uint32_t offsets_len);
size_t orc_ii_num_blocks(const OrcInvIndex *ii);
size_t orc_ii_num_docs(const OrcInvIndex *ii);
/* Block view for the product's host decoder tests: pointers stay valid until the next add. */
void orc_ii_block(const OrcInvIndex *ii, size_t b, uint64_t *first_id, uint64_t *last_id, uint16_t *num_entries,
                  const uint8_t **buf, size_t *len);

typedef struct OrcReader OrcReader; /* IndexReaderCore: next_record / seek_record / skip_to */
OrcReader *orc_reader_new(const OrcInvIndex *ii, uint32_t field_mask_filter /* 0 = no filter */);
/* u128 field masks (the *Wide codecs): mask / filter as (lo, hi) 64-bit halves */
OrcReader *orc_reader_new_wide(const OrcInvIndex *ii, uint64_t filter_lo, uint64_t filter_hi);
size_t orc_ii_add_wide(OrcInvIndex *ii, uint64_t doc_id, uint32_t freq, uint64_t mask_lo, uint64_t mask_hi, const uint8_t *offsets,
                       uint32_t offsets_len);
int orc_reader_next_wide(OrcReader *r, uint64_t *doc_id, uint32_t *freq, uint64_t *mask_lo, uint64_t *mask_hi);
void orc_reader_free(OrcReader *r);
void orc_reader_rewind(OrcReader *r);
/* 1 = record produced, 0 = EOF.  (reader/core.rs:245-277) */
int orc_reader_next(OrcReader *r, uint64_t *doc_id, uint32_t *freq, uint32_t *field_mask);
/* first record with docId >= target; 1 = produced, 0 = EOF.  (reader/core.rs:279-345) */
int orc_reader_seek(OrcReader *r, uint64_t target, uint64_t *doc_id, uint32_t *freq, uint32_t *field_mask);

/* Iterator algebra over readers (RS/rqe_iterators/src/intersection.rs, union_flat.rs).
 * A child is a reader plus its query term weight; results carry per-child freq so the scorer
 * oracle can recurse over "term leaves" in the reference's child order. */
typedef struct {
    uint64_t doc_id;
    uint32_t n_children;        /* matched children at this doc */
    uint32_t child_index[16];   /* index into the ORIGINAL children array, in aggregate order */
    uint32_t child_freq[16];
} OrcHit;
/* Runs the whole iterator to EOF.  hits may be NULL to only count.  Children are sorted by
 * num_estimated like Intersection::new (intersection.rs:103-169) unless in_order. */
size_t orc_intersect(OrcReader **children, size_t n, OrcHit *hits, size_t cap);
size_t orc_union(OrcReader **children, size_t n, int quick_exit, OrcHit *hits, size_t cap);
/* skip_to driven variants used by the contract tests: position the iterator with a list of
 * targets, recording status (0 OK, 1 NOTFOUND, 2 EOF) and the docId landed on. */
size_t orc_intersect_skipto(OrcReader **children, size_t n, const uint64_t *targets, size_t nt, int *status,
                            uint64_t *landed);
size_t orc_union_skipto(OrcReader **children, size_t n, const uint64_t *targets, size_t nt, int *status,
                        uint64_t *landed);

/* Numeric codec (RS/inverted_index/src/codec/numeric.rs): one record = header, docId delta (< 2^56), value.  Returns the bytes
 * written / consumed.  compress_floats selects NumericFloatCompression. */
size_t orc_numeric_encode(uint64_t delta, double value, int compress_floats, uint8_t *out /* >= 16 bytes */);
size_t orc_numeric_decode(const uint8_t *in, uint64_t *delta, double *value);
int orc_numeric_in_range(double value, double min, double max, int min_inclusive, int max_inclusive);

/* IDF (RS/idf/src/lib.rs:36-110). */
double orc_idf(uint64_t total_docs, uint64_t term_docs);
double orc_idf_bm25(uint64_t total_docs, uint64_t term_docs);
/* proximity.rs is_within_range over term children: offsets[i] / lens[i] = the varint-delta position bytes of child i for the
 * document (len 0 = the child carries no offsets); has_slop 0 = no slop limit (in-order only) */
int orc_within_range(size_t n_children, const uint8_t *const *offsets, const size_t *lens, int has_slop, uint32_t max_slop,
                     int in_order);

/* ---- scorer_oracle.c ------------------------------------------------------------------------ */
/* One matched document seen by a scorer: a flat intersection/union of term leaves
 * (src/ext/default.c recursion collapses to this for the query shapes on the hot path). */
typedef struct {
    uint32_t n_terms;
    const uint32_t *freq;   /* per matched term */
    const double *idf;      /* QueryTerm_GetIDF      (legacy log2 idf)     */
    const double *bm25_idf; /* QueryTerm_GetBM25_IDF                       */
    const double *weight;   /* leaf weight (query node weight, default 1)  */
    double agg_weight;      /* weight of the aggregate node                */
    uint32_t doc_len;       /* dmd->docLen        */
    uint32_t max_freq;      /* dmd->maxTermFreq   */
    float doc_score;        /* dmd->score         */
} OrcScoreDoc;
typedef struct {
    uint64_t num_docs;
    uint64_t num_terms;
    double avg_doc_len;
} OrcIndexStats;
enum { ORC_SCORER_BM25STD = 0, ORC_SCORER_BM25, ORC_SCORER_TFIDF, ORC_SCORER_TFIDF_DOCNORM, ORC_SCORER_DOCSCORE,
       ORC_SCORER_BM25STD_TANH, ORC_SCORER_DISMAX };
/* slop: value GetSlop would return (1 when offsets are not modelled); min_score as passed by
 * RPScorer; tanh_factor for BM25STD.TANH. */
double orc_score(int scorer, const OrcIndexStats *st, const OrcScoreDoc *d, int slop, double min_score,
                 double tanh_factor);

/* GetSlop = IndexResult_MinOffsetDelta (src/index_result/index_result.c:51-108) over an aggregate of n term leaves / virtual
 * results: decoded positions of child i at pos[i*stride .. i*stride + npos[i]). */
int orc_min_offset_delta(size_t n, const uint32_t *npos, const uint32_t *pos, size_t stride, const int *is_virtual);

/* ---- tree_oracle.c: result TREES (nested aggregates) ---------------------------------------------
 * Node 0 is the root; parent[i] < i; the children of a node are the nodes naming it as parent, in index order (= the
 * aggregate's child order).  Term nodes carry their varint-delta position bytes (off_len 0 = none). */
enum { ORC_KIND_TERM = 0, ORC_KIND_INTERSECTION, ORC_KIND_UNION, ORC_KIND_VIRTUAL, ORC_KIND_NUMERIC };
typedef struct {
    size_t n_nodes;
    const int32_t *parent; /* -1 for the root */
    const int32_t *kind;   /* ORC_KIND_* */
    const uint32_t *freq;
    const double *weight, *idf, *bm25_idf;
    const uint32_t *off_start, *off_len;
    const uint8_t *bytes;
} OrcTree;
/* slop < 0: GetSlop = orc_tree_min_offset_delta of the tree */
double orc_score_tree(int scorer, const OrcIndexStats *st, const OrcTree *t, uint32_t doc_len, uint32_t max_freq, float doc_score,
                      int slop, double min_score, double tanh_factor);
/* RSIndexResult_IterateOffsets of a node: every position in the order yielded; returns the count (out filled up to cap) */
size_t orc_tree_offsets(const OrcTree *t, size_t node, uint32_t *out, size_t cap);
int orc_tree_has_offsets(const OrcTree *t, size_t node);
int orc_tree_min_offset_delta(const OrcTree *t);
int orc_tree_within_range(const OrcTree *t, int has_slop, uint32_t max_slop, int in_order);

/* Bulk-add every member of vocabulary rank `rank` (synthetic Zipf corpus below); returns the count. */
size_t orc_ii_fill_synth(OrcInvIndex *ii, uint64_t n_docs, uint64_t rank);
/* CPU baseline: nq 3-term AND + BM25STD + top-N queries (terms = nq*3 indexes), one query per thread. */
double orc_time_search3(OrcInvIndex **terms, size_t nq, const uint32_t *doc_len, uint64_t n_docs, double avg_doc_len,
                        size_t top_n, int nthreads, uint64_t *out_ids, double *out_scores, size_t *out_hits);

/* Synthetic Zipf postings shared with the CUDA path. */
uint64_t orc_synth_df(uint64_t n_docs, uint64_t rank);
int orc_synth_member(uint64_t n_docs, uint64_t rank, uint64_t doc, uint32_t *tf);
uint32_t orc_synth_doclen(uint64_t doc);

#ifdef __cplusplus
}
#endif
#endif
