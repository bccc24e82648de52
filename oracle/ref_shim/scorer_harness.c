/* TEST INFRASTRUCTURE ONLY — never linked into the product.
 *
 * Harness around the REFERENCE's default scorers: oracle/Makefile compiles
 * /root/reference/src/ext/default.c in place and links it with this file into
 * oracle/_ref/libscorers_ref.so.  default.c reads RSIndexResult nodes directly (data.tag, freq,
 * weight) and goes through a handful of Rust-FFI accessors for aggregates and query terms
 * (SURVEY.md §8c / Appendix B.6); those accessors are provided here over plain C side tables.
 * Everything else default.c references belongs to the query expanders and aborts if reached.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "redisearch.h"
#include "index_result_rs.h"
#include "types_ffi.h"
#include "query_term_ffi.h"
#include "score_explain.h"

#define MAX_TERMS 32

typedef struct {
  double idf, bm25_idf;
  char str[8];
} HarnessTerm;

static RSIndexResult g_leaves[MAX_TERMS];
static const RSIndexResult *g_leaf_ptrs[MAX_TERMS];
static HarnessTerm g_terms[MAX_TERMS];
static size_t g_nleaves;
static RSIndexResult g_root;

/* ---- accessors default.c needs on the scoring path ---------------------------------------- */
const RSAggregateResult *IndexResult_AggregateRefUnchecked(const RSIndexResult *r) {
  return (const RSAggregateResult *)&r->data;
}
struct AggregateRecordsSlice AggregateResult_GetRecordsSlice(const RSAggregateResult *agg) {
  struct AggregateRecordsSlice s = {g_leaf_ptrs, g_nleaves};
  return s;
}
size_t AggregateResult_NumChildren(const RSAggregateResult *agg) { return g_nleaves; }
const RSIndexResult *AggregateResult_Get(const RSAggregateResult *agg, size_t index) { return g_leaf_ptrs[index]; }
struct RSQueryTerm *IndexResult_QueryTermRef(const RSIndexResult *r) {
  return (struct RSQueryTerm *)&g_terms[r - g_leaves];
}
double QueryTerm_GetIDF(const struct RSQueryTerm *t) { return ((const HarnessTerm *)t)->idf; }
double QueryTerm_GetBM25_IDF(const struct RSQueryTerm *t) { return ((const HarnessTerm *)t)->bm25_idf; }
const char *QueryTerm_GetStrAndLen(const struct RSQueryTerm *t, size_t *out_len) {
  *out_len = 1;
  return ((const HarnessTerm *)t)->str;
}
void explain(RSScoreExplain *scrExp, char *fmt, ...) {}

/* ---- capture the static scorer function pointers ------------------------------------------- */
#define MAX_SCORERS 16
static struct { char name[48]; RSScoringFunction fn; } g_scorers[MAX_SCORERS];
static int g_nscorers;
static int reg_scorer(const char *alias, RSScoringFunction func, RSFreeFunction ff, void *privdata) {
  if (g_nscorers < MAX_SCORERS) {
    strncpy(g_scorers[g_nscorers].name, alias, sizeof(g_scorers[0].name) - 1);
    g_scorers[g_nscorers++].fn = func;
  }
  return REDISEARCH_OK;
}
static int reg_expander(const char *alias, RSQueryTokenExpander exp, RSFreeFunction ff, void *privdata) {
  return REDISEARCH_OK;
}
int DefaultExtensionInit(RSExtensionCtx *ctx);

static int g_slop = 1;
static int harness_slop(const RSIndexResult *r) { return g_slop; }

/* Score one document: an aggregate (is_union ? Union : Intersection) of n term leaves.
 * Returns NaN if the scorer name is unknown. */
double RefScore(const char *scorer, int is_union, size_t n, const uint32_t *freq, const double *idf,
                const double *bm25_idf, const double *weight, double agg_weight, uint32_t doc_len, uint32_t max_freq,
                float doc_score, size_t num_docs, double avg_doc_len, int slop, double min_score,
                uint64_t tanh_factor) {
  if (!g_nscorers) {
    RSExtensionCtx ctx = {reg_scorer, reg_expander};
    DefaultExtensionInit(&ctx);
  }
  RSScoringFunction fn = NULL;
  for (int i = 0; i < g_nscorers; i++)
    if (!strcmp(g_scorers[i].name, scorer)) fn = g_scorers[i].fn;
  if (!fn || n > MAX_TERMS) return NAN;
  memset(&g_root, 0, sizeof(g_root));
  memset(g_leaves, 0, sizeof(g_leaves));
  g_nleaves = n;
  uint32_t total = 0;
  for (size_t i = 0; i < n; i++) {
    g_leaves[i].data.tag = RSResultData_Term;
    g_leaves[i].freq = freq[i];
    g_leaves[i].weight = weight[i];
    g_leaf_ptrs[i] = &g_leaves[i];
    g_terms[i].idf = idf[i];
    g_terms[i].bm25_idf = bm25_idf[i];
    g_terms[i].str[0] = 't';
    total += freq[i];
  }
  g_root.data.tag = is_union ? RSResultData_Union : RSResultData_Intersection;
  g_root.freq = total;
  g_root.weight = agg_weight;
  RSDocumentMetadata dmd;
  memset(&dmd, 0, sizeof(dmd));
  dmd.score = doc_score;
  dmd.docLen = doc_len;
  dmd.maxTermFreq = max_freq;
  ScoringFunctionArgs args;
  memset(&args, 0, sizeof(args));
  args.indexStats.numDocs = num_docs;
  args.indexStats.avgDocLen = avg_doc_len;
  args.GetSlop = harness_slop;
  args.tanhFactor = tanh_factor;
  g_slop = slop;
  return fn(&args, &g_root, &dmd, min_score);
}


/* ---- GetSlop: the reference's own IndexResult_MinOffsetDelta (src/index_result/index_result.c:51-108, compiled in place by
 * oracle/Makefile) over DECODED term positions held in a side table.  The accessors below are the ones that file needs;
 * RSIndexResult_IterateOffsets (src/offset_vector.c:147-180) is reduced to its term / virtual cases: an aggregate of term
 * leaves is the only shape the hot path produces. ------------------------------------------------------------------------- */
#define MAX_POS 256
static uint32_t g_pos[MAX_TERMS][MAX_POS];
static uint32_t g_npos[MAX_TERMS];
typedef struct { uint32_t leaf, i; } PosIter;
static uint32_t positer_next(void *ctx, RSQueryTerm **t) {
  PosIter *it = ctx;
  if (it->i >= g_npos[it->leaf]) return RS_OFFSETVECTOR_EOF;
  return g_pos[it->leaf][it->i++];
}
static void positer_rewind(void *ctx) { ((PosIter *)ctx)->i = 0; }
static void positer_free(void *ctx) { free(ctx); }
static uint32_t emptyiter_next(void *ctx, RSQueryTerm **t) { return RS_OFFSETVECTOR_EOF; }
static void emptyiter_noop(void *ctx) {}
RSOffsetIterator RSIndexResult_IterateOffsets(const RSIndexResult *res) {
  if (res->data.tag != RSResultData_Term) {
    RSOffsetIterator e = {.ctx = NULL, .Next = emptyiter_next, .Rewind = emptyiter_noop, .Free = emptyiter_noop};
    return e;
  }
  PosIter *it = malloc(sizeof(*it));
  it->leaf = (uint32_t)(res - g_leaves);
  it->i = 0;
  RSOffsetIterator r = {.ctx = it, .Next = positer_next, .Rewind = positer_rewind, .Free = positer_free};
  return r;
}
const RSOffsetSlice *IndexResult_TermOffsetsRef(const RSIndexResult *r) { return (const RSOffsetSlice *)&g_npos[r - g_leaves]; }
uint32_t RSOffsetVector_Len(const RSOffsetSlice *offsets) { return *(const uint32_t *)offsets; }
const RSAggregateResult *IndexResult_AggregateRef(const RSIndexResult *r) {
  return (r->data.tag & (RSResultData_Intersection | RSResultData_Union)) ? (const RSAggregateResult *)&r->data : NULL;
}
const RSIndexResult *AggregateResult_GetUnchecked(const RSAggregateResult *agg, size_t index) { return g_leaf_ptrs[index]; }
uint8_t AggregateResult_KindMask(const RSAggregateResult *agg) {
  uint8_t m = 0;
  for (size_t i = 0; i < g_nleaves; i++) m |= (uint8_t)g_leaves[i].data.tag;
  return m;
}
int IndexResult_MinOffsetDelta(const RSIndexResult *r);

/* n children of an intersection / union; child i is a term leaf with npos[i] decoded positions pos[i*stride ..] (npos 0 = a term
 * without offsets), or a virtual result when is_virtual[i] (NOT / absent OPTIONAL children).  Returns what GetSlop returns. */
int RefMinOffsetDelta(int is_union, size_t n, const uint32_t *npos, const uint32_t *pos, size_t stride, const int *is_virtual) {
  if (n > MAX_TERMS) return -1;
  memset(&g_root, 0, sizeof(g_root));
  memset(g_leaves, 0, sizeof(g_leaves));
  g_nleaves = n;
  for (size_t i = 0; i < n; i++) {
    if (npos[i] > MAX_POS) return -1;
    g_leaves[i].data.tag = (is_virtual && is_virtual[i]) ? RSResultData_Virtual : RSResultData_Term;
    g_leaf_ptrs[i] = &g_leaves[i];
    g_npos[i] = npos[i];
    memcpy(g_pos[i], pos + i * stride, npos[i] * sizeof(uint32_t));
  }
  g_root.data.tag = is_union ? RSResultData_Union : RSResultData_Intersection;
  return IndexResult_MinOffsetDelta(&g_root);
}


/* HAMMING through the reference's own HammingDistanceScorer (src/ext/default.c:475-497): payload may be NULL (no payload). */
double RefHamming(const char *payload, size_t payload_len, const char *qdata, size_t qdatalen) {
  if (!g_nscorers) {
    RSExtensionCtx ctx = {reg_scorer, reg_expander};
    DefaultExtensionInit(&ctx);
  }
  RSScoringFunction fn = NULL;
  for (int i = 0; i < g_nscorers; i++)
    if (!strcmp(g_scorers[i].name, "HAMMING")) fn = g_scorers[i].fn;
  if (!fn) return NAN;
  RSPayload pl = {.data = (char *)payload, .len = payload_len};
  RSDocumentMetadata dmd;
  memset(&dmd, 0, sizeof(dmd));
  if (payload) {
    dmd.flags |= Document_HasPayload;
    dmd.payload = &pl;
  }
  ScoringFunctionArgs args;
  memset(&args, 0, sizeof(args));
  args.qdata = qdata;
  args.qdatalen = qdatalen;
  memset(&g_root, 0, sizeof(g_root));
  return fn(&args, &g_root, &dmd, 0);
}
