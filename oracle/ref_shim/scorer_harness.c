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
#include "varint_ffi.h"
#include "buffer.h"

#include <stddef.h>
#include <stdarg.h>

#define MAX_TERMS 32
#define MAX_NODES 512

typedef struct {
  double idf, bm25_idf;
  char str[8];
} HarnessTerm;

/* Result nodes live in a pool; what the Rust side of the reference keeps inside RSIndexResult.data (the children of an
 * aggregate, the term's offsets and query term) sits in a side table at the same index. */
typedef struct {
  size_t nkids;
  const RSIndexResult *kids[MAX_TERMS];
  HarnessTerm term;
  const char *off; /* varint-delta position bytes */
  uint32_t off_len;
} Side;
static RSIndexResult g_nodes[MAX_NODES];
static Side g_side[MAX_NODES];
#define g_root g_nodes[0]
static size_t node_index(const RSIndexResult *r) { return (size_t)(r - g_nodes); }
static Side *side_of_agg(const RSAggregateResult *agg) {
  const RSIndexResult *r = (const RSIndexResult *)((const char *)agg - offsetof(RSIndexResult, data));
  return &g_side[node_index(r)];
}
static void pool_reset(void) {
  memset(g_nodes, 0, sizeof(g_nodes));
  memset(g_side, 0, sizeof(g_side));
}

/* ---- accessors default.c / index_result.c / offset_vector.c need ---------------------------- */
const RSAggregateResult *IndexResult_AggregateRefUnchecked(const RSIndexResult *r) {
  return (const RSAggregateResult *)&r->data;
}
const RSAggregateResult *IndexResult_AggregateRef(const RSIndexResult *r) {
  return (r->data.tag & (RSResultData_Intersection | RSResultData_Union)) ? (const RSAggregateResult *)&r->data : NULL;
}
struct AggregateRecordsSlice AggregateResult_GetRecordsSlice(const RSAggregateResult *agg) {
  Side *sd = side_of_agg(agg);
  struct AggregateRecordsSlice s = {sd->kids, sd->nkids};
  return s;
}
size_t AggregateResult_NumChildren(const RSAggregateResult *agg) { return side_of_agg(agg)->nkids; }
const RSIndexResult *AggregateResult_Get(const RSAggregateResult *agg, size_t index) { return side_of_agg(agg)->kids[index]; }
const RSIndexResult *AggregateResult_GetUnchecked(const RSAggregateResult *agg, size_t index) { return side_of_agg(agg)->kids[index]; }
uint8_t AggregateResult_KindMask(const RSAggregateResult *agg) {
  Side *sd = side_of_agg(agg);
  uint8_t m = 0;
  for (size_t i = 0; i < sd->nkids; i++) m |= (uint8_t)sd->kids[i]->data.tag;
  return m;
}
struct RSQueryTerm *IndexResult_QueryTermRef(const RSIndexResult *r) {
  return (struct RSQueryTerm *)&g_side[node_index(r)].term;
}
double QueryTerm_GetIDF(const struct RSQueryTerm *t) { return ((const HarnessTerm *)t)->idf; }
double QueryTerm_GetBM25_IDF(const struct RSQueryTerm *t) { return ((const HarnessTerm *)t)->bm25_idf; }
const char *QueryTerm_GetStrAndLen(const struct RSQueryTerm *t, size_t *out_len) {
  *out_len = 1;
  return ((const HarnessTerm *)t)->str;
}
/* src/score_explain.c:62-72 (the file itself pulls in the reply machinery; this is its one function the scorers call) */
void explain(RSScoreExplain *scrExp, char *fmt, ...) {
  char *old = scrExp->str;
  va_list ap;
  va_start(ap, fmt);
  if (vasprintf(&scrExp->str, fmt, ap) < 0) scrExp->str = NULL;
  va_end(ap);
  free(old);
}
/* term offsets: the slice handle is the side entry itself */
const RSOffsetSlice *IndexResult_TermOffsetsRef(const RSIndexResult *r) { return (const RSOffsetSlice *)&g_side[node_index(r)]; }
uint32_t RSOffsetVector_Len(const RSOffsetSlice *offsets) { return ((const Side *)offsets)->off_len; }
const char *RSOffsetVector_GetData(const RSOffsetSlice *offsets, uint32_t *len) {
  *len = ((const Side *)offsets)->off_len;
  return ((const Side *)offsets)->off;
}
/* RS/varint (Rust): 7-bit groups, most significant first, +1 per continuation byte (golden vectors: tests/golden) */
uint32_t ReadVarint(BufferReader *b) {
  unsigned char c = (unsigned char)BUFFER_READ_BYTE(b);
  uint32_t val = c & 127;
  while (c >> 7) {
    ++val;
    c = (unsigned char)BUFFER_READ_BYTE(b);
    val = (val << 7) | (c & 127);
  }
  return val;
}
static size_t write_varint(uint32_t value, char *out) { /* the inverse, for callers that hand in decoded positions */
  unsigned char tmp[8];
  size_t pos = sizeof(tmp) - 1;
  tmp[pos] = value & 127;
  while (value >>= 7) tmp[--pos] = 128 | (--value & 127);
  memcpy(out, tmp + pos, sizeof(tmp) - pos);
  return sizeof(tmp) - pos;
}

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
  pool_reset();
  uint32_t total = 0;
  g_side[0].nkids = n;
  for (size_t i = 0; i < n; i++) {
    RSIndexResult *leaf = &g_nodes[1 + i];
    leaf->data.tag = RSResultData_Term;
    leaf->freq = freq[i];
    leaf->weight = weight[i];
    g_side[0].kids[i] = leaf;
    g_side[1 + i].term.idf = idf[i];
    g_side[1 + i].term.bm25_idf = bm25_idf[i];
    g_side[1 + i].term.str[0] = 't';
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


/* ---- GetSlop: the reference's own IndexResult_MinOffsetDelta (src/index_result/index_result.c:51-108) over the reference's own
 * offset iterators (src/offset_vector.c: the term iterator reading varint deltas, the aggregate iterator merging its children),
 * both compiled in place by oracle/Makefile. ------------------------------------------------------------------------------- */
int IndexResult_MinOffsetDelta(const RSIndexResult *r);
int RSIndexResult_HasOffsets(const RSIndexResult *res);
static char g_bytes[1 << 16];

/* n children of an intersection / union; child i is a term leaf with npos[i] decoded positions pos[i*stride ..] (npos 0 = a term
 * without offsets), or a virtual result when is_virtual[i] (NOT / absent OPTIONAL children).  Returns what GetSlop returns. */
int RefMinOffsetDelta(int is_union, size_t n, const uint32_t *npos, const uint32_t *pos, size_t stride, const int *is_virtual) {
  if (n > MAX_TERMS) return -1;
  pool_reset();
  size_t used = 0;
  g_side[0].nkids = n;
  for (size_t i = 0; i < n; i++) {
    RSIndexResult *leaf = &g_nodes[1 + i];
    leaf->data.tag = (is_virtual && is_virtual[i]) ? RSResultData_Virtual : RSResultData_Term;
    g_side[0].kids[i] = leaf;
    if (used + 5 * (size_t)npos[i] > sizeof(g_bytes)) return -1;
    g_side[1 + i].off = g_bytes + used;
    uint32_t last = 0;
    for (uint32_t k = 0; k < npos[i]; k++) {
      used += write_varint(pos[i * stride + k] - last, g_bytes + used);
      last = pos[i * stride + k];
    }
    g_side[1 + i].off_len = (uint32_t)((g_bytes + used) - g_side[1 + i].off);
  }
  g_root.data.tag = is_union ? RSResultData_Union : RSResultData_Intersection;
  return IndexResult_MinOffsetDelta(&g_root);
}

/* ---- result TREES: node 0 is the root, parent[i] < i, children in index order; kind 0 term, 1 intersection, 2 union, 3 virtual,
 * 4 numeric (the numbering of oracle.h's ORC_KIND_*).  Term nodes: varint-delta position bytes bytes[off_start .. +off_len). */
static const uint8_t kTag[5] = {RSResultData_Term, RSResultData_Intersection, RSResultData_Union, RSResultData_Virtual, RSResultData_Numeric};
int RefTreeLoad(size_t n_nodes, const int32_t *parent, const int32_t *kind, const uint32_t *freq, const double *weight,
                const double *idf, const double *bm25_idf, const uint32_t *off_start, const uint32_t *off_len, const char *bytes) {
  if (n_nodes == 0 || n_nodes > MAX_NODES) return -1;
  pool_reset();
  for (size_t i = 0; i < n_nodes; i++) {
    if (kind[i] < 0 || kind[i] > 4) return -1;
    g_nodes[i].data.tag = kTag[kind[i]];
    g_nodes[i].freq = freq[i];
    g_nodes[i].weight = weight[i];
    g_side[i].term.idf = idf[i];
    g_side[i].term.bm25_idf = bm25_idf[i];
    g_side[i].term.str[0] = 't';
    g_side[i].off = bytes + off_start[i];
    g_side[i].off_len = kind[i] == 0 ? off_len[i] : 0;
    if (i > 0) {
      if (parent[i] < 0 || (size_t)parent[i] >= i) return -1;
      Side *ps = &g_side[parent[i]];
      if (ps->nkids >= MAX_TERMS) return -1;
      ps->kids[ps->nkids++] = &g_nodes[i];
    }
  }
  return 0;
}
int RefTreeMinOffsetDelta(void) { return IndexResult_MinOffsetDelta(&g_root); }
int RefTreeHasOffsets(size_t node) { return RSIndexResult_HasOffsets(&g_nodes[node]); }
size_t RefTreeOffsets(size_t node, uint32_t *out, size_t cap) {
  RSOffsetIterator it = RSIndexResult_IterateOffsets(&g_nodes[node]);
  size_t n = 0;
  for (;;) {
    const uint32_t p = it.Next(it.ctx, NULL);
    if (p == RS_OFFSETVECTOR_EOF) break;
    if (n < cap) out[n] = p;
    n++;
  }
  it.Free(it.ctx);
  return n;
}
static int tree_slop(const RSIndexResult *r) { return IndexResult_MinOffsetDelta(r); }
/* the loaded tree through the reference's scorer; slop < 0: GetSlop is the reference's IndexResult_MinOffsetDelta
 * (src/extension.c:159 wires exactly that) */
double RefTreeScore(const char *scorer, uint32_t doc_len, uint32_t max_freq, float doc_score, size_t num_docs, double avg_doc_len,
                    int slop, double min_score, uint64_t tanh_factor) {
  if (!g_nscorers) {
    RSExtensionCtx ctx = {reg_scorer, reg_expander};
    DefaultExtensionInit(&ctx);
  }
  RSScoringFunction fn = NULL;
  for (int i = 0; i < g_nscorers; i++)
    if (!strcmp(g_scorers[i].name, scorer)) fn = g_scorers[i].fn;
  if (!fn) return NAN;
  RSDocumentMetadata dmd;
  memset(&dmd, 0, sizeof(dmd));
  dmd.score = doc_score;
  dmd.docLen = doc_len;
  dmd.maxTermFreq = max_freq;
  ScoringFunctionArgs args;
  memset(&args, 0, sizeof(args));
  args.indexStats.numDocs = num_docs;
  args.indexStats.avgDocLen = avg_doc_len;
  args.GetSlop = slop < 0 ? tree_slop : harness_slop;
  args.tanhFactor = tanh_factor;
  g_slop = slop;
  return fn(&args, &g_root, &dmd, min_score);
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
  pool_reset();
  return fn(&args, &g_root, &dmd, 0);
}

/* EXPLAINSCORE of the loaded tree: the reference's scorer run with ScoringFunctionArgs.scrExp set, the resulting RSScoreExplain tree
 * (rooted where ctx->scrExp points AFTER the call: strExpCreateParent re-roots it, default.c:58-65) serialised one node per line as
 * "<depth> <string>\n" in pre-order.  Returns the bytes needed. */
static size_t ser(const RSScoreExplain *e, int depth, char *buf, size_t cap, size_t at) {
  char head[16];
  const int hn = snprintf(head, sizeof(head), "%d ", depth);
  const char *str = e->str ? e->str : "(null)";
  const size_t sn = strlen(str);
  if (buf && at + hn + sn + 1 < cap) {
    memcpy(buf + at, head, hn);
    memcpy(buf + at + hn, str, sn);
    buf[at + hn + sn] = '\n';
  }
  at += hn + sn + 1;
  for (int i = 0; i < e->numChildren; i++) at = ser(&e->children[i], depth + 1, buf, cap, at);
  return at;
}
static void free_exp(RSScoreExplain *e) {
  for (int i = 0; i < e->numChildren; i++) free_exp(&e->children[i]);
  free(e->children);
  free(e->str);
}
size_t RefTreeExplain(const char *scorer, uint32_t doc_len, uint32_t max_freq, float doc_score, size_t num_docs, double avg_doc_len,
                      int slop, double min_score, uint64_t tanh_factor, double *score_out, char *buf, size_t cap) {
  if (!g_nscorers) {
    RSExtensionCtx ctx = {reg_scorer, reg_expander};
    DefaultExtensionInit(&ctx);
  }
  RSScoringFunction fn = NULL;
  for (int i = 0; i < g_nscorers; i++)
    if (!strcmp(g_scorers[i].name, scorer)) fn = g_scorers[i].fn;
  if (!fn) return 0;
  RSDocumentMetadata dmd;
  memset(&dmd, 0, sizeof(dmd));
  dmd.score = doc_score;
  dmd.docLen = doc_len;
  dmd.maxTermFreq = max_freq;
  ScoringFunctionArgs args;
  memset(&args, 0, sizeof(args));
  args.indexStats.numDocs = num_docs;
  args.indexStats.avgDocLen = avg_doc_len;
  args.GetSlop = slop < 0 ? tree_slop : harness_slop;
  args.tanhFactor = tanh_factor;
  args.scrExp = calloc(1, sizeof(RSScoreExplain));
  g_slop = slop;
  const double sc = fn(&args, &g_root, &dmd, min_score);
  if (score_out) *score_out = sc;
  RSScoreExplain *root = args.scrExp;
  const size_t n = ser(root, 0, buf, cap, 0);
  if (buf && cap) buf[n < cap ? n : cap - 1] = 0;
  free_exp(root);
  free(root);
  return n;
}
