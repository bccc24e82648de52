/* Test host for the drop-in boundary of libii_b200.so: plays the part of RediSearch (tests/test_boundary_harness.py).
 *
 *   - exports the InvertedIndex block accessors of RS/headers/inverted_index_ffi.h (+ IndexBlock_DataLen) over a toy
 *     in-memory index whose blocks it encodes itself (FreqsOnly: qint2[delta, freq], 100 entries per block);
 *   - dlopen()s the library like Extension_LoadDynamic (src/extension.c:121-145), calls RS_ExtensionInit with a capturing
 *     RSExtensionCtx;
 *   - builds  (A AND B AND foreign C [AND NOT D] [AND OPTIONAL E])  through NewIntersectionIterator with the reference's
 *     signature and ownership rules, the children being B200 term leaves (II_NewTermIterator_FromIndex through a term
 *     cache) and one FOREIGN iterator implemented here;
 *   - walks the result with Read() and asks the registered "BM25STD.B200" RSScoringFunction for every result;
 *   - prints "docId score(hex float) freq" lines for the Python side to compare with the oracle.
 * usage: ext_harness <libii_b200.so> <dir with a.bin b.bin c.bin d.bin e.bin doclen.bin> <n_docs> <avg_doc_len>
 * each list file: u32 n, then n x (u32 docId, u32 freq). */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ii_b200.h"

/* ---- toy host index -------------------------------------------------------------------------- */
typedef struct {
    uint64_t first, last;
    uint16_t n;
    uint8_t *data;
    size_t len;
} ToyBlock;
typedef struct {
    ToyBlock *blocks;
    size_t nblocks, entries;
    uint32_t gc_marker;
} ToyIndex;

size_t InvertedIndex_NumBlocks(const void *ii) { return ((const ToyIndex *)ii)->nblocks; }
const void *InvertedIndex_BlockRef(const void *ii, size_t i) { return &((const ToyIndex *)ii)->blocks[i]; }
uint32_t InvertedIndex_GcMarker(const void *ii) { return ((const ToyIndex *)ii)->gc_marker; }
size_t InvertedIndex_NumEntries(const void *ii) { return ((const ToyIndex *)ii)->entries; }
const char *IndexBlock_Data(const void *b) { return (const char *)((const ToyBlock *)b)->data; }
size_t IndexBlock_DataLen(const void *b) { return ((const ToyBlock *)b)->len; }
uint64_t IndexBlock_FirstId(const void *b) { return ((const ToyBlock *)b)->first; }
uint64_t IndexBlock_LastId(const void *b) { return ((const ToyBlock *)b)->last; }
uint16_t IndexBlock_NumEntries(const void *b) { return ((const ToyBlock *)b)->n; }

/* what NewInvIndIterator_TermQuery resolves in the host besides the block accessors */
uint32_t InvertedIndex_Flags(const void *ii) { (void)ii; return 0x10; /* Index_StoreFreqs: the FreqsOnly encoding */ }
uint32_t InvertedIndex_NumDocs(const void *ii) { return (uint32_t)((const ToyIndex *)ii)->entries; }
typedef struct { size_t numDocs, numTerms; double avgDocLen; } ToyStats;
typedef struct { ToyStats stats; } ToySpec;                    /* stands for IndexSpec */
typedef struct { void *redisCtx; ToySpec *spec; } ToySearchCtx; /* src/search_ctx.h:60-64 */
void IndexSpec_GetStats(void *sp, ToyStats *out) { *out = ((ToySpec *)sp)->stats; }
typedef struct { double idf, bm25_idf; int freed; } ToyTerm;    /* stands for RSQueryTerm */
void QueryTerm_SetIDFs(void *t, double idf, double bm25_idf) { ((ToyTerm *)t)->idf = idf, ((ToyTerm *)t)->bm25_idf = bm25_idf; }
static int g_terms_freed;
void Term_Free(void *t) { ((ToyTerm *)t)->freed = 1; g_terms_freed++; }

static size_t put(uint8_t *p, uint32_t v) { /* minimal little-endian bytes, at least one */
    size_t n = 0;
    do {
        p[n++] = (uint8_t)v;
        v >>= 8;
    } while (v);
    return n;
}
static ToyIndex *toy_from_file(const char *path, uint32_t **ids_out, uint32_t **freqs_out, uint32_t *n_out) {
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    uint32_t n = 0;
    if (fread(&n, 4, 1, f) != 1) return NULL;
    uint32_t *pairs = malloc((size_t)n * 8 + 8);
    if (n && fread(pairs, 8, n, f) != n) return NULL;
    fclose(f);
    ToyIndex *ix = calloc(1, sizeof(*ix));
    ix->nblocks = (n + 99) / 100;
    ix->blocks = calloc(ix->nblocks ? ix->nblocks : 1, sizeof(ToyBlock));
    ix->entries = n;
    uint32_t *ids = malloc((size_t)n * 4 + 4), *fr = malloc((size_t)n * 4 + 4);
    for (uint32_t i = 0; i < n; i++) ids[i] = pairs[2 * i], fr[i] = pairs[2 * i + 1];
    for (size_t b = 0; b < ix->nblocks; b++) {
        ToyBlock *blk = &ix->blocks[b];
        const uint32_t lo = (uint32_t)b * 100, hi = lo + 100 < n ? lo + 100 : n;
        blk->data = malloc(900);
        blk->first = ids[lo];
        uint32_t last = ids[lo];
        size_t pos = 0;
        for (uint32_t i = lo; i < hi; i++) { /* qint2: lead byte, 2 bits per value = byte length - 1 */
            uint8_t *lead = &blk->data[pos++];
            const size_t a = put(blk->data + pos, ids[i] - last);
            pos += a;
            const size_t c = put(blk->data + pos, fr[i]);
            pos += c;
            *lead = (uint8_t)((a - 1) | ((c - 1) << 2));
            last = ids[i];
        }
        blk->last = last;
        blk->n = (uint16_t)(hi - lo);
        blk->len = pos;
    }
    free(pairs);
    if (ids_out) *ids_out = ids, *freqs_out = fr, *n_out = n;
    return ix;
}

/* ---- a FOREIGN iterator: what a numeric / tag leaf of the host looks like to the library ------- */
typedef struct {
    II_QueryIterator base;
    II_IndexResult res;
    uint32_t *ids, *freqs, n, pos;
    int *freed;
} Foreign;
static size_t f_est(const II_QueryIterator *b) { return ((const Foreign *)b)->n; }
static IteratorStatus f_read(II_QueryIterator *b) {
    Foreign *f = (Foreign *)b;
    if (f->pos >= f->n) {
        b->atEOF = true;
        b->current = NULL;
        return ITERATOR_EOF;
    }
    f->res.docId = f->ids[f->pos];
    f->res.freq = f->freqs[f->pos];
    b->lastDocId = f->ids[f->pos];
    b->current = &f->res;
    f->pos++;
    return ITERATOR_OK;
}
static IteratorStatus f_skip(II_QueryIterator *b, t_docId d) {
    Foreign *f = (Foreign *)b;
    while (f->pos < f->n && f->ids[f->pos] < d) f->pos++;
    if (f_read(b) == ITERATOR_EOF) return ITERATOR_EOF;
    return b->lastDocId == d ? ITERATOR_OK : ITERATOR_NOTFOUND;
}
static void f_rewind(II_QueryIterator *b) {
    Foreign *f = (Foreign *)b;
    f->pos = 0;
    b->atEOF = false;
    b->lastDocId = 0;
    b->current = NULL;
}
static void f_free(II_QueryIterator *b) {
    Foreign *f = (Foreign *)b;
    if (f->freed) (*f->freed)++;
    free(f);
}
static ValidateStatus f_reval(II_QueryIterator *b, struct IndexSpec *s) {
    (void)b;
    (void)s;
    return VALIDATE_OK;
}

/* ---- extension registration ------------------------------------------------------------------- */
typedef struct {
    void *extdata;
    const void *qdata;
    size_t qdatalen;
    struct {
        size_t numDocs, numTerms;
        double avgDocLen;
    } indexStats;
    void *scrExp;
    int (*GetSlop)(const void *);
    uint64_t tanhFactor;
} ScoringFunctionArgs;
typedef double (*RSScoringFunction)(const ScoringFunctionArgs *, const void *res, const void *dmd, double minScore);
static struct {
    char name[64];
    RSScoringFunction fn;
} g_scorers[16];
static int g_nscorers;
static int reg_scorer(const char *alias, RSScoringFunction fn, void (*ff)(void *), void *priv) {
    (void)ff;
    (void)priv;
    for (int i = 0; i < g_nscorers; i++)
        if (!strcmp(g_scorers[i].name, alias)) return 1; /* REDISEARCH_ERR: names are unique (extension.c:76-80) */
    snprintf(g_scorers[g_nscorers].name, 64, "%s", alias);
    g_scorers[g_nscorers++].fn = fn;
    return 0;
}
static int reg_expander(const char *a, void *e, void (*ff)(void *), void *p) {
    (void)a, (void)e, (void)ff, (void)p;
    return 0;
}
static RSScoringFunction scorer(const char *name) {
    for (int i = 0; i < g_nscorers; i++)
        if (!strcmp(g_scorers[i].name, name)) return g_scorers[i].fn;
    return NULL;
}

#define SYM(T, name) T name = (T)dlsym(lib, #name); if (!name) { fprintf(stderr, "missing %s\n", #name); return 2; }

typedef struct ExplainNode { /* src/score_explain.h:20-24 */
    char *str;
    int numChildren;
    struct ExplainNode *children;
} ExplainNode;
static void print_explain(const ExplainNode *e, int depth) {
    printf("%d %s\n", depth, e->str ? e->str : "(null)");
    for (int i = 0; i < e->numChildren; i++) print_explain(&e->children[i], depth + 1);
}
static void free_explain(ExplainNode *e) { /* recExplainDestroy, src/score_explain.c:34-41 */
    for (int i = 0; i < e->numChildren; i++) free_explain(&e->children[i]);
    free(e->children);
    free(e->str);
}

int main(int argc, char **argv) {
    if (argc < 5) return 2;
    void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    const size_t n_docs = strtoull(argv[3], NULL, 10);
    const double avg = atof(argv[4]);
    typedef int (*InitFn)(void *);
    SYM(InitFn, RS_ExtensionInit);
    struct {
        int (*RegisterScoringFunction)(const char *, RSScoringFunction, void (*)(void *), void *);
        int (*RegisterQueryExpander)(const char *, void *, void (*)(void *), void *);
    } ext = {reg_scorer, reg_expander};
    if (RS_ExtensionInit(&ext) != 0) return 3;
    printf("registered %d:", g_nscorers);
    for (int i = 0; i < g_nscorers; i++) printf(" %s", g_scorers[i].name);
    printf("\n");
    if (argc > 5 && !strcmp(argv[5], "register-only")) return 0;

    typedef II_QueryIterator *(*AndFn)(II_QueryIterator **, size_t, int32_t, bool, double);
    typedef II_QueryIterator *(*OrFn)(II_QueryIterator **, int32_t, bool, double, int, const char *, const void *);
    typedef II_QueryIterator *(*LeafFn)(const void *, II_Codec, double, double, double, II_TermCache *);
    typedef II_QueryIterator *(*WrapFn)(II_QueryIterator *, t_docId, double);
    typedef II_TermCache *(*CacheNewFn)(size_t);
    typedef void (*CacheFreeFn)(II_TermCache *);
    typedef II_TermCacheStats (*CacheStatsFn)(II_TermCache *);
    typedef II_DocTable *(*DtNewFn)(size_t, const uint32_t *, const float *, const uint32_t *);
    typedef void (*DtSetFn)(const II_DocTable *);
    typedef double (*IdfFn)(size_t, size_t);
    SYM(AndFn, NewIntersectionIterator);
    SYM(OrFn, NewUnionIterator);
    SYM(LeafFn, II_NewTermIterator_FromIndex);
    SYM(WrapFn, II_NewNotIterator);
    SYM(WrapFn, II_NewOptionalIterator);
    SYM(CacheNewFn, II_TermCache_New);
    SYM(CacheFreeFn, II_TermCache_Free);
    SYM(CacheStatsFn, II_TermCache_GetStats);
    SYM(DtNewFn, II_DocTable_New);
    SYM(DtSetFn, II_SetDefaultDocTable);
    SYM(IdfFn, II_CalculateIDF);
    SYM(IdfFn, II_CalculateIDF_BM25);

    char path[1024];
    const char *names[5] = {"a", "b", "c", "d", "e"};
    ToyIndex *ix[5];
    uint32_t *ids[5], *fr[5], n[5];
    for (int i = 0; i < 5; i++) {
        snprintf(path, sizeof(path), "%s/%s.bin", argv[2], names[i]);
        ix[i] = toy_from_file(path, &ids[i], &fr[i], &n[i]);
        if (!ix[i]) return 4;
    }
    snprintf(path, sizeof(path), "%s/doclen.bin", argv[2]);
    FILE *f = fopen(path, "rb");
    uint32_t *doc_len = malloc((n_docs + 1) * 4);
    if (!f || fread(doc_len, 4, n_docs + 1, f) != n_docs + 1) return 4;
    fclose(f);
    II_DocTable *dt = II_DocTable_New(n_docs, doc_len, NULL, NULL);
    if (!dt) return 5;
    II_SetDefaultDocTable(dt);
    II_TermCache *cache = II_TermCache_New((size_t)1 << 30);

    RSScoringFunction bm25 = scorer("BM25STD.B200"), tfidf = scorer("TFIDF.B200");
    if (!bm25 || !tfidf) return 6;
    ScoringFunctionArgs args;
    memset(&args, 0, sizeof(args));
    args.indexStats.numDocs = n_docs;
    args.indexStats.avgDocLen = avg;
    args.tanhFactor = 4;

    for (int variant = 0; variant < 4; variant++) {
        /* 0: A & B & foreign C;  1: the same again (term cache hits);  2: A & B & NOT D;  3: A & OPTIONAL E (weight 2) & B */
        int freed = 0;
        II_QueryIterator **its = malloc(4 * sizeof(*its)); /* ownership goes to the constructor */
        size_t k = 0;
        II_QueryIterator *la = II_NewTermIterator_FromIndex(ix[0], II_CODEC_FREQS_ONLY, 1.0, II_CalculateIDF(n_docs, n[0]), II_CalculateIDF_BM25(n_docs, n[0]), cache);
        II_QueryIterator *lb = II_NewTermIterator_FromIndex(ix[1], II_CODEC_FREQS_ONLY, 1.0, II_CalculateIDF(n_docs, n[1]), II_CalculateIDF_BM25(n_docs, n[1]), cache);
        if (!la || !lb) return 7;
        its[k++] = la;
        if (variant == 3) {
            II_QueryIterator *le = II_NewTermIterator_FromIndex(ix[4], II_CODEC_FREQS_ONLY, 1.0, II_CalculateIDF(n_docs, n[4]), II_CalculateIDF_BM25(n_docs, n[4]), cache);
            its[k++] = II_NewOptionalIterator(le, n_docs, 2.0);
        }
        its[k++] = lb;
        if (variant <= 1) {
            Foreign *fo = calloc(1, sizeof(*fo));
            fo->base.type = 0; /* IteratorType_InvIdxNumeric */
            fo->base.NumEstimated = f_est, fo->base.Read = f_read, fo->base.SkipTo = f_skip, fo->base.Rewind = f_rewind;
            fo->base.Free = f_free, fo->base.Revalidate = f_reval;
            fo->ids = ids[2], fo->freqs = fr[2], fo->n = n[2], fo->freed = &freed;
            fo->res.data.tag = II_ResultData_Numeric, fo->res.weight = 1.0; /* what the host's numeric iterator yields */
            its[k++] = &fo->base;
        } else if (variant == 2) {
            II_QueryIterator *ld = II_NewTermIterator_FromIndex(ix[3], II_CODEC_FREQS_ONLY, 1.0, II_CalculateIDF(n_docs, n[3]), II_CalculateIDF_BM25(n_docs, n[3]), cache);
            its[k++] = II_NewNotIterator(ld, n_docs, 1.0);
        }
        II_QueryIterator *it = NewIntersectionIterator(its, k, -1, false, 1.0);
        if (!it) return 8;
        if (variant <= 1 && freed != 1) return 9; /* the constructor owns and frees its children */
        printf("variant %d estimated %zu\n", variant, it->NumEstimated(it));
        t_docId prev = 0;
        while (it->Read(it) == ITERATOR_OK) {
            if (it->lastDocId <= prev || it->current->docId != it->lastDocId) return 10;
            prev = it->lastDocId;
            const double s = bm25(&args, it->current, NULL, 0.0);
            printf("%llu %a %u\n", (unsigned long long)it->lastDocId, s, it->current->freq);
        }
        if (!it->atEOF) return 11;
        /* SkipTo after Rewind: OK on a hit, NOTFOUND on the next greater */
        it->Rewind(it);
        if (prev > 1) {
            IteratorStatus st = it->SkipTo(it, prev);
            if (st != ITERATOR_OK || it->lastDocId != prev) return 12;
            if (tfidf(&args, it->current, NULL, 0.0) < 0) return 13; /* a second scorer re-scores the set */
        }
        it->Free(it);
        II_TermCacheStats cs = II_TermCache_GetStats(cache);
        printf("cache hits %zu misses %zu\n", cs.hits, cs.misses);
    }
    /* variant 4: A & B with the leaves built by NewInvIndIterator_TermQuery — the reference's own constructor signature; codec,
     * IDFs and ownership of the term come from the host accessors above */
    {
        typedef II_QueryIterator *(*TermFn)(const void *, const void *, II_FieldMaskOrIndex, void *, double);
        typedef void (*SetCacheFn)(II_TermCache *);
        SYM(TermFn, NewInvIndIterator_TermQuery);
        SYM(SetCacheFn, II_SetDefaultTermCache);
        II_SetDefaultTermCache(cache);
        ToySpec spec = {{n_docs, 0, avg}};
        ToySearchCtx sctx = {NULL, &spec};
        ToyTerm *ta = calloc(1, sizeof(*ta)), *tb = calloc(1, sizeof(*tb));
        II_FieldMaskOrIndex all;
        memset(&all, 0, sizeof(all));
        all.tag = 1;
        all.mask = ~(unsigned __int128)0; /* RS_FIELDMASK_ALL */
        II_QueryIterator **its = malloc(2 * sizeof(*its));
        its[0] = NewInvIndIterator_TermQuery(ix[0], &sctx, all, ta, 1.0);
        its[1] = NewInvIndIterator_TermQuery(ix[1], &sctx, all, tb, 1.0);
        if (!its[0] || !its[1]) return 20;
        if (its[0]->type != 1) return 21;
        if (ta->idf != II_CalculateIDF(n_docs, n[0]) || tb->bm25_idf != II_CalculateIDF_BM25(n_docs, n[1])) return 22; /* stored into the terms */
        II_QueryIterator *it = NewIntersectionIterator(its, 2, -1, false, 1.0);
        if (!it) return 23;
        if (g_terms_freed != 2) return 24; /* the leaves owned the terms; the constructor consumed the leaves */
        printf("variant 4 estimated %zu\n", it->NumEstimated(it));
        while (it->Read(it) == ITERATOR_OK) printf("%llu %a %u\n", (unsigned long long)it->lastDocId, bm25(&args, it->current, NULL, 0.0), it->current->freq);
        it->Free(it);
        II_TermCacheStats cs = II_TermCache_GetStats(cache);
        printf("cache hits %zu misses %zu\n", cs.hits, cs.misses);
        II_SetDefaultTermCache(NULL);
    }
    /* variant 5: (A | E) & B — a NESTED union under the intersection, both through the reference's constructor signatures.  The
     * union's result set moves into the intersection; the scorers recurse into it (weight 0.5 * sum over its matching children) */
    {
        II_QueryIterator **inner = malloc(2 * sizeof(*inner));
        inner[0] = II_NewTermIterator_FromIndex(ix[0], II_CODEC_FREQS_ONLY, 1.0, II_CalculateIDF(n_docs, n[0]), II_CalculateIDF_BM25(n_docs, n[0]), cache);
        inner[1] = II_NewTermIterator_FromIndex(ix[4], II_CODEC_FREQS_ONLY, 2.0, II_CalculateIDF(n_docs, n[4]), II_CalculateIDF_BM25(n_docs, n[4]), cache);
        II_QueryIterator *un = NewUnionIterator(inner, 2, false, 0.5, 0, NULL, NULL);
        if (!un) return 30;
        II_QueryIterator **its = malloc(2 * sizeof(*its));
        its[0] = un;
        its[1] = II_NewTermIterator_FromIndex(ix[1], II_CODEC_FREQS_ONLY, 1.0, II_CalculateIDF(n_docs, n[1]), II_CalculateIDF_BM25(n_docs, n[1]), cache);
        II_QueryIterator *it = NewIntersectionIterator(its, 2, -1, false, 1.5);
        if (!it) return 31;
        printf("variant 5 estimated %zu\n", it->NumEstimated(it));
        int explained = 0;
        while (it->Read(it) == ITERATOR_OK) {
            if (!explained) { /* EXPLAINSCORE: the node handed in comes back as the root of the reference's explanation tree */
                ExplainNode exp = {NULL, 0, NULL};
                args.scrExp = &exp;
                const double se = tfidf(&args, it->current, NULL, 0.0);
                args.scrExp = NULL;
                if (!exp.str || se != tfidf(&args, it->current, NULL, 0.0)) return 32;
                printf("explain %llu\n", (unsigned long long)it->lastDocId);
                print_explain(&exp, 0);
                printf("explain-end\n");
                free_explain(&exp);
                explained = 1;
            }
            const double s1 = bm25(&args, it->current, NULL, 0.0);
            const double s2 = tfidf(&args, it->current, NULL, 0.0);
            printf("%llu %a %u %a\n", (unsigned long long)it->lastDocId, s1, it->current->freq, s2);
        }
        it->Free(it);
        II_TermCacheStats cs = II_TermCache_GetStats(cache);
        printf("cache hits %zu misses %zu\n", cs.hits, cs.misses);
    }
    /* union of A and foreign C through the reference's NewUnionIterator signature */
    {
        II_QueryIterator **its = malloc(2 * sizeof(*its));
        its[0] = II_NewTermIterator_FromIndex(ix[0], II_CODEC_FREQS_ONLY, 1.0, 1.0, 1.0, cache);
        Foreign *fo = calloc(1, sizeof(*fo));
        fo->base.NumEstimated = f_est, fo->base.Read = f_read, fo->base.SkipTo = f_skip, fo->base.Rewind = f_rewind;
        fo->base.Free = f_free, fo->base.Revalidate = f_reval;
        fo->ids = ids[2], fo->freqs = fr[2], fo->n = n[2];
        fo->res.data.tag = II_ResultData_Numeric, fo->res.weight = 1.0;
        its[1] = &fo->base;
        II_QueryIterator *it = NewUnionIterator(its, 2, false, 1.0, 0, NULL, NULL);
        if (!it) return 14;
        size_t cnt = 0;
        while (it->Read(it) == ITERATOR_OK) cnt++;
        printf("union %zu\n", cnt);
        it->Free(it);
    }
    /* reduction rules: no children / a NULL child */
    {
        II_QueryIterator *e = NewIntersectionIterator(NULL, 0, -1, false, 1.0);
        if (!e || e->Read(e) != ITERATOR_EOF) return 15;
        e->Free(e);
        II_QueryIterator **its = malloc(2 * sizeof(*its));
        its[0] = II_NewTermIterator_FromIndex(ix[0], II_CODEC_FREQS_ONLY, 1.0, 1.0, 1.0, cache);
        its[1] = NULL;
        e = NewIntersectionIterator(its, 2, -1, false, 1.0);
        if (!e || e->type != II_IteratorType_Empty || e->Read(e) != ITERATOR_EOF) return 16;
        e->Free(e);
        /* phrase constraints are declined */
        its = malloc(sizeof(*its));
        its[0] = II_NewTermIterator_FromIndex(ix[0], II_CODEC_FREQS_ONLY, 1.0, 1.0, 1.0, cache);
        if (NewIntersectionIterator(its, 1, 0, true, 1.0) != NULL) return 17;
        its[0]->Free(its[0]);
        free(its);
    }
    II_TermCache_Free(cache);
    printf("HARNESS-OK\n");
    return 0;
}
