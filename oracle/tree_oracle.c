/* tree_oracle.c — CPU restatement of what the reference does with a result TREE (an aggregate whose children are themselves
 * aggregates: `(a|b) c`, stemming / synonym expansions under an AND, a phrase inside a larger query).
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Checked by tests/test_oracle_trees.py against
 *   - oracle/_ref/libscorers_ref.so: the reference's own src/ext/default.c, src/index_result/index_result.c and
 *     src/offset_vector.c compiled in place (RefTree* of ref_shim/scorer_harness.c), and
 *   - the known answers of RS/index_result/src/core/proximity.rs:409-470 (the Merge iterator).
 *
 * Restated here:
 *   scorers      src/ext/default.c  tfidfRecursive :68-106, bm25Recursive :164-211, bm25StdRecursive :253-301,
 *                dismaxRecursive :377-452 and their callers (final normalisation, min score, slop division)
 *   offsets      src/offset_vector.c _aoi_Next :216-239 (k-way merge, the FIRST smallest look-ahead advances, duplicates kept),
 *                RSIndexResult_IterateOffsets :155-189 (one child: that child's iterator), proximity.rs OffsetIter::Merge :53-67
 *   has offsets  src/index_result/index_result.c:19-42 / proximity.rs:72-91 (an aggregate: by the KIND MASK of its direct
 *                children — not Virtual-only, not exactly Numeric|Metric — whatever the streams hold)
 *   GetSlop      IndexResult_MinOffsetDelta, index_result.c:51-108, over the root's children
 *   proximity    proximity.rs within_range_in_order :134-180, within_range_unordered :184-220, is_within_range :262-299
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

enum { TAG_UNION = 1, TAG_INTERSECTION = 2, TAG_TERM = 4, TAG_VIRTUAL = 8, TAG_NUMERIC = 16, TAG_METRIC = 32 };
static int tag_of(int kind) {
    switch (kind) {
    case ORC_KIND_TERM: return TAG_TERM;
    case ORC_KIND_INTERSECTION: return TAG_INTERSECTION;
    case ORC_KIND_UNION: return TAG_UNION;
    case ORC_KIND_NUMERIC: return TAG_NUMERIC;
    default: return TAG_VIRTUAL;
    }
}
static int is_agg(const OrcTree *t, size_t n) { return t->kind[n] == ORC_KIND_INTERSECTION || t->kind[n] == ORC_KIND_UNION; }
/* children of `node` in index order */
static size_t children_of(const OrcTree *t, size_t node, size_t *out, size_t cap) {
    size_t n = 0;
    for (size_t i = node + 1; i < t->n_nodes; i++)
        if (t->parent[i] == (int32_t)node && n < cap) out[n++] = i;
    return n;
}
#define MAX_KIDS 64

/* ---------------------------------------------------------------- scorers ------------------- */
static double bm25std_calc(double idf, double f, int doc_len, double avg_doc_len, double weight) { /* default.c:241-250 */
    const float b = 0.75f, k1 = 1.2f;
    volatile float one_minus_b = 1.0f - b;
    volatile float b_len = b * (float)doc_len;
    volatile float k1p1 = k1 + 1;
    volatile double ratio = (double)b_len / avg_doc_len;
    volatile double inner = (double)one_minus_b + ratio;
    volatile double k1_inner = (double)k1 * inner;
    volatile double denom = f + k1_inner;
    volatile double num = weight * idf;
    num = num * f;
    num = num * (double)k1p1;
    return num / denom;
}
static double bm25_denominator(double f, double avg_doc_len) { /* f + k1 * (1.0f - b + b * avg) */
    const float b = 0.5f, k1 = 1.2f;
    volatile float one_minus_b = 1.0f - b;
    volatile double b_avg = (double)b * avg_doc_len;
    volatile double inner = (double)one_minus_b + b_avg;
    volatile double k1_inner = (double)k1 * inner;
    volatile double denom = f + k1_inner;
    return denom;
}
static double rec(int scorer, const OrcIndexStats *st, const OrcTree *t, size_t n, uint32_t doc_len) {
    const double w = t->weight[n], f = (double)t->freq[n];
    size_t kids[MAX_KIDS];
    const size_t nk = is_agg(t, n) ? children_of(t, n, kids, MAX_KIDS) : 0;
    switch (scorer) {
    case ORC_SCORER_TFIDF:
    case ORC_SCORER_TFIDF_DOCNORM:
        if (t->kind[n] == ORC_KIND_TERM) {
            volatile double r = w * f;
            r = r * t->idf[n];
            return r;
        }
        if (is_agg(t, n)) {
            volatile double ret = 0;
            for (size_t i = 0; i < nk; i++) ret = ret + rec(scorer, st, t, kids[i], doc_len);
            return w * ret;
        }
        return w * f;
    case ORC_SCORER_BM25:
        if (t->kind[n] == ORC_KIND_TERM) {
            volatile double num = w * t->idf[n];
            num = num * f;
            return num / bm25_denominator(f, st->avg_doc_len);
        }
        if (is_agg(t, n)) {
            volatile double ret = 0;
            for (size_t i = 0; i < nk; i++) ret = ret + rec(scorer, st, t, kids[i], doc_len);
            ret = ret * w;
            return ret;
        }
        if (t->freq[n]) {
            volatile double num = w * f;
            return num / bm25_denominator(f, st->avg_doc_len);
        }
        return 0;
    case ORC_SCORER_BM25STD:
    case ORC_SCORER_BM25STD_TANH:
        if (t->kind[n] == ORC_KIND_TERM) return bm25std_calc(t->bm25_idf[n], f, (int)doc_len, st->avg_doc_len, w);
        if (is_agg(t, n)) {
            volatile double ret = 0;
            for (size_t i = 0; i < nk; i++) ret = ret + rec(scorer, st, t, kids[i], doc_len);
            ret = ret * w;
            return ret;
        }
        if (t->kind[n] == ORC_KIND_VIRTUAL && t->freq[n] && w != 0) return bm25std_calc(1.0, 1.0, (int)doc_len, st->avg_doc_len, w);
        return 0;
    case ORC_SCORER_DISMAX: {
        volatile double ret = 0;
        if (t->kind[n] == ORC_KIND_INTERSECTION) {
            for (size_t i = 0; i < nk; i++) ret = ret + rec(scorer, st, t, kids[i], doc_len);
        } else if (t->kind[n] == ORC_KIND_UNION) {
            for (size_t i = 0; i < nk; i++) {
                const double c = rec(scorer, st, t, kids[i], doc_len);
                ret = ret > c ? ret : c; /* MAX(ret, child) */
            }
        } else {
            ret = f;
        }
        return w * ret;
    }
    }
    return NAN;
}
double orc_score_tree(int scorer, const OrcIndexStats *st, const OrcTree *t, uint32_t doc_len, uint32_t max_freq, float doc_score,
                      int slop, double min_score, double tanh_factor) {
    if (slop < 0) slop = orc_tree_min_offset_delta(t);
    switch (scorer) {
    case ORC_SCORER_BM25STD:
    case ORC_SCORER_BM25STD_TANH: {
        volatile double score = (double)doc_score * rec(scorer, st, t, 0, doc_len);
        if (scorer == ORC_SCORER_BM25STD_TANH) return tanh((1 / tanh_factor) * score);
        return score;
    }
    case ORC_SCORER_BM25: {
        volatile double score = (double)doc_score * rec(scorer, st, t, 0, doc_len);
        if (score < min_score) return 0;
        return score / slop;
    }
    case ORC_SCORER_TFIDF:
    case ORC_SCORER_TFIDF_DOCNORM: {
        if (doc_score == 0) return 0;
        const uint32_t norm = scorer == ORC_SCORER_TFIDF ? max_freq : doc_len;
        if (norm == 0) return 0;
        volatile double tfidf = (double)doc_score * rec(scorer, st, t, 0, doc_len);
        tfidf = tfidf / norm;
        if (tfidf < min_score) return 0;
        return tfidf / slop;
    }
    case ORC_SCORER_DOCSCORE: return (double)doc_score;
    case ORC_SCORER_DISMAX: return rec(scorer, st, t, 0, doc_len);
    }
    return NAN;
}

/* ---------------------------------------------------------------- offsets ------------------- */
typedef struct {
    uint32_t *v;
    size_t n;
} PosVec;
static void pv_push(PosVec *p, size_t *cap, uint32_t x) {
    if (p->n == *cap) {
        *cap = *cap ? *cap * 2 : 16;
        p->v = realloc(p->v, *cap * sizeof(uint32_t));
    }
    p->v[p->n++] = x;
}
/* every position RSIndexResult_IterateOffsets yields for the node, in the order it yields them */
static PosVec node_offsets(const OrcTree *t, size_t node) {
    PosVec out = {NULL, 0};
    size_t cap = 0;
    if (t->kind[node] == ORC_KIND_TERM) { /* varint deltas (RS/varint: 7-bit groups, most significant first, +1 per continuation) */
        const uint8_t *p = t->bytes + t->off_start[node], *end = p + t->off_len[node];
        uint32_t last = 0;
        while (p < end) {
            uint8_t b = *p++;
            uint64_t val = b & 0x7f;
            int bad = 0;
            while (b & 0x80) {
                if (p >= end) {
                    bad = 1;
                    break;
                }
                val += 1;
                b = *p++;
                val = (val << 7) | (b & 0x7f);
            }
            if (bad) break;
            last += (uint32_t)val;
            pv_push(&out, &cap, last);
        }
        return out;
    }
    if (!is_agg(t, node)) return out; /* virtual / numeric: the empty iterator */
    size_t kids[MAX_KIDS];
    const size_t nk = children_of(t, node, kids, MAX_KIDS);
    if (nk == 1) return node_offsets(t, kids[0]);
    PosVec ch[MAX_KIDS];
    size_t at[MAX_KIDS];
    for (size_t i = 0; i < nk; i++) {
        ch[i] = node_offsets(t, kids[i]);
        at[i] = 0;
    }
    for (;;) { /* _aoi_Next: the first child holding the smallest look-ahead yields it and advances */
        size_t mi = nk;
        uint32_t mv = 0xFFFFFFFFu;
        for (size_t i = 0; i < nk; i++)
            if (at[i] < ch[i].n && ch[i].v[at[i]] < mv) {
                mv = ch[i].v[at[i]];
                mi = i;
            }
        if (mi == nk) break;
        at[mi]++;
        pv_push(&out, &cap, mv);
    }
    for (size_t i = 0; i < nk; i++) free(ch[i].v);
    return out;
}
size_t orc_tree_offsets(const OrcTree *t, size_t node, uint32_t *out, size_t cap) {
    PosVec p = node_offsets(t, node);
    for (size_t i = 0; i < p.n && i < cap; i++) out[i] = p.v[i];
    free(p.v);
    return p.n;
}
int orc_tree_has_offsets(const OrcTree *t, size_t node) {
    if (t->kind[node] == ORC_KIND_TERM) return t->off_len[node] > 0;
    if (!is_agg(t, node)) return 0;
    size_t kids[MAX_KIDS];
    const size_t nk = children_of(t, node, kids, MAX_KIDS);
    int mask = 0;
    for (size_t i = 0; i < nk; i++) mask |= tag_of(t->kind[kids[i]]);
    return mask != TAG_VIRTUAL && mask != (TAG_NUMERIC | TAG_METRIC);
}

typedef struct {
    PosVec p;
    size_t i;
} Cur;
static uint32_t cur_next(Cur *c) { return c->i < c->p.n ? c->p.v[c->i++] : 0xFFFFFFFFu; }
static int cur_next_opt(Cur *c, uint32_t *pos) {
    if (c->i >= c->p.n) return 0;
    *pos = c->p.v[c->i++];
    return 1;
}

int orc_tree_min_offset_delta(const OrcTree *t) {
    if (!is_agg(t, 0)) return 1;
    size_t kids[MAX_KIDS];
    const size_t num = children_of(t, 0, kids, MAX_KIDS);
    if (num <= 1) return 1;
    int dist = 0;
    size_t i = 0;
    while (i < num) {
        while (i < num && !orc_tree_has_offsets(t, kids[i])) i++;
        if (i == num) break;
        Cur v1 = {node_offsets(t, kids[i]), 0};
        i++;
        while (i < num && !orc_tree_has_offsets(t, kids[i])) i++;
        if (i == num) {
            free(v1.p.v);
            break;
        }
        Cur v2 = {node_offsets(t, kids[i]), 0};
        uint32_t p1 = cur_next(&v1), p2 = cur_next(&v2);
        int cd = (int)(p2 > p1 ? p2 - p1 : p1 - p2);
        while (cd > 1 && p1 != 0xFFFFFFFFu && p2 != 0xFFFFFFFFu) {
            const uint32_t d = p2 > p1 ? p2 - p1 : p1 - p2;
            cd = (int)(d < (uint32_t)cd ? d : (uint32_t)cd);
            if (p2 > p1)
                p1 = cur_next(&v1);
            else
                p2 = cur_next(&v2);
        }
        free(v1.p.v);
        free(v2.p.v);
        dist += cd * cd;
    }
    return dist ? (int)sqrt((double)dist) : (int)(num - 1);
}

int orc_tree_within_range(const OrcTree *t, int has_slop, uint32_t max_slop_in, int in_order) {
    if (!is_agg(t, 0)) return 1;
    size_t kids[MAX_KIDS];
    const size_t num = children_of(t, 0, kids, MAX_KIDS);
    if (num <= 1) return 1;
    Cur it[MAX_KIDS];
    size_t n = 0;
    for (size_t i = 0; i < num; i++)
        if (orc_tree_has_offsets(t, kids[i])) {
            it[n].p = node_offsets(t, kids[i]);
            it[n].i = 0;
            n++;
        }
    int result = 0;
    const uint32_t max_slop = has_slop ? max_slop_in : 0xFFFFFFFFu;
    uint32_t positions[MAX_KIDS];
    if (n <= 1) {
        result = 1;
    } else if (in_order) {
        for (size_t i = 0; i < n; i++) positions[i] = 0;
        int done = 0;
        while (!done) {
            int32_t span = 0;
            int over = 0;
            for (size_t i = 0; i < n && !done; i++) {
                uint32_t pos;
                if (i == 0) {
                    if (!cur_next_opt(&it[0], &pos)) {
                        done = 1;
                        break;
                    }
                } else {
                    pos = positions[i];
                }
                const uint32_t last_pos = i == 0 ? 0u : positions[i - 1];
                while (pos < last_pos)
                    if (!cur_next_opt(&it[i], &pos)) {
                        done = 1;
                        break;
                    }
                if (done) break;
                positions[i] = pos;
                if (i > 0) {
                    span += (int32_t)pos - (int32_t)last_pos - 1;
                    if (span > 0 && (uint32_t)span > max_slop) {
                        over = 1;
                        break;
                    }
                }
            }
            if (done) break;
            if (!over) {
                result = 1;
                break;
            }
        }
    } else {
        int primed = 1;
        for (size_t i = 0; i < n && primed; i++) primed = cur_next_opt(&it[i], &positions[i]);
        if (primed) {
            uint32_t max_pos = 0;
            for (size_t i = 0; i < n; i++)
                if (positions[i] >= max_pos) max_pos = positions[i];
            for (;;) {
                uint32_t min_pos = 0xFFFFFFFFu;
                size_t min_idx = 0;
                for (size_t i = 0; i < n; i++)
                    if (positions[i] < min_pos) {
                        min_pos = positions[i];
                        min_idx = i;
                    }
                if (min_pos != max_pos) {
                    const int32_t span = (int32_t)max_pos - (int32_t)min_pos - ((int32_t)n - 1);
                    if (span < 0 || (uint32_t)span <= max_slop) {
                        result = 1;
                        break;
                    }
                }
                uint32_t np;
                if (!cur_next_opt(&it[min_idx], &np)) break;
                positions[min_idx] = np;
                if (np > max_pos) max_pos = np;
            }
        }
    }
    for (size_t i = 0; i < n; i++) free(it[i].p.v);
    return result;
}
