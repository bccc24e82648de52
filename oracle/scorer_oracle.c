/* scorer_oracle.c — CPU restatement of the reference's default scorers (src/ext/default.c) for the
 * result shapes on the hot path: an aggregate (intersection or union) of term leaves.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Checked against oracle/_ref/libscorers_ref.so (the
 * reference's own default.c, compiled in place) by tests/test_oracle_scorers.py, and against the
 * golden scores of the reference's flow tests (tests/pytests/test_scorers.py:198-242).
 *
 * The C expression trees are reproduced operation by operation, including which sub-expressions are
 * evaluated in float (`1.0f - b`, `b * (float)doc_len`, `k1 + 1`) before being promoted to double.
 */
#include "oracle.h"

#include <math.h>

/* default.c:241-250 */
static double bm25std_term(double idf, double f, int doc_len, double avg_doc_len, double weight) {
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

/* default.c:164-211 (term leaf) */
static double bm25_term(double idf, double f, double avg_doc_len, double weight) {
    const float b = 0.5f, k1 = 1.2f;
    volatile float one_minus_b = 1.0f - b;
    volatile double b_avg = (double)b * avg_doc_len;
    volatile double inner = (double)one_minus_b + b_avg;
    volatile double k1_inner = (double)k1 * inner;
    volatile double denom = f + k1_inner;
    volatile double num = weight * idf;
    num = num * f;
    return num / denom;
}

double orc_score(int scorer, const OrcIndexStats *st, const OrcScoreDoc *d, int slop, double min_score,
                 double tanh_factor) {
    switch (scorer) {
    case ORC_SCORER_BM25STD:
    case ORC_SCORER_BM25STD_TANH: { /* :253-316, :339-359 */
        volatile double ret = 0;
        for (uint32_t i = 0; i < d->n_terms; i++)
            ret = ret + bm25std_term(d->bm25_idf[i], (double)d->freq[i], (int)d->doc_len, st->avg_doc_len, d->weight[i]);
        ret = ret * d->agg_weight;
        volatile double score = (double)d->doc_score * ret;
        if (scorer == ORC_SCORER_BM25STD_TANH) return tanh((1 / tanh_factor) * score);
        return score;
    }
    case ORC_SCORER_BM25: { /* :164-233 */
        volatile double ret = 0;
        for (uint32_t i = 0; i < d->n_terms; i++)
            ret = ret + bm25_term(d->idf[i], (double)d->freq[i], st->avg_doc_len, d->weight[i]);
        ret = ret * d->agg_weight;
        volatile double score = (double)d->doc_score * ret;
        if (score < min_score) return 0;
        return score / slop;
    }
    case ORC_SCORER_TFIDF:
    case ORC_SCORER_TFIDF_DOCNORM: { /* :68-153 */
        if (d->doc_score == 0) return 0;
        uint32_t norm = (scorer == ORC_SCORER_TFIDF) ? d->max_freq : d->doc_len;
        if (norm == 0) return 0;
        volatile double raw = 0;
        for (uint32_t i = 0; i < d->n_terms; i++) {
            volatile double leaf = d->weight[i] * (double)d->freq[i];
            leaf = leaf * d->idf[i];
            raw = raw + leaf;
        }
        raw = d->agg_weight * raw;
        volatile double tfidf = (double)d->doc_score * raw;
        tfidf = tfidf / norm;
        if (tfidf < min_score) return 0;
        return tfidf / slop;
    }
    case ORC_SCORER_DOCSCORE: return (double)d->doc_score; /* :366-371 */
    case ORC_SCORER_DISMAX: { /* :378-461, intersection: sum of weight*freq */
        volatile double ret = 0;
        for (uint32_t i = 0; i < d->n_terms; i++) ret = ret + d->weight[i] * (double)d->freq[i];
        return d->agg_weight * ret;
    }
    }
    return NAN;
}

/* GetSlop = IndexResult_MinOffsetDelta (src/index_result/index_result.c:51-108; wired at src/extension.c:159): over the
 * aggregate's children that carry offsets (term leaves with a non-empty offset vector; virtual results never do, :19-42),
 * taken as CONSECUTIVE pairs (child i with the next such child, which then opens the following pair), the smallest position
 * distance a two-pointer walk sees before it drops to <= 1 or either side ends; sqrt of the sum of squares, truncated.
 * npos[i] = 0: a term without offsets; is_virtual[i]: NOT / absent OPTIONAL child. */
int orc_min_offset_delta(size_t n, const uint32_t *npos, const uint32_t *pos, size_t stride, const int *is_virtual) {
    if (n <= 1) return 1;
#define ORC_HAS(i) (!(is_virtual && is_virtual[i]) && npos[i] > 0)
#define ORC_NEXT(c, k) ((k) < npos[c] ? pos[(c) * stride + (k)++] : 0xFFFFFFFFu)
    int dist = 0;
    size_t i = 0;
    while (i < n) {
        while (i < n && !ORC_HAS(i)) i++;
        if (i == n) break;
        const size_t a = i++;
        while (i < n && !ORC_HAS(i)) i++;
        if (i == n) break;
        const size_t b = i;
        uint32_t ka = 0, kb = 0;
        uint32_t p1 = ORC_NEXT(a, ka), p2 = ORC_NEXT(b, kb);
        int cd = (int)(p2 > p1 ? p2 - p1 : p1 - p2);
        while (cd > 1 && p1 != 0xFFFFFFFFu && p2 != 0xFFFFFFFFu) {
            const uint32_t d = p2 > p1 ? p2 - p1 : p1 - p2;
            cd = (int)(d < (uint32_t)cd ? d : (uint32_t)cd); /* MIN(unsigned, int) compares as unsigned */
            if (p2 > p1)
                p1 = ORC_NEXT(a, ka);
            else
                p2 = ORC_NEXT(b, kb);
        }
        dist += cd * cd;
    }
#undef ORC_HAS
#undef ORC_NEXT
    return dist ? (int)sqrt((double)dist) : (int)(n - 1);
}
