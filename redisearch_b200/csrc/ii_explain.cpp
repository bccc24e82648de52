// See ii_explain.h.  Every format string and the order of the arithmetic are the reference's (src/ext/default.c); the file:line of
// each is given where it is used.  Pinned by tests/test_oracle_trees.py against the reference's own default.c compiled in place.
#include "ii_explain.h"

#include <cmath>
#include <cstdarg>
#include <cstdio>

namespace iiexplain {
namespace {

std::string fmt(const char *f, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, f);
    const int n = vsnprintf(buf, sizeof(buf), f, ap);
    va_end(ap);
    if (n < (int)sizeof(buf)) return std::string(buf, n > 0 ? (size_t)n : 0);
    std::string s((size_t)n + 1, '\0');
    va_start(ap, f);
    vsnprintf(&s[0], s.size(), f, ap);
    va_end(ap);
    s.resize((size_t)n);
    return s;
}
inline bool is_agg(const TreeNode &n) { return n.kind == Intersection || n.kind == Union; }

// default.c:241-250 CalculateBM25Std
double bm25std_calc(double idf, double f, int doc_len, double avg, double weight, const std::string &term, Explain &e) {
    const float b = 0.75f, k1 = 1.2f;
    const double ret = weight * idf * f * (k1 + 1) / (f + k1 * (1.0f - b + b * (float)doc_len / avg));
    e.str = fmt("%.*s: (%.2f = Weight %.2f * IDF %.2f * (F %.2f * (k1 1.2 + 1)) / (F %.2f + k1 1.2 * (1 - b 0.75 + b 0.75 *"
                " Doc Len %d / Average Doc Len %.2f)))",
                (int)term.size(), term.c_str(), ret, weight, idf, f, f, doc_len, avg);
    return ret;
}

double tfidf_rec(const TreeNode &r, Explain &e) { // :68-106
    if (r.kind == Term) {
        const double res = r.weight * ((double)r.freq) * r.idf;
        e.str = fmt("(TFIDF %.2f = Weight %.2f * TF %d * IDF %.2f)", res, r.weight, r.freq, r.idf);
        return res;
    }
    if (is_agg(r)) {
        double ret = 0;
        e.kids.resize(r.kids.size());
        for (size_t i = 0; i < r.kids.size(); i++) ret += tfidf_rec(r.kids[i], e.kids[i]);
        e.str = fmt("(Weight %.2f * total children TFIDF %.2f)", r.weight, ret);
        return r.weight * ret;
    }
    e.str = fmt("(TFIDF %.2f = Weight %.2f * Frequency %d)", r.weight * (double)r.freq, r.weight, r.freq);
    return r.weight * (double)r.freq;
}

double bm25_rec(const TreeNode &r, const DocParams &d, Explain &e) { // :164-211
    static const float b = 0.5f, k1 = 1.2f;
    const double f = (double)r.freq;
    double ret = 0;
    if (r.kind == Term) {
        ret = r.weight * r.idf * f / (f + k1 * (1.0f - b + b * d.avg_doc_len));
        e.str = fmt("(%.2f = Weight %.2f * IDF %.2f * F %d / (F %d + k1 1.2 * (1 - b 0.5 + b 0.5 * Average Len %.2f)))", ret, r.weight, r.idf,
                    r.freq, r.freq, d.avg_doc_len);
    } else if (is_agg(r)) {
        e.kids.resize(r.kids.size());
        for (size_t i = 0; i < r.kids.size(); i++) ret += bm25_rec(r.kids[i], d, e.kids[i]);
        e.str = fmt("(Weight %.2f * children BM25 %.2f)", r.weight, ret);
        ret *= r.weight;
    } else if (f) {
        ret = r.weight * f / (f + k1 * (1.0f - b + b * d.avg_doc_len));
        e.str = fmt("(%.2f = Weight %.2f * F %d / (F %d + k1 1.2 * (1 - b 0.5 + b 0.5 * Average Len %.2f)))", ret, r.weight, r.freq, r.freq,
                    d.avg_doc_len);
    } else {
        e.str = "Frequency 0 -> value 0";
    }
    return ret;
}

double bm25std_rec(const TreeNode &r, const DocParams &d, Explain &e) { // :253-301
    const double f = (double)r.freq;
    double ret = 0;
    if (r.kind == Term) {
        ret = bm25std_calc(r.bm25_idf, f, (int)d.doc_len, d.avg_doc_len, r.weight, r.term, e);
    } else if (is_agg(r)) {
        e.kids.resize(r.kids.size());
        for (size_t i = 0; i < r.kids.size(); i++) ret += bm25std_rec(r.kids[i], d, e.kids[i]);
        e.str = fmt("(Weight %.2f * children BM25 %.2f)", r.weight, ret);
        ret *= r.weight;
    } else if (r.kind == Virtual && f && r.weight) {
        ret = bm25std_calc(1.0, 1, (int)d.doc_len, d.avg_doc_len, r.weight, "*", e);
    } else {
        e.str = "Irrelevant token -> score is 0";
    }
    return ret;
}

double dismax_rec(const TreeNode &r, Explain &e) { // :377-452
    double ret = 0;
    if (r.kind == Intersection || r.kind == Union) {
        e.kids.resize(r.kids.size());
        for (size_t i = 0; i < r.kids.size(); i++) {
            const double c = dismax_rec(r.kids[i], e.kids[i]);
            if (r.kind == Intersection)
                ret += c;
            else
                ret = ret > c ? ret : c; // MAX(ret, child)
        }
        e.str = fmt("%.2f = Weight %.2f * children DISMAX %.2f", r.weight * ret, r.weight, ret);
    } else {
        ret = r.freq;
        e.str = fmt("DISMAX %.2f = Weight %.2f * Frequency %d", r.weight * ret, r.weight, r.freq);
    }
    return r.weight * ret;
}

// strExpCreateParent :58-65: the current root becomes the only child of a new one
void wrap(Explain &e) {
    Explain parent;
    parent.kids.push_back(std::move(e));
    e = std::move(parent);
}

} // namespace

double explain_score(int scorer, const TreeNode &root, const DocParams &d, int slop, double min_score, uint64_t tanh_factor, Explain &out) {
    out = Explain();
    switch (scorer) {
    case 0:   // BM25STD :304-316
    case 5: { // BM25STD.TANH :339-359
        const double bm25res = bm25std_rec(root, d, out);
        const double score = d.doc_score * bm25res;
        wrap(out);
        out.str = fmt("Final BM25 : words BM25 %.2f * document score %.2f", bm25res, d.doc_score);
        if (scorer == 0) return score;
        const double normalized = tanh((1 / (double)tanh_factor) * score);
        wrap(out);
        out.str = fmt("Final Normalized BM25 : tanh(stretch factor 1/%d * Final BM25 %.2f)", (int)tanh_factor, score);
        return normalized;
    }
    case 1: { // BM25 :214-233
        const double bm25res = bm25_rec(root, d, out);
        double score = d.doc_score * bm25res;
        wrap(out);
        if (score < min_score) {
            out.str = fmt("BM25 score of %.2f is smaller than minimum score %.2f", bm25res, score);
            return 0;
        }
        score /= slop;
        out.str = fmt("Final BM25 : words BM25 %.2f * document score %.2f / slop %d", bm25res, d.doc_score, slop);
        return score;
    }
    case 2:   // TFIDF :108-146
    case 3: { // TFIDF.DOCNORM
        if (d.doc_score == 0) {
            out.str = "Document score is 0";
            return 0;
        }
        const uint32_t norm = scorer == 2 ? d.max_freq : d.doc_len;
        if (norm == 0) {
            out.str = fmt("Document %s is 0", scorer == 2 ? "max frequency" : "length");
            return 0;
        }
        const double raw = tfidf_rec(root, out);
        double tfidf = d.doc_score * raw / norm;
        wrap(out);
        if (tfidf < min_score) {
            out.str = fmt("TFIDF score of %.2f is smaller than minimum score %.2f", tfidf, min_score);
            return 0;
        }
        tfidf /= slop;
        out.str = fmt("Final TFIDF : words TFIDF %.2f * document score %.2f / norm %d / slop %d", raw, d.doc_score, norm, slop);
        return tfidf;
    }
    case 4: // DOCSCORE :366-371
        out.str = fmt("Document's score is %.2f", d.doc_score);
        return d.doc_score;
    case 6: // DISMAX :455-459
        return dismax_rec(root, out);
    }
    return 0;
}

void explain_hamming(double result, size_t qdatalen, Explain &out) {
    out = Explain();
    if (result == 0) {
        out.str = "Payloads provided to scorer vary in length";
        return;
    }
    const size_t bits = (size_t)llround(1.0 / result - 1.0);
    out.str = fmt("String length is %zu. Bit count is %zu. Result is (1 / count + 1) = %.2f", qdatalen, bits, result);
}

} // namespace iiexplain
