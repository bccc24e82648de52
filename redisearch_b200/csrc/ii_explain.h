// EXPLAINSCORE: the explanation tree the reference's scorers build next to the score (src/ext/default.c, the EXPLAIN macro and
// strExpCreateParent :58-65; reply side src/score_explain.c).  Host code: the *.B200 scorers evaluate the whole result set on
// the device; when a caller also wants the explanation of ONE result, its result tree is read back and walked here with the
// reference's own format strings.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace iiexplain {

enum Kind : int { Term = 0, Intersection = 1, Union = 2, Virtual = 3, Numeric = 4 };

struct TreeNode { // one RSIndexResult of the document's result tree, children in aggregate order
    int kind = Term;
    uint32_t freq = 0;
    double weight = 1.0, idf = 0.0, bm25_idf = 0.0;
    std::string term; // QueryTerm_GetStrAndLen of a term leaf (BM25STD prints it)
    std::vector<TreeNode> kids;
};
struct DocParams {
    uint32_t doc_len = 0, max_freq = 0; // dmd->docLen, dmd->maxTermFreq
    float doc_score = 1.0f;             // dmd->score
    double avg_doc_len = 0.0;           // ctx->indexStats.avgDocLen
};
struct Explain { // RSScoreExplain
    std::string str;
    std::vector<Explain> kids;
};

// The score the reference computes for the tree and, in `out`, the explanation rooted where ctx->scrExp points after the call
// (strExpCreateParent wraps the tree's own node once, BM25STD.TANH twice).  scorer: II_Scorer numbering (0 BM25STD, 1 BM25,
// 2 TFIDF, 3 TFIDF.DOCNORM, 4 DOCSCORE, 5 BM25STD.TANH, 6 DISMAX).  slop = what ctx->GetSlop returns for the result.
double explain_score(int scorer, const TreeNode &root, const DocParams &d, int slop, double min_score, uint64_t tanh_factor, Explain &out);
// HAMMING (default.c:475-497): result = the score (0 = payloads vary in length)
void explain_hamming(double result, size_t qdatalen, Explain &out);

} // namespace iiexplain
