// The HybridIterator state machine of src/iterators/hybrid_reader.c over the B200 index: mode choice (:668-691),
// batches mode with its batch-size formula and alternating merge (prepareResults :372-443, alternatingIterate :140-169,
// HR_ReadInBatch / HR_SkipToInBatch :60-88), the policy review that may switch to ad-hoc mid-query
// (reviewHybridSearchPolicy :346-370) and the ad-hoc mode (computeDistances_RAM :289-335).  The child is ANY iterator with
// the reference's QueryIterator vtable (src/iterators/iterator_api.h:46-151): a B200 AND / OR result, or a host iterator.
//
// What runs where: every VecSim call is the same C-ABI entry point hybrid_reader.c calls (VecSimIndex_PreferAdHocSearch,
// VecSimBatchIterator_{New,Next,HasNext,Free}, VecSimQueryReply_* — i.e. the device scans of vecsim_index.cpp); the ad-hoc
// mode replaces the per-document VecSimIndex_GetDistanceFrom_Unsafe round trips by ONE fused device call over the drained
// child docIds (FlatIndex::topk_filtered).  The heap is the reference's: ordered by cmpVecSimResByScore (:35-44).
#include "vecsim_index.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {
// layout of QueryIterator up to the members used here (src/iterators/iterator_api.h:46-151)
struct ChildIter {
    uint32_t type;
    bool atEOF;
    uint64_t lastDocId;
    void *current;
    size_t (*NumEstimated)(const ChildIter *);
    int (*Read)(ChildIter *);
    int (*SkipTo)(ChildIter *, uint64_t);
    int (*Revalidate)(ChildIter *, void *);
    void (*Free)(ChildIter *);
    void (*Rewind)(ChildIter *);
};
enum { IT_OK = 0, IT_NOTFOUND = 1, IT_EOF = 2, IT_TIMEOUT = 3 };

struct Hit {
    double score;
    uint64_t doc;
};
// cmpVecSimResByScore: score ascending; among equal scores the entry with the SMALLER docId compares greater
inline bool hit_less(const Hit &a, const Hit &b) {
    if (a.score < b.score) return true;
    if (a.score > b.score) return false;
    return !(a.doc < b.doc) && a.doc != b.doc; // a < b  <=>  a.doc > b.doc
}
struct TopHeap { // k is small: a sorted vector plays the min-max heap (mmh_*), same order relation
    size_t k;
    std::vector<Hit> v; // ascending by hit_less: back() is the max
    bool full() const { return v.size() >= k; }
    void insert(const Hit &h) { v.insert(std::upper_bound(v.begin(), v.end(), h, hit_less), h); }
    void exchange_max(const Hit &h) {
        v.pop_back();
        insert(h);
    }
    double upper() const { return v.back().score; }
};
} // namespace

extern "C" {

// Returns a VecSimQueryReply_Code (0 OK, 1 timed out) or -1.  Results ordered by (score asc, docId asc) like
// HR_ReadKnnUnsorted's consumers sort them; *out_mode = the VecSearchMode the query ended in, *out_iterations = batches run.
int VecSimB200_HybridTopK(VecSimIndex *index, const void *queryBlob, size_t k, void *child_iterator, VecSimQueryParams *qp,
                          size_t *out_labels, double *out_scores, size_t *out_count, int *out_mode, size_t *out_iterations) {
    using rsb200::FlatIndex;
    FlatIndex *ix = reinterpret_cast<FlatIndex *>(index);
    ChildIter *child = static_cast<ChildIter *>(child_iterator);
    if (out_count) *out_count = 0;
    if (out_iterations) *out_iterations = 0;
    if (!ix || !out_labels || !out_scores || !out_count) return -1;
    VecSimQueryParams local{};
    if (qp) local = *qp;
    int mode;
    if (!child || k == 0) { // :668-671: no child, or nothing to return -> plain KNN
        mode = STANDARD_KNN;
        VecSimQueryReply *rep = VecSimIndex_TopKQuery(index, queryBlob, k, &local, BY_SCORE);
        const int code = VecSimQueryReply_GetCode(rep);
        VecSimQueryReply_Iterator *it = VecSimQueryReply_GetIterator(rep);
        size_t w = 0;
        while (VecSimQueryReply_IteratorHasNext(it) && w < k) {
            VecSimQueryResult *r = VecSimQueryReply_IteratorNext(it);
            out_labels[w] = (size_t)VecSimQueryResult_GetId(r);
            out_scores[w] = VecSimQueryResult_GetScore(r);
            w++;
        }
        VecSimQueryReply_IteratorFree(it);
        VecSimQueryReply_Free(rep);
        *out_count = w;
        if (out_mode) *out_mode = mode;
        return code;
    }
    // :672-691 mode choice
    size_t subset = child->NumEstimated(child);
    const size_t index_size = VecSimIndex_IndexSize(index);
    if (subset > index_size) subset = index_size;
    if (local.searchMode)
        mode = (int)local.searchMode;
    else
        mode = VecSimIndex_PreferAdHocSearch(index, subset, k, true) ? HYBRID_ADHOC_BF : HYBRID_BATCHES;

    TopHeap heap{k, {}};
    int code = VecSim_QueryReply_OK;
    auto adhoc = [&]() -> int { // computeDistances_RAM: child docIds in ascending order -> distances -> the k best
        std::vector<uint32_t> ids;
        int st;
        while ((st = child->Read(child)) != IT_EOF) {
            if (st == IT_TIMEOUT) return VecSim_QueryReply_TimedOut;
            if (child->lastDocId <= 0xFFFFFFFEull) ids.push_back((uint32_t)child->lastDocId);
        }
        std::vector<size_t> lab(k);
        std::vector<double> sc(k);
        size_t cnt = 0;
        const int rc = ix->topk_filtered(queryBlob, k, ids.data(), ids.size(), false, lab.data(), sc.data(), &cnt);
        if (rc != 0) return -1;
        heap.v.clear();
        for (size_t i = 0; i < cnt; i++) heap.insert(Hit{sc[i], lab[i]});
        return VecSim_QueryReply_OK;
    };

    if (mode == HYBRID_ADHOC_BF) {
        code = adhoc();
    } else {
        mode = HYBRID_BATCHES;
        if (child->NumEstimated(child) != 0) { // :385-387
            VecSimBatchIterator *bit = VecSimBatchIterator_New(index, queryBlob, &local);
            double upper_bound = INFINITY;
            size_t child_num_estimated = child->NumEstimated(child);
            if (child_num_estimated > index_size) child_num_estimated = index_size;
            const size_t child_upper_bound = child_num_estimated;
            bool switched = false;
            while (VecSimBatchIterator_HasNext(bit)) {
                if (out_iterations) (*out_iterations)++;
                const size_t n_res_left = k - heap.v.size();
                size_t batch_size = local.batchSize;
                if (batch_size == 0) batch_size = (size_t)(n_res_left * ((float)VecSimIndex_IndexSize(index) / child_num_estimated) + 1); // :400-404
                VecSimQueryReply *rep = VecSimBatchIterator_Next(bit, batch_size, BY_ID);
                code = VecSimQueryReply_GetCode(rep);
                if (code == VecSim_QueryReply_TimedOut) {
                    VecSimQueryReply_Free(rep);
                    break;
                }
                VecSimQueryReply_Iterator *it = VecSimQueryReply_GetIterator(rep);
                child->Rewind(child);
                // alternatingIterate :140-169
                Hit cur{0, 0};
                auto read_in_batch = [&]() -> int {
                    if (!VecSimQueryReply_IteratorHasNext(it)) return IT_EOF;
                    VecSimQueryResult *r = VecSimQueryReply_IteratorNext(it);
                    cur = Hit{VecSimQueryResult_GetScore(r), (uint64_t)VecSimQueryResult_GetId(r)};
                    return IT_OK;
                };
                auto skip_in_batch = [&](uint64_t doc) -> int {
                    while (VecSimQueryReply_IteratorHasNext(it)) {
                        VecSimQueryResult *r = VecSimQueryReply_IteratorNext(it);
                        const uint64_t id = (uint64_t)VecSimQueryResult_GetId(r);
                        if (doc > id) continue;
                        cur = Hit{VecSimQueryResult_GetScore(r), id};
                        return IT_OK;
                    }
                    return IT_EOF;
                };
                int child_status = child->Read(child);
                int vec_status = read_in_batch();
                while (child_status == IT_OK && vec_status == IT_OK) {
                    if (cur.doc == child->lastDocId) {
                        if (!heap.full() || cur.score < upper_bound) {
                            if (!heap.full())
                                heap.insert(cur);
                            else
                                heap.exchange_max(cur);
                            upper_bound = heap.upper();
                        }
                        child_status = child->Read(child);
                        vec_status = read_in_batch();
                    } else if (cur.doc > child->lastDocId) {
                        child_status = child->SkipTo(child, cur.doc);
                        if (child_status == IT_NOTFOUND) child_status = IT_OK;
                    } else if (VecSimQueryReply_IteratorHasNext(it)) {
                        vec_status = skip_in_batch(child->lastDocId);
                    } else {
                        break;
                    }
                }
                VecSimQueryReply_IteratorFree(it);
                VecSimQueryReply_Free(rep);
                if (heap.v.size() == k) break;
                // reviewHybridSearchPolicy :346-370
                bool change = false;
                if (!((int)local.searchMode == HYBRID_BATCHES && local.batchSize)) {
                    const size_t new_results = heap.v.size() - (k - n_res_left);
                    const float cur_ratio = (float)new_results / n_res_left;
                    const size_t cur_est = (size_t)(cur_ratio * VecSimIndex_IndexSize(index));
                    child_num_estimated = (child_num_estimated + cur_est) / 2;
                    if (child_num_estimated > child_upper_bound) child_num_estimated = child_upper_bound;
                    if ((int)local.searchMode != HYBRID_BATCHES) change = VecSimIndex_PreferAdHocSearch(index, child_num_estimated, k, false);
                    if (child_num_estimated == 0) child_num_estimated = 1; // the reference would divide by zero in the next batch size
                }
                if (change) {
                    switched = true;
                    break;
                }
            }
            VecSimBatchIterator_Free(bit);
            if (switched) { // :430-438 batches -> ad-hoc: drop what was found, restart
                mode = HYBRID_BATCHES_TO_ADHOC_BF;
                heap.v.clear();
                child->Rewind(child);
                code = adhoc();
            }
        }
    }
    ix->set_last_mode((VecSearchMode)mode);
    if (code < 0) return -1;
    std::vector<Hit> res = heap.v;
    std::sort(res.begin(), res.end(), [](const Hit &a, const Hit &b) { return a.score < b.score || (a.score == b.score && a.doc < b.doc); });
    for (size_t i = 0; i < res.size(); i++) {
        out_labels[i] = (size_t)res[i].doc;
        out_scores[i] = res[i].score;
    }
    *out_count = res.size();
    if (out_mode) *out_mode = mode;
    return code;
}

} // extern "C"
