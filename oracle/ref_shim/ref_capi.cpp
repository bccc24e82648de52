// TEST INFRASTRUCTURE ONLY — never linked into, loaded by, or called from the product path.
//
// A thin extern "C" face over the *reference's own* brute-force VecSim code, compiled from the
// sources where they lie under /root/reference (see oracle/Makefile) into
// oracle/_ref/libvecsim_ref.so.  It plays the role of deps/VectorSimilarity/src/VecSim/vec_sim.cpp
// (which cannot be compiled offline because it includes the SVS headers, vec_sim.cpp:18) for the
// subset of calls the FLAT path needs (vec_sim.cpp:213-432).  Symbols carry a Ref_ prefix so the
// library can sit in the same process as the product's libvecsim_b200.so.
#include "VecSim/index_factories/brute_force_factory.h"
#include "VecSim/query_result_definitions.h"
#include "VecSim/batch_iterator.h"
#include "VecSim/spaces/spaces.h"
#include "VecSim/types/bfloat16.h"
#include "VecSim/types/float16.h"
#include "VecSim/vec_sim_interface.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

namespace {
struct Quiet {
    Quiet() { VecSimIndexInterface::logCallback = nullptr; }
} quiet_logger;

size_t drain(VecSimQueryReply *rep, size_t cap, size_t *ids, double *scores, int *code) {
    size_t n = rep->results.size();
    if (code) *code = (int)rep->code;
    for (size_t i = 0; i < n && i < cap; i++) {
        ids[i] = rep->results[i].id;
        scores[i] = rep->results[i].score;
    }
    delete rep;
    return n;
}
} // namespace

extern "C" {

void *Ref_IndexNew(int type, size_t dim, int metric, int multi, size_t blockSize) {
    BFParams p{};
    p.type = (VecSimType)type;
    p.dim = dim;
    p.metric = (VecSimMetric)metric;
    p.multi = multi != 0;
    p.initialCapacity = 0;
    p.blockSize = blockSize ? blockSize : DEFAULT_BLOCK_SIZE;
    return BruteForceFactory::NewIndex(&p, false);
}

void Ref_IndexFree(void *idx) {
    auto *index = (VecSimIndex *)idx;
    auto allocator = index->getAllocator(); // vec_sim.cpp:371-375: keep the allocator alive across delete
    delete index;
}

int Ref_AddVector(void *idx, const void *blob, size_t label) {
    return ((VecSimIndex *)idx)->addVector(blob, label);
}

// labels == NULL -> labels are label0 + i
void Ref_AddVectors(void *idx, const void *blobs, size_t n, size_t stride, const size_t *labels,
                    size_t label0) {
    auto *index = (VecSimIndex *)idx;
    for (size_t i = 0; i < n; i++)
        index->addVector((const char *)blobs + i * stride, labels ? labels[i] : label0 + i);
}

int Ref_DeleteVector(void *idx, size_t label) { return ((VecSimIndex *)idx)->deleteVector(label); }

size_t Ref_IndexSize(void *idx) { return ((VecSimIndex *)idx)->indexSize(); }

size_t Ref_TopK(void *idx, const void *q, size_t k, int order, size_t cap, size_t *ids,
                double *scores, int *code) {
    auto *index = (VecSimIndex *)idx;
    auto *rep = index->topKQuery(q, k, nullptr);
    // vec_sim.cpp:353-355
    if ((VecSimQueryReply_Order)order == BY_ID) sort_results_by_id(rep);
    return drain(rep, cap, ids, scores, code);
}

size_t Ref_Range(void *idx, const void *q, double radius, int order, size_t cap, size_t *ids,
                 double *scores, int *code) {
    auto *index = (VecSimIndex *)idx;
    auto *rep = index->rangeQuery(q, radius, nullptr, (VecSimQueryReply_Order)order);
    return drain(rep, cap, ids, scores, code);
}

double Ref_GetDistanceFrom(void *idx, size_t label, const void *q) {
    return ((VecSimIndex *)idx)->getDistanceFrom_Unsafe(label, q);
}

int Ref_PreferAdHoc(void *idx, size_t subset, size_t k, int initial) {
    return ((VecSimIndex *)idx)->preferAdHocSearch(subset, k, initial != 0);
}

void *Ref_BatchNew(void *idx, const void *q) {
    return ((VecSimIndex *)idx)->newBatchIterator(q, nullptr);
}
size_t Ref_BatchNext(void *it, size_t n, int order, size_t cap, size_t *ids, double *scores) {
    auto *rep = ((VecSimBatchIterator *)it)->getNextResults(n, (VecSimQueryReply_Order)order);
    return drain(rep, cap, ids, scores, nullptr);
}
int Ref_BatchHasNext(void *it) { return !((VecSimBatchIterator *)it)->isDepleted(); }
void Ref_BatchReset(void *it) { ((VecSimBatchIterator *)it)->reset(); }
void Ref_BatchFree(void *it) {
    auto *bi = (VecSimBatchIterator *)it;
    auto allocator = bi->getAllocator();
    delete bi;
}

// One distance with the tier the reference dispatches to on this CPU (spaces.cpp:50-146).
float Ref_Distance(int type, int metric, size_t dim, const void *a, const void *b) {
    using namespace spaces;
    unsigned char al = 0;
    auto m = (VecSimMetric)metric;
    switch ((VecSimType)type) {
    case VecSimType_FLOAT32:
        return GetDistFunc<float, float>(m, dim, &al)(a, b, dim);
    case VecSimType_FLOAT16:
        return GetDistFunc<vecsim_types::float16, float>(m, dim, &al)(a, b, dim);
    case VecSimType_BFLOAT16:
        return GetDistFunc<vecsim_types::bfloat16, float>(m, dim, &al)(a, b, dim);
    case VecSimType_INT8:
        return GetDistFunc<int8_t, float>(m, dim, &al)(a, b, dim);
    case VecSimType_UINT8:
        return GetDistFunc<uint8_t, float>(m, dim, &al)(a, b, dim);
    default:
        return __builtin_nanf("");
    }
}

// n distances of rows a[i] (stride bytes apart) against one b.
void Ref_Distances(int type, int metric, size_t dim, const void *a, size_t stride, size_t n,
                   const void *b, float *out) {
    for (size_t i = 0; i < n; i++)
        out[i] = Ref_Distance(type, metric, dim, (const char *)a + i * stride, b);
}

// vec_sim.cpp:238-254
void Ref_Normalize(void *blob, size_t dim, int type) {
    using namespace spaces;
    switch ((VecSimType)type) {
    case VecSimType_FLOAT32:
        GetNormalizeFunc<float>()(blob, dim);
        break;
    case VecSimType_FLOAT16:
        GetNormalizeFunc<vecsim_types::float16>()(blob, dim);
        break;
    case VecSimType_BFLOAT16:
        GetNormalizeFunc<vecsim_types::bfloat16>()(blob, dim);
        break;
    case VecSimType_INT8:
        GetNormalizeFunc<int8_t>()(blob, dim);
        break;
    case VecSimType_UINT8:
        GetNormalizeFunc<uint8_t>()(blob, dim);
        break;
    default:
        break;
    }
}

// CPU-baseline timer: nthreads worker threads pull queries from a shared counter and run the
// reference's single-threaded topKQuery on a frozen index (models WORKERS=nthreads, the only
// parallelism the reference has on this path: brute_force.h:243-291 is one thread per query).
// Returns wall seconds for all nq queries.  ids/scores (nq*k, may be NULL) receive the results.
double Ref_TimeTopK(void *idx, const void *queries, size_t qstride, size_t nq, size_t k,
                    int nthreads, size_t *ids, double *scores) {
    auto *index = (VecSimIndex *)idx;
    std::atomic<size_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
        th.emplace_back([&] {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= nq) break;
                auto *rep = index->topKQuery((const char *)queries + i * qstride, k, nullptr);
                for (size_t j = 0; j < rep->results.size() && j < k; j++) {
                    if (ids) ids[i * k + j] = rep->results[j].id;
                    if (scores) scores[i * k + j] = rep->results[j].score;
                }
                delete rep;
            }
        });
    for (auto &t : th) t.join();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

// Streaming form of BruteForceIndex::topKQuery (brute_force.h:243-291) over a FLAT array of stored-form rows that
// the caller hands over chunk by chunk (bench.py copies the device's own corpus back 1M rows at a time): the
// reference's dispatched distance kernel (spaces::GetDistFunc, same call as calcDistance) and the reference's heap
// class with its admission rule, per query, `nthreads` queries at a time.  The heap state travels between chunks in
// ids/scores/counts (counts[i] entries valid, any order); labels are label0 + row index.  `queries` are stored-form
// (already normalised for cosine).  After the last chunk the caller sorts by (score, label).
void Ref_ScanTopKChunk(int type, int metric, size_t dim, const void *rows, size_t stride, size_t n, size_t label0,
                       const void *queries, size_t qstride, size_t nq, size_t k, int nthreads, size_t *ids,
                       float *scores, size_t *counts) {
    using namespace spaces;
    unsigned char al = 0;
    auto m = (VecSimMetric)metric;
    dist_func_t<float> fn = nullptr;
    switch ((VecSimType)type) {
    case VecSimType_FLOAT32: fn = GetDistFunc<float, float>(m, dim, &al); break;
    case VecSimType_FLOAT16: fn = GetDistFunc<vecsim_types::float16, float>(m, dim, &al); break;
    case VecSimType_BFLOAT16: fn = GetDistFunc<vecsim_types::bfloat16, float>(m, dim, &al); break;
    case VecSimType_INT8: fn = GetDistFunc<int8_t, float>(m, dim, &al); break;
    case VecSimType_UINT8: fn = GetDistFunc<uint8_t, float>(m, dim, &al); break;
    default: return;
    }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < (nthreads < 1 ? 1 : nthreads); t++)
        th.emplace_back([&] {
            auto alloc = VecSimAllocator::newVecsimAllocator();
            for (;;) {
                const size_t qi = next.fetch_add(1);
                if (qi >= nq) break;
                const void *q = (const char *)queries + qi * qstride;
                vecsim_stl::max_priority_queue<float, labelType> heap(alloc);
                for (size_t j = 0; j < counts[qi]; j++) heap.emplace(scores[qi * k + j], ids[qi * k + j]);
                float upperBound = heap.size() ? heap.top().first : std::numeric_limits<float>::lowest();
                for (size_t r = 0; r < n; r++) {
                    const float score = fn((const char *)rows + r * stride, q, dim);
                    if (score < upperBound || heap.size() < k) { // brute_force.h:271-278
                        heap.emplace(score, (labelType)(label0 + r));
                        if (heap.size() > k) heap.pop();
                        upperBound = heap.top().first;
                    }
                }
                size_t c = 0;
                while (!heap.empty()) {
                    scores[qi * k + c] = heap.top().first;
                    ids[qi * k + c] = heap.top().second;
                    heap.pop();
                    c++;
                }
                counts[qi] = c;
            }
        });
    for (auto &t : th) t.join();
}

} // extern "C"
