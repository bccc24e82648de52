// Host-side FLAT index object behind the VecSim C API: label<->row bookkeeping, preprocessing,
// staging of appended rows, query contexts (stream + scratch), reply objects.  The vectors
// themselves live only in HBM.  Mirrors the host logic of
//   VS/algorithms/brute_force/brute_force.h            (append / swap-delete / queries / heuristics)
//   VS/algorithms/brute_force/brute_force_single.h     (label -> id, update in place)
//   VS/algorithms/brute_force/brute_force_multi.h      (label -> ids)
//   VS/algorithms/brute_force/bf_batch_iterator.h      (batch iterator state machine)
#pragma once
#include "../../include/vecsim_b200.h"
#include "vecsim_kernels.h"

#include <atomic>
#include <map>
#include <memory>
#include <limits>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

struct VecSimQueryResult {
    size_t id;
    double score;
};
struct VecSimQueryReply {
    std::vector<VecSimQueryResult> results;
    VecSimQueryReply_Code code = VecSim_QueryReply_OK;
};
struct VecSimQueryReply_Iterator {
    VecSimQueryReply *reply;
    size_t pos;
};
struct VecSimDebugInfoIterator {
    std::vector<VecSim_InfoField> fields;
    std::vector<std::string> owned;
    size_t pos = 0;
};

namespace rsb200 {

struct Globals {
    std::atomic<timeoutCallbackFunction> timeout_cb{nullptr};
    std::atomic<logCallbackFunction> log_cb{nullptr};
    VecSimMemoryFunctions mem{};
    std::mutex test_ctx_mu; // VecSim_SetTestLogContext
    std::string test_name, test_type;
};
Globals &globals();
void set_coarse_mode(int mode); // -1 env default, 0 exact scans only, 1 coarse pass on fp16 shadow rows, 2 TF32 coarse pass

// Device + pinned scratch for one in-flight query (or query batch).  Checked out of a pool so
// that many RediSearch worker threads can query one index concurrently (SURVEY.md §8b threading).
struct QueryCtx {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    uint8_t *d_query = nullptr, *h_query = nullptr; // staged query blobs
    size_t query_cap = 0;
    uint64_t *d_cand = nullptr;
    size_t cand_cap = 0;
    uint64_t *d_out = nullptr, *h_out = nullptr;
    size_t out_cap = 0;
    float *d_scores = nullptr; // unfused path: one score per row
    size_t scores_cap = 0;
    uint32_t *d_count = nullptr, *h_count = nullptr;
    uint32_t *d_last_ok = nullptr; // coarse path: per-query verification flags of the last batch (inside d_cand)
    uint32_t last_ok_n = 0;
    uint32_t *d_ids = nullptr, *h_ids = nullptr;
    float *d_dist = nullptr, *h_dist = nullptr;
    size_t ids_cap = 0;
    bool abandoned = false; // a timed-out caller left while kernels were still running on `stream`: synchronise before reuse
    // raised by the host when a caller times out: the exact-scan kernel polls it (mapped pinned memory) and winds down, so the GPU
    // does not finish a pass nobody waits for
    uint32_t *h_abort = nullptr;
    const uint32_t *d_abort = nullptr;
    ~QueryCtx();
    bool init();
    bool need_query(size_t bytes);
    bool need_cand(size_t elems);
    bool need_out(size_t elems);
    bool need_scores(size_t n);
    bool need_ids(size_t n);
};

class FlatIndex;

struct BatchIter {
    FlatIndex *index;
    std::vector<uint8_t> query; // stored-form query blob (normalised for cosine)
    void *timeout_ctx;
    std::unique_ptr<QueryCtx> ctx; // owns the score array between Next calls
    bool scored = false;
    uint32_t n_rows = 0;        // rows at scoring time
    size_t label_count = 0;     // labels at scoring time (bf_batch_iterator.h:29)
    size_t returned = 0;
    bool has_cursor = false;
    uint64_t cursor = 0;                  // last composite handed out
    std::unordered_set<size_t> seen;      // multi-value: labels already returned
    std::vector<size_t> id_to_label_snap; // label table at scoring time
};

struct AdhocCtx {
    FlatIndex *index;
    std::vector<uint8_t> query; // stored-form query
    std::unique_ptr<QueryCtx> ctx;
    bool query_on_device = false;
};

class FlatIndex {
  public:
    static FlatIndex *create(const BFParams &p, void *log_ctx);
    ~FlatIndex();

    int add(const void *blob, size_t label);
    int add_bulk(const void *blobs, size_t stride, size_t n, const size_t *labels, size_t label0);
    int add_bulk_device(const void *d_rows, size_t n, size_t label0);
    int remove(size_t label);
    size_t size() const { return count_; }
    size_t label_count() const { return multi_ ? label_to_ids_.size() : label_to_id_.size(); }
    bool reserve(size_t rows);
    bool flush();

    VecSimQueryReply *topk(const void *q, size_t k, VecSimQueryParams *qp, VecSimQueryReply_Order order);
    // the same through the request combiner (micro_batcher.h): concurrent callers share one corpus pass.  Opt-in:
    // VECSIM_B200_MICROBATCH_US=<collection window in microseconds> (default 0 = off).
    VecSimQueryReply *topk_combined(const void *q, size_t k, VecSimQueryParams *qp, VecSimQueryReply_Order order);
    static int microbatch_window_us();
    int topk_batch(const void *qs, size_t qstride, size_t nq, size_t k, VecSimQueryParams *qp, size_t *out_labels,
                   double *out_scores);
    int topk_batch_device(const void *d_q, size_t nq, size_t k, int64_t *d_labels, float *d_scores, cudaStream_t s);
    VecSimQueryReply *range(const void *q, double radius, VecSimQueryParams *qp, VecSimQueryReply_Order order);
    double distance_from(size_t label, const void *stored_form_blob);
    bool prefer_adhoc(size_t subset, size_t k, bool initial);

    BatchIter *batch_new(const void *q, VecSimQueryParams *qp);
    VecSimQueryReply *batch_next(BatchIter *it, size_t n, VecSimQueryReply_Order order);

    AdhocCtx *adhoc_new(const void *q);
    void adhoc_distances(AdhocCtx *c, const size_t *labels, double *out, size_t n);
    // fused hybrid ad-hoc query: k nearest among the rows whose labels are listed (ascending) in doc_ids
    int topk_filtered(const void *q, size_t k, const uint32_t *doc_ids, size_t n, bool ids_on_device, size_t *out_labels,
                      double *out_scores, size_t *out_count);
    int topk_filtered_batch(const void *const *queries, size_t nq, size_t k, const uint32_t *const *d_doc_ids, const size_t *counts,
                            size_t *out_labels, double *out_scores, size_t *out_counts);

    VecSimIndexBasicInfo basic_info() const;
    VecSimIndexStatsInfo stats_info() const;
    VecSimIndexDebugInfo debug_info() const; // BruteForceIndex::debugInfo, brute_force.h:318-325
    VecSimDebugInfoIterator *debug_iterator() const;
    void set_last_mode(VecSearchMode m) { last_mode_ = m; }
    VecSimB200_Stats get_stats(bool reset);
    // debug: verification flags of the last device-batch call (1 = answered by the tensor-core path)
    int last_coarse_flags(uint32_t *out, size_t n);
    const void *device_rows(size_t *pitch, size_t *rows) {
        flush();
        *pitch = pitch_;
        *rows = count_;
        return d_rows_;
    }

    // stored-form rows [first, first + n) -> tightly packed host buffer (stored_bytes_ per row); false if out of range
    bool read_rows(size_t first, size_t n, void *host_dst);
    size_t query_blob_bytes() const { return stored_bytes_; } // VecSimParams_GetQueryBlobSize
    void preprocess_query(const void *blob, uint8_t *dst) const;   // -> stored form
    void preprocess_storage(const void *blob, uint8_t *dst) const; // -> stored form

    VecSimType type_;
    VecSimMetric metric_;
    size_t dim_;
    bool multi_;
    size_t block_size_;
    void *log_ctx_;

    void log(const char *level, const char *fmt, ...) const;

  private:
    FlatIndex() = default;
    CorpusView view() const;
    std::unique_ptr<QueryCtx> checkout();
    void checkin(std::unique_ptr<QueryCtx> c);
    bool grow_to(size_t rows);
    bool sync_labels_to_device();
    bool timed_out(void *ctx) const;
    // Wait for `s`, polling the timeout callback (VecSim_SetTimeoutCallbackFunction) every ~50 us while kernels run —
    // the reference checks it per vector (brute_force.h:265-269); a device pass cannot be interrupted, but the caller
    // is released as soon as the deadline passes.  0 = done, 1 = timed out (the stream is still busy), -1 = CUDA error.
    int wait_polling(cudaStream_t s, void *timeout_ctx) const;
    // k smallest composites (> cursor) over ctx->d_scores[0..n) into ctx->h_out, chunked by
    // kMaxFusedK; returns the number found or -1.
    long select_from_scores(QueryCtx &c, uint32_t n, bool has_cursor, uint64_t cursor, size_t want);
    bool upload_query(QueryCtx &c, const uint8_t *stored_q, size_t nq);
    bool batch_scan(QueryCtx &c, const void *d_q, size_t qpitch, uint32_t nq, uint32_t ke, cudaStream_t st, LaunchCounters &lc,
                    uint64_t **d_result);
    void finish_reply(VecSimQueryReply *rep, VecSimQueryReply_Order order) const;

    DType dtype_;
    MetricKind mkind_;
    size_t elem_bytes_ = 0;   // sizeof element
    size_t stored_bytes_ = 0; // VS/utils/vec_utils.cpp:296-302
    size_t pitch_ = 0;        // HBM row pitch (>= stored_bytes_)

    uint8_t *d_rows_ = nullptr;
    // fp16 copy of the fp32 rows for the tensor-core coarse pass, stored tile by tile in the swizzled layout
    // the kernel streams (coarse_tc.cu); built lazily by the first eligible batch, rows [0, shadow_rows_)
    // valid except shadow_dirty_
    uint8_t *d_shadow_ = nullptr;
    size_t shadow_cap_ = 0, shadow_rows_ = 0;
    // L2 / raw inner-product indexes: |row|^2 per row and the running maxima the error bound needs
    float *d_norm2_ = nullptr;
    uint32_t *d_stats_ = nullptr;
    float shadow_max_norm_ = 0.0f, shadow_max_abs_ = std::numeric_limits<float>::infinity();
    std::atomic<bool> coarse_disabled_{false}; // L2 / IP index with values outside the fp16 range: exact scans only
    // cosine index that took an in-place update (brute_force_single.h:139-144 stores the caller's RAW blob): its rows are
    // no longer all unit vectors, the coarse proof uses the norm-scaled bound from then on
    std::atomic<bool> raw_rows_{false};
    bool unit_rows() const { return metric_ == VecSimMetric_Cosine && !raw_rows_; }
    std::vector<idType> shadow_dirty_;
    bool ensure_shadow(cudaStream_t st);
    bool single_query_takes_coarse(uint32_t ke);
    size_t capacity_ = 0; // rows of HBM allocated
    size_t count_ = 0;    // rows in the index (incl. staged)
    size_t resident_ = 0; // rows already copied to HBM
    uint8_t *h_stage_ = nullptr; // pinned; rows [resident_, count_)
    size_t stage_cap_rows_ = 0;
    cudaStream_t copy_stream_ = nullptr;

    std::vector<size_t> id_to_label_;
    std::unordered_map<size_t, idType> label_to_id_;
    std::unordered_map<size_t, std::vector<idType>> label_to_ids_;
    uint64_t *d_id_to_label_ = nullptr;
    size_t d_labels_cap_ = 0;
    bool labels_dirty_ = true;
    // dense label -> row id table on the device (labels are RediSearch docIds: small integers), for topk_filtered
    uint32_t *d_label_to_id_ = nullptr;
    size_t l2i_size_ = 0, l2i_cap_ = 0;
    bool l2i_dirty_ = true;
    bool sync_label_table();

    mutable std::mutex mu_;      // guards mutation + staging
    std::mutex pool_mu_;
    std::vector<std::unique_ptr<QueryCtx>> pool_;
    mutable VecSearchMode last_mode_ = EMPTY_MODE;

    std::atomic<uint64_t> launches_total_{0};
    std::atomic<uint64_t> coarse_batches_{0};
    bool last_batch_coarse_ = false;
    int last_batch_path_ = 0; // 0 exact scan, 1 tensor-core coarse pass + proof, 2 tensor-core direct (16-bit corpora)
  public:
    int last_batch_path() const { return last_batch_path_; }
  private:
    std::unique_ptr<QueryCtx> dev_ctx_; // scratch of topk_batch_device (stream-ordered)
    std::mutex dev_mu_;
    bool dev_timing_pending_ = false;
    uint64_t dev_timing_bytes_ = 0;
    void collect_dev_timing_locked();
    std::atomic<uint64_t> scan_launches_{0};
    std::atomic<uint64_t> scan_bytes_{0};
    double scan_us_ = 0;
    std::mutex stats_mu_;
    struct TopkReq;
    std::mutex mb_mu_;
    std::map<size_t, std::shared_ptr<void>> batchers_; // one MicroBatcher<TopkReq> per k
    friend struct BatchIter;
};

} // namespace rsb200
