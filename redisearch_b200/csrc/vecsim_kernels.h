// Host-visible launchers of the sm_100a KNN kernels (vecsim_kernels.cu).  Plain CUDA runtime
// types only; no torch.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rsb200 {

enum DType : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2, DT_I8 = 3, DT_U8 = 4 };
// MT_COS only differs from MT_IP for the integer types (division by the stored norms,
// VS/spaces/IP/IP.cpp:264-285); float corpora are normalised at ingest and use IP arithmetic
// (VS/spaces/spaces.cpp:50-61).
enum MetricKind : int { MT_L2 = 0, MT_IP = 1, MT_COS = 2 };

struct CorpusView {
    const void *rows;   // device, row-major, `pitch` bytes per row
    size_t pitch;       // >= stored row size, multiple of 16 for the 16/8-bit types
    uint32_t n_rows;
    uint32_t dim;
    DType dtype;
    MetricKind metric;
};

struct LaunchCounters {
    uint64_t launches = 0;
};

int device_sm_count();

// ---- fused scan + per-warp top-k (k <= kMaxFusedK) -------------------------------------------
// Shape of the candidate buffer the scan writes: lists_per_query lists of k composites per query.
struct ScanPlan {
    uint32_t grid_x, grid_y, wq, qt, lists_per_query;
    size_t smem_bytes;
    size_t cand_elems; // uint64 elements needed in d_cand
};
ScanPlan plan_scan_topk(const CorpusView &c, uint32_t nq, uint32_t k);
// d_queries: nq device blobs, qpitch bytes apart, already in stored form (normalised etc.).
cudaError_t launch_scan_topk(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq,
                             uint32_t k, const ScanPlan &plan, uint64_t *d_cand, cudaStream_t s,
                             LaunchCounters *ctr, const uint32_t *d_q_ok = nullptr, const uint32_t *d_abort = nullptr);
// out[q] = ok[q] ? a[q] : b[q] for [nq][k] composite arrays
cudaError_t launch_blend(const uint32_t *d_ok, const uint64_t *d_a, const uint64_t *d_b, uint32_t nq, uint32_t k, uint64_t *d_out,
                         cudaStream_t s, LaunchCounters *ctr);
// Reduce m_per_query candidate composites per query to the k smallest, ascending.
// d_out: [nq][k] composites (kEmptySlot-padded when fewer than k real candidates exist).
// d_nq_dev (nullable): only the first *d_nq_dev queries are processed (count known on the device only).
cudaError_t launch_final_select(const uint64_t *d_cand, uint32_t nq, uint32_t m_per_query, uint32_t k,
                                uint64_t *d_out, cudaStream_t s, LaunchCounters *ctr, const uint32_t *d_nq_dev = nullptr);

// ---- unfused path: all distances of one query, then cursor-select / range-compact ------------
cudaError_t launch_scan_scores(const CorpusView &c, const void *d_query, float *d_scores,
                               cudaStream_t s, LaunchCounters *ctr);
// k (<= kMaxFusedK) smallest composites strictly greater than *d_cursor (d_cursor may be NULL =
// no lower bound) among scores[0..n).  Writes lists to d_cand (size from plan_select_scores), to
// be reduced with launch_final_select.
uint32_t plan_select_scores_lists(uint32_t n);
cudaError_t launch_select_scores(const float *d_scores, uint32_t n, const uint64_t *d_cursor, uint32_t k,
                                 uint64_t *d_cand, cudaStream_t s, LaunchCounters *ctr);
// Append composite(score,id) of every score <= radius to d_out (capacity n), count in *d_count.
cudaError_t launch_range_compact(const float *d_scores, uint32_t n, float radius, uint64_t *d_out,
                                 uint32_t *d_count, cudaStream_t s, LaunchCounters *ctr);

// ---- ad-hoc: distances of listed rows ----------------------------------------------------------
// d_ids[i] == 0xFFFFFFFF -> NaN.
cudaError_t launch_gather_distances(const CorpusView &c, const void *d_query, const uint32_t *d_ids,
                                    uint32_t count, float *d_out, cudaStream_t s, LaunchCounters *ctr);

// filter-set plumbing of the fused hybrid query: docId -> row id through a dense table; selected positions -> docIds
cudaError_t launch_map_labels(const uint32_t *d_labels, uint32_t n, const uint32_t *d_table, uint32_t table_size, uint32_t *d_ids,
                              cudaStream_t s, LaunchCounters *ctr);
cudaError_t launch_pick_labels(const uint64_t *d_comp, uint32_t k, const uint32_t *d_labels, uint32_t *d_out, cudaStream_t s,
                               LaunchCounters *ctr);

// ---- result unpacking / shard merge -------------------------------------------------------------
// composites [nq][k] + id->label table -> labels (int64, -1 for empty) and float scores.
cudaError_t launch_unpack_results(const uint64_t *d_comp, uint32_t nq, uint32_t k,
                                  const uint64_t *d_id_to_label, int64_t *d_labels, float *d_scores,
                                  cudaStream_t s, LaunchCounters *ctr);
// [G][nq][k] (score,label) -> [nq][k] by (score asc, label asc); label -1 = empty.
// score_stride / label_stride: elements between consecutive shards' arrays (0 = nq*k, i.e. dense [G][nq][k]).
cudaError_t launch_merge_shards(const float *d_scores, const int64_t *d_labels, uint32_t G, uint32_t nq,
                                uint32_t k, float *d_out_scores, int64_t *d_out_labels, cudaStream_t s,
                                LaunchCounters *ctr, size_t score_stride = 0, size_t label_stride = 0);

} // namespace rsb200
