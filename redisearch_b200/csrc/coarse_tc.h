// Launchers of the tensor-core coarse pass (coarse_tc.cu): tcgen05 GEMM (fp16 shadow rows or TF32 on the
// fp32 rows) with fused candidate selection, exact rescoring, and the completeness proof.  See the header
// comment of coarse_tc.cu.
#pragma once
#include "vecsim_kernels.h"

namespace rsb200 {

constexpr uint32_t kCoarseKeep = 24;       // candidates kept per (CTA row range, query), first tier
constexpr uint32_t kCoarseKeepWide = 128;  // second tier (queries the first proof left open) and first tier of k > 16
constexpr uint32_t kCoarseTier1MaxK = 16;  // largest k the 24-entry lists serve
constexpr uint32_t kCoarseMaxK = 128;      // largest k served by the coarse path
constexpr uint32_t kCoarseSampleSlices = 32; // minima the sample pass publishes per (query, row range)
constexpr uint32_t kCoarseFixedCapDirect = 256; // list capacity of the fixed-bound pass on 16-bit corpora (k up to 128)
constexpr uint32_t kCoarseFixedCap = 96;   // list capacity of the fixed-bound main pass (rows below the bound per row range)
// |approx - exact| bounds for unit vectors (Cauchy-Schwarz over the dot product: sum |a_i b_i| <= 1):
//  TF32: each operand truncated to 11 significant bits -> 2 * 2^-10 relative per product, + accumulation slack;
//  F16 : each operand rounded to nearest, 11 significant bits -> 2 * 2^-11 = 9.8e-4 per product; elements below
//        the fp16 normal range (6.1e-5) add at most 2^-25 * sum|b_i| <= 2^-25 * sqrt(dim) <= 1e-6; fp32
//        accumulation of <= 1024 terms adds < 1.3e-4.
constexpr float kCoarseEpsTF32 = 2.5e-3f;
constexpr float kCoarseEpsF16 = 1.2e-3f;

enum CoarseKind : int { CoarseTF32 = 0, CoarseF16 = 1, CoarseDirect16 = 2, CoarseDirect8 = 3 };
inline float coarse_eps(CoarseKind k) { return k == CoarseF16 ? kCoarseEpsF16 : kCoarseEpsTF32; }

struct CoarsePlan {
    CoarseKind kind;
    uint32_t grid_x, grid_y, num_kb, tiles, keep;
    uint32_t stages; // depth of the row-tile ring in shared memory
    uint32_t epl;    // candidate-list entries per lane of the compacting warp (3 or 8)
    uint32_t tile_stride; // 1 = every row tile; n = every n-th (the sample pass)
    int mode;        // CoarseF16: 0 adaptive top-`keep` lists, 1 fixed admission bound per query (main pass), 2 sample pass (slice minima)
    uint32_t csize;  // thread-block cluster size along y (query groups sharing multicast row tiles); 1 = none
    bool pair;       // the two query groups of a row range run as a CTA PAIR: tcgen05.mma.cta_group::2 (M = 256 across the two SMs of a
                     // TPC), each CTA loads only its half of every row tile — no multicast, half the operand traffic per SM
    size_t cand_elems; // uint64 per (query, list, keep)
    size_t scratch_elems; // uint64 of per-CTA candidate-list scratch (CoarseF16), 0 otherwise
    size_t smem_bytes;
};

// operands of the coarse GEMM in the element type of `kind`: corpus rows and the query batch
struct CoarseOperands {
    const void *rows;
    size_t pitch;
    const void *queries;
    size_t qpitch;
    int elem_variant; // CoarseDirect16: 1 = bfloat16 (else IEEE half); CoarseDirect8: 1 = int8 (else uint8)
    int int_cosine;   // CoarseDirect8: cosine (rows carry their fp32 norm after the payload) instead of inner product;
                      // CoarseF16: squared-L2 epilogue (needs the two arrays below)
    const float *row_norm2; // CoarseF16 / L2: |row|^2 per row (fp32 rows)
    const float *q_norm2;   //                 |q|^2 per query
};
bool coarse_supported(const CorpusView &c, uint32_t nq, uint32_t k, CoarseKind kind);
// keep_override != 0 (CoarseF16 only): candidates per list instead of the default for k
CoarsePlan plan_coarse(const CorpusView &c, uint32_t nq, CoarseKind kind, uint32_t k, uint32_t keep_override = 0, uint32_t tile_stride = 1,
                       int mode = 0);
// d_nq_dev (nullable): the number of live queries is read from device memory (min with nq); 0 = the kernel exits at once
cudaError_t launch_coarse(const CoarseOperands &o, uint32_t n_rows, uint32_t dim, uint32_t nq, const CoarsePlan &p, uint64_t *d_cand,
                          uint64_t *d_scratch, cudaStream_t s, const uint32_t *d_nq_dev = nullptr, const float *d_thr_fixed = nullptr,
                          uint32_t *d_overflow = nullptr);
// bound of the fixed pass from the sample pass's candidate lists: d_thr[q] = k-th smallest approximate distance + 2 eps;
// clears d_overflow[q]
cudaError_t launch_threshold(const uint64_t *d_cand, uint32_t nq, uint32_t lists_per_query, uint32_t keep, uint32_t k, float eps,
                             const float *d_q_norm2, float max_norm, uint32_t dim, int l2, float *d_thr, uint32_t *d_overflow, cudaStream_t s);
// fp32 rows [first, first+n) -> the tiled fp16 shadow copy read by the CoarseF16 kernel (layout in coarse_tc.cu);
// the buffer holds coarse_shadow_bytes(capacity_rows, dim) bytes
size_t coarse_shadow_bytes(uint32_t rows, uint32_t dim);
cudaError_t launch_to_f16_tiled(const void *src, size_t spitch, uint32_t dim, uint32_t first, uint32_t n, void *dst, cudaStream_t s);
// fp32 rows [first, first+n) -> row-major fp16 rows (the query batch; dim % 8 == 0)
cudaError_t launch_to_f16(const void *src, size_t spitch, uint32_t dim, uint32_t first, uint32_t n, void *dst, size_t dpitch,
                          cudaStream_t s);
// One CTA per query: exact rescoring of the candidates within 2 eps of the k-th best approximate distance, exact top-k
// (d_out[q][k], ascending, kEmptySlot padded) and the completeness proof (d_ok[q]).  d_q_norm2 == NULL: unit vectors,
// |approx - exact| <= eps; otherwise the bound scales with max_norm (over all rows) and |q| (refine_kernel).  Second tier:
// d_q_index[i] = query of the original batch whose lists sit at position i (d_q_norm2 is indexed by position), *d_nq_dev
// live positions.
cudaError_t launch_refine(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, uint32_t lists_per_query, uint32_t keep,
                          uint32_t k, const uint64_t *d_cand, float eps, const float *d_q_norm2, float max_norm, uint32_t *d_ok,
                          uint64_t *d_out, const uint32_t *d_q_index, const uint32_t *d_nq_dev, cudaStream_t s,
                          const float *d_thr_T = nullptr, const uint32_t *d_overflow = nullptr);
// direct 16-bit route: d_ok[q] = d_overflow[q] ? 0 : 1
cudaError_t launch_flags_from_overflow(const uint32_t *d_overflow, uint32_t nq, uint32_t *d_ok, cudaStream_t s);
// second tier of the direct route: row i of src ([.][k] composites) -> row d_idx[i] of dst, d_ok[d_idx[i]] = 2, for i < *d_count
cudaError_t launch_scatter_rows(const uint64_t *d_src, const uint32_t *d_idx, const uint32_t *d_count, uint32_t max_n, uint32_t k,
                                uint64_t *d_dst, uint32_t *d_ok, cudaStream_t s);
// d_idx[0, *d_count) = the queries with d_ok == 0, ascending
cudaError_t launch_compact_unproven(const uint32_t *d_ok, uint32_t nq, uint32_t *d_idx, uint32_t *d_count, cudaStream_t s);
// row i of dst = row d_idx[i] of src (pitch % 16 == 0), squared norms likewise (nullable), for i < *d_count
cudaError_t launch_gather_queries(const void *d_src, size_t pitch, const float *d_src_n2, const uint32_t *d_idx, const uint32_t *d_count,
                                  uint32_t max_n, void *d_dst, float *d_dst_n2, cudaStream_t s);
// |row|^2 of fp32 rows [first, first+n) into d_norm2[first..]; d_stats (nullable) = {max |row|^2, max |x|} as float bits
cudaError_t launch_row_stats(const void *rows, size_t pitch, uint32_t dim, uint32_t first, uint32_t n, float *d_norm2, uint32_t *d_stats,
                             cudaStream_t s);

} // namespace rsb200
