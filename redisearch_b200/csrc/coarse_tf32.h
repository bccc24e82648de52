// Launchers of the tensor-core coarse pass (coarse_tf32.cu): tcgen05 TF32 GEMM with fused candidate
// selection, exact rescoring, and the completeness proof.  See the header comment of coarse_tf32.cu.
#pragma once
#include "vecsim_kernels.h"

namespace rsb200 {

constexpr uint32_t kCoarseKeep = 24;   // candidates kept per (CTA row range, query)
constexpr uint32_t kCoarseMaxK = 16;   // largest k served by the coarse path
// |approx - exact| bound for unit vectors under TF32 operand truncation (2 * 2^-10 relative per
// product, Cauchy-Schwarz over the dot product) plus accumulation slack.
constexpr float kCoarseEpsUnit = 2.5e-3f;

struct CoarsePlan {
    uint32_t grid_x, grid_y, num_kb, tiles, keep;
    uint32_t csize; // thread-block cluster size along y (1 = no multicast)
    size_t cand_elems; // uint64 per (query, list, keep)
    size_t smem_bytes;
};

bool coarse_supported(const CorpusView &c, uint32_t nq, uint32_t k);
CoarsePlan plan_coarse(const CorpusView &c, uint32_t nq);
cudaError_t launch_coarse(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, const CoarsePlan &p,
                          uint64_t *d_cand, cudaStream_t s);
cudaError_t launch_rescore(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, uint32_t per_query,
                           const uint64_t *d_cand, uint64_t *d_exact, cudaStream_t s);
cudaError_t launch_verify(const uint64_t *d_cand, const uint64_t *d_topk, uint32_t nq, uint32_t lists_per_query, uint32_t keep,
                          uint32_t k, float eps, uint32_t *d_ok, cudaStream_t s);

} // namespace rsb200
