// Launchers of the posting-list kernels (ii_kernels.cu).  Plain CUDA runtime types only.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace rsb200 {

constexpr int kIIThreads = 256;
constexpr int kIIItems = 4;
constexpr int kIIChunk = kIIThreads * kIIItems; // entries of the driving list per CTA
constexpr int kIISmemElems = 8192;              // 32 KB window of the probed list staged per CTA
constexpr int kIIMaxLists = 16;

struct IntersectArgs {
    const uint32_t *ids[kIIMaxLists]; // [0] = the shortest list (drives), others ascending by length
    uint32_t len[kIIMaxLists];
    uint32_t n;
    uint32_t *tmp_idx;  // [nchunks*kIIChunk] survivors of chunk c at tmp_idx[c*kIIChunk + r] (index into list 0)
    uint32_t *tmp_pos;  // [n][stride]: position of entry idx of list 0 inside list j (valid for survivors)
    uint32_t *counts;   // [nchunks]
    size_t stride;      // nchunks*kIIChunk
};

struct GatherArgs {
    const uint32_t *ids0;
    const uint32_t *freqs[kIIMaxLists];
    uint32_t n;
    const uint32_t *tmp_idx, *tmp_pos, *counts, *offsets;
    size_t stride;   // of tmp_pos
    uint32_t *out_doc;
    uint32_t *out_freq; // [n][fstride]
    size_t fstride;
};

struct ScoreArgs {
    int scorer; // II_Scorer numbering
    int is_union;
    uint32_t n_children;
    double weight[kIIMaxLists], idf[kIIMaxLists], bm25_idf[kIIMaxLists]; // aggregate child order
    double agg_weight, avg_doc_len, min_score;
    uint64_t tanh_factor;
    const uint32_t *doc_len;   // by docId, may be NULL
    const float *doc_score;    // by docId, may be NULL
    const uint32_t *max_freq;  // by docId, may be NULL
};

cudaError_t ii_launch_decode(const uint8_t *d_bytes, const uint64_t *d_byte_off, const uint64_t *d_first_id,
                             const uint32_t *d_entry_off, uint32_t nblocks, int codec, uint32_t *d_ids, uint32_t *d_freqs,
                             uint32_t *d_masks, cudaStream_t s);
cudaError_t ii_launch_mask_filter(const uint32_t *d_ids, const uint32_t *d_freqs, const uint32_t *d_masks, uint32_t n,
                                  uint32_t filter, uint32_t *d_counts, uint32_t *d_offsets, uint32_t *d_total,
                                  uint32_t *d_out_ids, uint32_t *d_out_freqs, cudaStream_t s);
cudaError_t ii_launch_intersect(const IntersectArgs &a, uint32_t nchunks, uint32_t *d_offsets, uint32_t *d_total,
                                cudaStream_t s);
cudaError_t ii_launch_gather(const GatherArgs &g, uint32_t nchunks, cudaStream_t s);
cudaError_t ii_launch_score(const ScoreArgs &sa, const uint32_t *d_docs, const uint32_t *d_freqs, size_t fstride,
                            const uint32_t *d_len, uint32_t cap_len, double *d_scores, cudaStream_t s);
cudaError_t ii_launch_union(const uint32_t *const *d_ids, const uint32_t *const *d_freqs, const uint32_t *lens, uint32_t n,
                            uint32_t nwords, uint32_t *d_bitmap, uint32_t *d_blocksum, uint32_t *d_blockoff,
                            uint32_t *d_wordoff, uint32_t *d_total, uint32_t *d_out_doc, uint32_t *d_out_freq,
                            size_t fstride, bool want_freqs, cudaStream_t s);
uint32_t ii_topn_lists(uint32_t m);
cudaError_t ii_launch_topn(const uint32_t *d_docs, const double *d_scores, const uint32_t *d_len, uint32_t cap_len, uint32_t k,
                           uint64_t *d_keys, uint32_t *d_ids, cudaStream_t s);

} // namespace rsb200
