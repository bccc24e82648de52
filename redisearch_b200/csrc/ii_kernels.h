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
constexpr int kIIMaxLists = 32; // children of one AND / OR (the reference switches its union to the heap variant above 20: same docIds)

struct IntersectArgs {
    const uint32_t *ids[kIIMaxLists]; // [0] = the shortest list (drives), others ascending by length
    uint32_t len[kIIMaxLists];
    uint8_t mode[kIIMaxLists]; // per list (slot 0 = the driver, always required): 0 = required (AND), 1 = NOT (the docId must be
                               // absent: not.rs as a child of an intersection), 2 = OPTIONAL (never rejects: optional.rs)
    uint32_t n;
    uint32_t *tmp_idx;  // [nchunks*kIIChunk] survivors of chunk c at tmp_idx[c*kIIChunk + r] (index into list 0)
    uint32_t *tmp_pos;  // [n][stride]: position of entry idx of list 0 inside list j (valid for survivors)
    uint32_t *counts;   // [nchunks]
    size_t stride;      // nchunks*kIIChunk
};

struct GatherArgs {
    const uint32_t *ids0;
    const uint32_t *freqs[kIIMaxLists];
    uint8_t mode[kIIMaxLists]; // as IntersectArgs: NOT children yield freq 0, OPTIONAL children freq 0 where absent (virtual results)
    uint32_t n;
    const uint32_t *tmp_idx, *tmp_pos, *counts, *offsets;
    size_t stride;   // of tmp_pos
    uint32_t *out_doc;
    uint32_t *out_freq; // [n][fstride]
    uint32_t *out_pos;  // nullable, [n][fstride]: position of the hit inside list j (phrase checks read the term positions there)
    size_t fstride;
};

// phrase constraints (slop / in-order): the term positions of every hit, children in aggregate order
constexpr int kPhraseMaxLists = 8;
struct PhraseArgs {
    const uint8_t *bytes[kPhraseMaxLists];    // gathered block bytes of list j (the offsets payloads live inside)
    const uint32_t *off_pos[kPhraseMaxLists]; // per posting: start of its offsets payload in bytes[j]
    const uint32_t *off_len[kPhraseMaxLists]; // per posting: length (0 / NULL array = the child carries no offsets)
    const uint32_t *pos;                      // [n][fstride] posting position of hit o, rows in KERNEL-slot order (GatherArgs::out_pos)
    uint32_t row[kPhraseMaxLists];            // aggregate child j -> its row of pos
    size_t fstride;
    uint32_t n;
    uint32_t max_slop; // 0xFFFFFFFF = no limit
    int in_order;
};
cudaError_t ii_launch_phrase_filter(const PhraseArgs &a, const uint32_t *d_len, uint32_t cap_len, uint32_t *d_flags, uint32_t *d_counts,
                                    uint32_t *d_offsets, uint32_t *d_total, const uint32_t *d_docs, const uint32_t *d_freqs, size_t fstride,
                                    uint32_t *d_out_docs, uint32_t *d_out_freqs, size_t out_fstride, cudaStream_t s);

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

// ---- fused batch search: AND + scorer + top-N of MANY queries in two launches --------------------------------------
constexpr int kFusedMaxLists = 8;   // more children: the per-query kernel chain
constexpr int kFusedMaxTopN = 128;
// One query of the batch.  Children are in the reference's aggregate order (ascending num_estimated, stable:
// RS/rqe_iterators/src/intersection.rs:110-145), which for the fused path is also ascending ACTUAL length, so child 0 drives.
struct FusedQuery {
    const uint32_t *ids[kFusedMaxLists];
    const uint32_t *freqs[kFusedMaxLists];
    uint32_t len[kFusedMaxLists];
    double weight[kFusedMaxLists], idf[kFusedMaxLists], bm25_idf[kFusedMaxLists];
    uint32_t n;       // children (1..kFusedMaxLists)
    uint32_t item0;   // first work item (1024-entry chunk of child 0) of this query in the batch
    uint32_t nchunks; // work items of this query
    uint32_t _pad;
};
struct FusedCommon {
    int scorer;
    double agg_weight, avg_doc_len;
    uint64_t tanh_factor;
    const uint32_t *doc_len;
    const float *doc_score;
    const uint32_t *max_freq;
};
// Launch 1: one CTA per work item — membership of the chunk's docIds in every other child (window located by two warp-wide
// searches, staged in shared memory), freqs of the matches, the scorer, and the CTA's best `top_n` hits into
// cand_keys / cand_ids [item][top_n] (padded with ~0); hits[q] += survivors.  Launch 2: one CTA per query selects the best
// top_n by (score desc, docId asc) over its items' candidates into out_keys / out_ids [nq][top_n].
cudaError_t ii_launch_fused_search(const FusedQuery *d_queries, uint32_t nq, uint32_t total_items, const FusedCommon &fc, uint32_t top_n,
                                   uint64_t *d_cand_keys, uint32_t *d_cand_ids, uint32_t *d_hits, uint64_t *d_out_keys,
                                   uint32_t *d_out_ids, cudaStream_t s);

cudaError_t ii_launch_decode(const uint8_t *d_bytes, const uint64_t *d_byte_off, const uint64_t *d_first_id,
                             const uint32_t *d_entry_off, uint32_t nblocks, int codec, uint32_t *d_ids, uint32_t *d_freqs,
                             uint32_t *d_masks, cudaStream_t s);
// batch decode: 32-bit tables (byte_off[nblocks+1] into the 16-byte-aligned, 16-byte-padded gathered stream, first_id[nblocks],
// entry_off[nblocks+1] into the output arrays), blocks of MANY lists back to back
cudaError_t ii_launch_decode_staged(const uint8_t *d_bytes, const uint32_t *d_byte_off, const uint32_t *d_first_id,
                                    const uint32_t *d_entry_off, uint32_t nblocks, int codec, uint32_t *d_ids, uint32_t *d_freqs,
                                    uint32_t *d_masks, uint32_t *d_off_pos, uint32_t *d_off_len, cudaStream_t s);
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
