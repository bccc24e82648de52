// Bench / test tooling (NOT part of the drop-in surface): counter-based synthetic corpora generated
// directly in HBM, so that a 10M x 768 fp32 corpus (30.7 GB) never has to cross PCIe, plus the
// reference's cosine normaliser mirrored bit for bit on device
// (VS/spaces/normalize/normalize_naive.h:23-37: double sum in element order, norm rounded to float,
// float division).  The generator is the same 64-bit mix as oracle/vecsim_oracle.c orc_synth_f32, so
// host and device can regenerate any element.
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t seed, uint64_t a, uint64_t b) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (a + 1) + 0xD1B54A32D192ED03ULL * (b + 1);
    z ^= z >> 30;
    z *= 0xBF58476D1CE4E5B9ULL;
    z ^= z >> 27;
    z *= 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z;
}
__host__ __device__ inline float synth_f32(uint64_t seed, uint64_t row, uint64_t col) {
    const uint32_t m = (uint32_t)(mix64(seed, row, col) >> 40);
    return (float)m * (2.0f / 16777216.0f) - 1.0f; // exact: 24-bit integer scaled by a power of two
}

// fp32 -> fp16 exactly as VS/types/float16.h:62-117 (restated in host_numeric.h)
__device__ inline uint16_t f32_to_half_ref(float x) {
    uint32_t u = __float_as_uint(x);
    const uint32_t sign = u & 0x80000000u;
    u ^= sign;
    const uint32_t inf = 255u << 23;
    uint32_t o = (u > inf) ? 0x7e00u : 0x7c00u;
    const uint32_t keep = ~0xfffu;
    float scaled = __fmul_rn(__uint_as_float(u & keep), __uint_as_float(15u << 23));
    const float cap = __uint_as_float((31u << 23) - 0x1000u);
    if (cap < scaled) scaled = cap;
    const int32_t shifted = (int32_t)__float_as_uint(scaled) - (int32_t)keep;
    if (u < inf) o = (uint32_t)(shifted >> 13);
    return (uint16_t)(o | (sign >> 16));
}
__device__ inline uint16_t f32_to_bf16_ref(float x) {
    uint32_t u = __float_as_uint(x);
    u += ((u >> 16) & 1u) + 0x7FFFu;
    return (uint16_t)(u >> 16);
}

// type codes == VecSimType
__global__ void fill_rows_kernel(uint8_t *rows, size_t pitch, int type, uint64_t seed, uint64_t row0, uint64_t nrows,
                                 uint32_t dim) {
    const uint64_t total = nrows * dim;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = i / dim;
        const uint32_t c = (uint32_t)(i - r * dim);
        const float x = synth_f32(seed, row0 + r, c);
        uint8_t *p = rows + r * pitch;
        switch (type) {
        case 0: reinterpret_cast<float *>(p)[c] = x; break;
        case 3: reinterpret_cast<uint16_t *>(p)[c] = f32_to_half_ref(x); break;
        case 2: reinterpret_cast<uint16_t *>(p)[c] = f32_to_bf16_ref(x); break;
        case 4: reinterpret_cast<int8_t *>(p)[c] = (int8_t)__float2int_rn(127.0f * x); break;
        default: p[c] = (uint8_t)__float2int_rn(__fmaf_rn(127.5f, x, 127.5f)); break;
        }
    }
}

// one thread per row: sequential double accumulation in element order, like the host code
__global__ void normalize_f32_rows_kernel(uint8_t *rows, size_t pitch, uint64_t nrows, uint32_t dim) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    float *v = reinterpret_cast<float *>(rows + r * pitch);
    double s = 0.0;
    for (uint32_t i = 0; i < dim; i++) {
        const double d = (double)v[i];
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    const float norm = (float)sqrt(s);
    for (uint32_t i = 0; i < dim; i++) v[i] = __fdiv_rn(v[i], norm);
}

} // namespace

extern "C" {

// Fill nrows x dim synthetic elements of VecSimType `type` at d_rows (row pitch in bytes).
int Synth_FillRows(void *d_rows, size_t pitch, int type, uint64_t seed, uint64_t row0, uint64_t nrows, uint32_t dim,
                   void *stream) {
    if (nrows == 0) return 0;
    const uint64_t total = nrows * dim;
    const unsigned grid = (unsigned)((total + 255) / 256 < 148u * 32u ? (total + 255) / 256 : 148u * 32u);
    fill_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t *)d_rows, pitch, type, seed, row0, nrows, dim);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// In-place cosine normalisation of fp32 rows with the reference's arithmetic.
int Synth_NormalizeRowsF32(void *d_rows, size_t pitch, uint64_t nrows, uint32_t dim, void *stream) {
    if (nrows == 0) return 0;
    normalize_f32_rows_kernel<<<(unsigned)((nrows + 127) / 128), 128, 0, (cudaStream_t)stream>>>((uint8_t *)d_rows, pitch,
                                                                                                nrows, dim);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

float Synth_ElementF32(uint64_t seed, uint64_t row, uint64_t col) { return synth_f32(seed, row, col); }

} // extern "C"

// ================================================================================================
// Synthetic Zipf postings, same arithmetic as oracle/postings_oracle.c
// orc_synth_member / orc_synth_doclen so host and device agree entry for entry.
// ================================================================================================
namespace {

__device__ __forceinline__ bool synth_member(uint64_t rank, uint64_t doc, uint64_t thresh, uint32_t *tf) {
    const uint64_t h = mix64(7, rank, doc);
    if ((h >> 32) >= thresh) return false;
    uint32_t low = (uint32_t)h, g = 0;
    while (g < 31 && (low & 1u)) {
        g++;
        low >>= 1;
    }
    *tf = 1 + g;
    return true;
}

// docs 1..n_docs in chunks of 1024: count members per chunk
__global__ void postings_count_kernel(uint64_t n_docs, uint64_t rank, uint64_t thresh, uint32_t *counts) {
    __shared__ uint32_t s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * 1024u + 1;
    uint32_t c = 0, tf;
    for (uint32_t i = threadIdx.x; i < 1024u; i += blockDim.x)
        if (base + i <= n_docs && synth_member(rank, base + i, thresh, &tf)) c++;
    atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) counts[blockIdx.x] = s_cnt;
}
__global__ void __launch_bounds__(1024) synth_scan_kernel(const uint32_t *counts, uint32_t n, uint32_t *offsets, uint32_t *total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = (i < n) ? counts[i] : 0;
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        uint32_t wb = 0;
        for (int w = 0; w < warp; w++) wb += s_warp[w];
        const uint32_t carry = s_carry;
        if (i < n) offsets[i] = carry + wb + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + wb + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
// one warp per 1024-doc chunk, ordered scatter (32 docs per step, ballot ranks)
__global__ void postings_scatter_kernel(uint64_t n_docs, uint64_t rank, uint64_t thresh, const uint32_t *offsets,
                                        uint32_t *ids, uint32_t *freqs) {
    const uint64_t base = (uint64_t)blockIdx.x * 1024u + 1;
    uint32_t o = offsets[blockIdx.x];
    const int lane = threadIdx.x;
    for (uint32_t s = 0; s < 1024u; s += 32) {
        const uint64_t doc = base + s + lane;
        uint32_t tf = 0;
        const bool m = doc <= n_docs && synth_member(rank, doc, thresh, &tf);
        const unsigned ball = __ballot_sync(0xffffffffu, m);
        if (m) {
            const uint32_t r = __popc(ball & ((1u << lane) - 1u));
            ids[o + r] = (uint32_t)doc;
            freqs[o + r] = tf;
        }
        o += __popc(ball);
    }
}
__global__ void doclen_kernel(uint64_t n_docs, uint32_t *out) {
    for (uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; d <= n_docs; d += (uint64_t)gridDim.x * blockDim.x)
        out[d] = 50u + (uint32_t)(mix64(11, d, 0) % 451u);
}

} // namespace

extern "C" {

uint64_t Synth_DocFreq(uint64_t n_docs, uint64_t rank) {
    uint64_t df = (uint64_t)((double)n_docs * 0.2 / (double)rank);
    return df > n_docs ? n_docs : df;
}

// Posting list of vocabulary rank `rank` over docs 1..n_docs, ascending, into d_ids/d_freqs
// (capacity >= Synth_DocFreq*1.2 + 4096).  d_scratch: 2*(n_docs/1024+1) uint32.  Count -> *h_count.
int Synth_Postings(uint64_t n_docs, uint64_t rank, uint32_t *d_ids, uint32_t *d_freqs, uint32_t *d_scratch,
                   uint32_t *d_total, uint32_t *h_count, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t chunks = (uint32_t)((n_docs + 1023) / 1024);
    const uint64_t thresh = (Synth_DocFreq(n_docs, rank) << 32) / n_docs;
    uint32_t *counts = d_scratch, *offsets = d_scratch + chunks;
    postings_count_kernel<<<chunks, 256, 0, s>>>(n_docs, rank, thresh, counts);
    synth_scan_kernel<<<1, 1024, 0, s>>>(counts, chunks, offsets, d_total);
    postings_scatter_kernel<<<chunks, 32, 0, s>>>(n_docs, rank, thresh, offsets, d_ids, d_freqs);
    if (cudaMemcpyAsync(h_count, d_total, 4, cudaMemcpyDeviceToHost, s) != cudaSuccess) return -1;
    return cudaStreamSynchronize(s) == cudaSuccess ? 0 : -1;
}

int Synth_DocLens(uint64_t n_docs, uint32_t *d_out, void *stream) {
    doclen_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(n_docs, d_out);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// Host-side encoder producing the reference's FreqsOnly IndexBlocks (qint2[delta,freq], 100 entries per
// block, first record delta 0: RS/inverted_index/src/codec/freqs_only.rs:23-45, index/core.rs:235-358)
// from decoded arrays — the INPUT format the drop-in receives from RediSearch.  Returns bytes written;
// block tables sized n/100+1.
size_t Synth_EncodeFreqsOnlyBlocks(const uint32_t *ids, const uint32_t *freqs, size_t n, uint8_t *out, uint64_t *blk_first,
                                   uint64_t *blk_last, uint16_t *blk_n, uint64_t *blk_off, size_t *nblocks) {
    size_t pos = 0, nb = 0;
    uint32_t last = 0;
    for (size_t i = 0; i < n; i++) {
        if (i % 100 == 0) {
            blk_first[nb] = ids[i];
            blk_off[nb] = pos;
            blk_n[nb] = 0;
            last = ids[i];
            nb++;
        }
        const uint32_t vals[2] = {ids[i] - last, freqs[i]};
        uint8_t lead = 0;
        const size_t lead_pos = pos++;
        for (int k = 0; k < 2; k++) {
            uint32_t v = vals[k];
            int bytes = 0;
            do {
                out[pos++] = (uint8_t)v;
                bytes++;
                v >>= 8;
            } while (v);
            lead |= (uint8_t)((bytes - 1) << (2 * k));
        }
        out[lead_pos] = lead;
        last = ids[i];
        blk_last[nb - 1] = ids[i];
        blk_n[nb - 1]++;
    }
    blk_off[nb] = pos;
    *nblocks = nb;
    return pos;
}

} // extern "C"
