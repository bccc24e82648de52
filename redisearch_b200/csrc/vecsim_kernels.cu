// sm_100a kernels of the FLAT KNN path.  See DESIGN.md §3 for the roofline of each kernel.
//
//   scan_topk_kernel     fused distance scan + per-warp top-k lists   (HBM-bound: N*rowbytes)
//   final_select_kernel  candidates -> k smallest, sorted             (tiny)
//   scan_scores_kernel   all N distances of one query -> HBM          (HBM-bound; batch iterator,
//                                                                      range query, k > 128)
//   select_scores_kernel cursor-select over a score array             (N*4 bytes per pass)
//   range_compact_kernel scores <= radius -> compacted composites
//   gather_kernel        distances of listed rows (ad-hoc / hybrid)
//   unpack / merge       reply formatting, G-way shard merge
//
// Replaces, on device: BruteForceIndex::topKQuery (VS/algorithms/brute_force/brute_force.h:243-291),
// rangeQuery (:293-326), BFS_BatchIterator::calculateScores (bfs_batch_iterator.h:24-40),
// BF_BatchIterator::getNextResults (bf_batch_iterator.h:176-200), getDistanceFrom_Unsafe
// (brute_force_single.h:200-212) and the distance functions of VS/spaces/.
#include "vecsim_kernels.h"
#include "distance_core.cuh"
#include "topk_common.cuh"

#include <algorithm>
#include <mutex>

namespace rsb200 {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

int device_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// ------------------------------------------------------------------------------------------------
// per-warp list state in shared memory
// ------------------------------------------------------------------------------------------------
struct ListState {
    uint64_t *slots; // [k]
    uint64_t *worst; // [1]
    uint32_t *wpos;  // [1]
};

// All 32 lanes; cand is warp-uniform.  Replaces the current worst entry, then rescans.
__device__ __forceinline__ void list_admit(const ListState &ls, uint32_t k, uint64_t cand, int lane) {
    if (lane == 0) ls.slots[*ls.wpos] = cand;
    __syncwarp();
    uint64_t best = 0;
    uint32_t pos = 0;
    for (uint32_t p = lane; p < k; p += 32) {
        uint64_t v = ls.slots[p];
        if (v >= best) {
            best = v;
            pos = p;
        }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        uint64_t ob = shfl_xor_u64(best, m);
        uint32_t op = __shfl_xor_sync(0xffffffffu, pos, m);
        if (ob > best || (ob == best && op < pos)) {
            best = ob;
            pos = op;
        }
    }
    if (lane == 0) {
        *ls.worst = best;
        *ls.wpos = pos;
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// fused scan + top-k
// ------------------------------------------------------------------------------------------------
struct ScanArgs {
    const uint8_t *rows;
    size_t pitch;
    uint32_t n_rows, dim;
    const uint8_t *queries; // device, 16B-aligned, qpitch (multiple of 16) apart
    size_t qpitch;
    uint32_t q_smem_pitch; // round16(query blob bytes)
    uint32_t nq, k, wq, lists_per_query;
    uint64_t *cand;
    const uint32_t *q_ok; // optional: queries already answered (coarse path verified) are skipped
    const uint32_t *abort; // optional (mapped host memory): non-zero = the caller has left (timeout), stop scanning
    uint32_t poll_mask;    // the flag is read every poll_mask + 1 tiles of a warp
};

template <int DT, int MT, int RT, int QT, bool QSMEM>
__global__ void __launch_bounds__(kScanThreads) scan_topk_kernel(const ScanArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    using Tile = DistTile<DT, MT, RT, QT>;
    using Map = typename Tile::Map;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t WQ = a.wq, WR = kScanWarps / WQ;
    const uint32_t qg = warp % WQ, rg = warp / WQ;
    const uint32_t q_cta0 = blockIdx.y * WQ * QT;
    const uint32_t q0 = q_cta0 + qg * QT;
    const uint32_t k = a.k;

    const size_t qs_bytes = QSMEM ? (size_t)WQ * QT * a.q_smem_pitch : 0;
    uint8_t *qs = smem;
    uint64_t *slots = reinterpret_cast<uint64_t *>(smem + qs_bytes);
    uint64_t *worst = slots + (size_t)kScanWarps * QT * k;
    uint32_t *wpos = reinterpret_cast<uint32_t *>(worst + kScanWarps * QT);

    if (QSMEM) {
        const uint32_t vec_per_q = a.q_smem_pitch >> 4;
        const uint32_t total = WQ * QT * vec_per_q;
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            const uint32_t qi = i / vec_per_q, vi = i - qi * vec_per_q;
            const uint32_t q = min(q_cta0 + qi, a.nq - 1);
            reinterpret_cast<uint4 *>(qs)[i] = reinterpret_cast<const uint4 *>(a.queries + (size_t)q * a.qpitch)[vi];
        }
    }
    {
        uint64_t *my = slots + (size_t)warp * QT * k;
        for (uint32_t p = lane; p < QT * k; p += 32) my[p] = kEmptySlot;
        if (lane < QT) {
            worst[warp * QT + lane] = kEmptySlot;
            wpos[warp * QT + lane] = 0;
        }
    }
    __syncthreads();

    const uint8_t *qb[QT];
#pragma unroll
    for (int j = 0; j < QT; j++) {
        if (QSMEM)
            qb[j] = qs + (size_t)(qg * QT + j) * a.q_smem_pitch;
        else
            qb[j] = a.queries + (size_t)min(q0 + j, a.nq - 1) * a.qpitch;
    }

    const uint32_t ntiles = (a.n_rows + RT - 1) / RT;
    bool active = q0 < a.nq;
    if (active && a.q_ok) { // exact fallback of the tensor-core path: only unverified queries are scanned
        bool all_ok = true;
#pragma unroll
        for (int j = 0; j < QT; j++)
            if (q0 + j < a.nq && !a.q_ok[q0 + j]) all_ok = false;
        active = !all_ok;
    }
    if (active) {
        uint32_t since_poll = 0;
        for (uint32_t t = blockIdx.x * WR + rg; t < ntiles; t += gridDim.x * WR) {
            if (a.abort && (since_poll++ & a.poll_mask) == a.poll_mask) { // a host flag over PCIe: one lane reads it now and then
                uint32_t f = 0;
                if (lane == 0) f = *reinterpret_cast<const volatile uint32_t *>(a.abort);
                if (__shfl_sync(0xffffffffu, f, 0)) break; // the partial lists are published and then ignored by the host
            }
            const uint32_t r0 = t * RT;
            const uint8_t *rowb[RT];
#pragma unroll
            for (int i = 0; i < RT; i++) rowb[i] = a.rows + (size_t)min(r0 + i, a.n_rows - 1) * a.pitch;
            float d[Map::kPerLane];
            Tile::run(rowb, qb, a.dim, lane, d);
#pragma unroll
            for (int t2 = 0; t2 < Map::kPerLane; t2++) {
                const int idx = Map::value_index(lane, t2);
                const uint32_t row = r0 + idx / QT;
                const uint32_t j = idx % QT;
                const bool valid = Map::primary(lane) && row < a.n_rows && (q0 + j) < a.nq;
                const uint64_t comp = make_composite(d[t2], row);
                const uint64_t w = worst[warp * QT + j];
                unsigned pending = __ballot_sync(0xffffffffu, valid && comp < w);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const uint64_t c = shfl_u64(comp, src);
                    const uint32_t js = __shfl_sync(0xffffffffu, j, src);
                    ListState ls{slots + ((size_t)warp * QT + js) * k, worst + warp * QT + js, wpos + warp * QT + js};
                    if (c < *ls.worst) list_admit(ls, k, c, lane);
                }
            }
        }
    }
    __syncwarp();
    // publish this warp's lists
#pragma unroll
    for (int j = 0; j < QT; j++) {
        if (q0 + j < a.nq) {
            uint64_t *dst = a.cand + ((size_t)(q0 + j) * a.lists_per_query + blockIdx.x * WR + rg) * k;
            const uint64_t *src = slots + ((size_t)warp * QT + j) * k;
            for (uint32_t p = lane; p < k; p += 32) dst[p] = src[p];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// candidates -> k smallest per query, ascending
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) final_select_kernel(const uint64_t *__restrict__ cand, uint32_t m,
                                                                    uint32_t k, uint64_t *__restrict__ out,
                                                                    const uint32_t *__restrict__ nq_dev) {
    extern __shared__ __align__(16) uint8_t smem[];
    if (nq_dev && blockIdx.x >= *nq_dev) return; // second tier: only the first *nq_dev positions hold lists
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sortn = next_pow2(kScanWarps * k);
    uint64_t *sortbuf = reinterpret_cast<uint64_t *>(smem); // [sortn]; first 8*k double as the lists
    uint64_t *worst = sortbuf + sortn;
    uint32_t *wpos = reinterpret_cast<uint32_t *>(worst + kScanWarps);
    const uint64_t *src = cand + (size_t)blockIdx.x * m;

    for (uint32_t p = threadIdx.x; p < sortn; p += blockDim.x) sortbuf[p] = kEmptySlot;
    if (threadIdx.x < kScanWarps) {
        worst[threadIdx.x] = kEmptySlot;
        wpos[threadIdx.x] = 0;
    }
    __syncthreads();
    ListState ls{sortbuf + (size_t)warp * k, worst + warp, wpos + warp};
    for (uint32_t base = warp * 32; base < m; base += kScanThreads) {
        const uint32_t i = base + lane;
        const uint64_t c = (i < m) ? src[i] : kEmptySlot;
        unsigned pending = __ballot_sync(0xffffffffu, c < *ls.worst);
        while (pending) {
            const int s = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t cc = shfl_u64(c, s);
            if (cc < *ls.worst) list_admit(ls, k, cc, lane);
        }
    }
    bitonic_sort_smem(sortbuf, sortn);
    for (uint32_t p = threadIdx.x; p < k; p += blockDim.x) out[(size_t)blockIdx.x * k + p] = sortbuf[p];
}

// ------------------------------------------------------------------------------------------------
// unfused: all scores of one query
// ------------------------------------------------------------------------------------------------
template <int DT, int MT>
__global__ void __launch_bounds__(kScanThreads) scan_scores_kernel(const uint8_t *rows, size_t pitch, uint32_t n_rows,
                                                                   uint32_t dim, const uint8_t *query,
                                                                   uint32_t q_smem_pitch, float *scores) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int RT = 4;
    using Tile = DistTile<DT, MT, RT, 1>;
    using Map = typename Tile::Map;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (q_smem_pitch >> 4); i += blockDim.x)
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(query)[i];
    __syncthreads();
    const uint8_t *qb[1] = {smem};
    const uint32_t ntiles = (n_rows + RT - 1) / RT;
    for (uint32_t t = blockIdx.x * kScanWarps + warp; t < ntiles; t += gridDim.x * kScanWarps) {
        const uint32_t r0 = t * RT;
        const uint8_t *rowb[RT];
#pragma unroll
        for (int i = 0; i < RT; i++) rowb[i] = rows + (size_t)min(r0 + i, n_rows - 1) * pitch;
        float d[Map::kPerLane];
        Tile::run(rowb, qb, dim, lane, d);
        const uint32_t row = r0 + Map::value_index(lane, 0);
        if (Map::primary(lane) && row < n_rows) scores[row] = d[0];
    }
}

// k smallest composites > cursor over a score array -> per-warp lists
__global__ void __launch_bounds__(kScanThreads) select_scores_kernel(const float *__restrict__ scores, uint32_t n,
                                                                     const uint64_t *__restrict__ cursor, uint32_t k,
                                                                     uint64_t *__restrict__ cand) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint64_t *slots = reinterpret_cast<uint64_t *>(smem);
    uint64_t *worst = slots + (size_t)kScanWarps * k;
    uint32_t *wpos = reinterpret_cast<uint32_t *>(worst + kScanWarps);
    for (uint32_t p = threadIdx.x; p < kScanWarps * k; p += blockDim.x) slots[p] = kEmptySlot;
    if (threadIdx.x < kScanWarps) {
        worst[threadIdx.x] = kEmptySlot;
        wpos[threadIdx.x] = 0;
    }
    __syncthreads();
    const bool has_cursor = cursor != nullptr;
    const uint64_t lo = has_cursor ? *cursor : 0;
    ListState ls{slots + (size_t)warp * k, worst + warp, wpos + warp};
    const uint32_t gw = blockIdx.x * kScanWarps + warp, nw = gridDim.x * kScanWarps;
    for (uint64_t base = (uint64_t)gw * 32; base < n; base += (uint64_t)nw * 32) {
        const uint32_t i = (uint32_t)base + lane;
        bool valid = i < n;
        uint64_t c = kEmptySlot;
        if (valid) {
            c = make_composite(scores[i], i);
            valid = !has_cursor || c > lo;
        }
        unsigned pending = __ballot_sync(0xffffffffu, valid && c < *ls.worst);
        while (pending) {
            const int s = __ffs(pending) - 1;
            pending &= pending - 1;
            const uint64_t cc = shfl_u64(c, s);
            if (cc < *ls.worst) list_admit(ls, k, cc, lane);
        }
    }
    __syncwarp();
    uint64_t *dst = cand + (size_t)gw * k;
    for (uint32_t p = lane; p < k; p += 32) dst[p] = ls.slots[p];
}

__global__ void range_compact_kernel(const float *__restrict__ scores, uint32_t n, float radius,
                                     uint64_t *__restrict__ out, uint32_t *__restrict__ count) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float s = scores[i];
        if (s <= radius) { // brute_force.h:315 (NaN never passes)
            const uint32_t pos = atomicAdd(count, 1u);
            out[pos] = make_composite(s, i);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// ad-hoc gather: one warp per listed row
// ------------------------------------------------------------------------------------------------
template <int DT, int MT>
__global__ void __launch_bounds__(kScanThreads) gather_kernel(const uint8_t *rows, size_t pitch, uint32_t dim,
                                                              const uint8_t *query, uint32_t q_smem_pitch,
                                                              const uint32_t *ids, uint32_t count, float *out) {
    extern __shared__ __align__(16) uint8_t smem[];
    using Tile = DistTile<DT, MT, 1, 1>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x; i < (q_smem_pitch >> 4); i += blockDim.x)
        reinterpret_cast<uint4 *>(smem)[i] = reinterpret_cast<const uint4 *>(query)[i];
    __syncthreads();
    const uint8_t *qb[1] = {smem};
    for (uint32_t w = blockIdx.x * kScanWarps + warp; w < count; w += gridDim.x * kScanWarps) {
        const uint32_t id = ids[w];
        if (id == 0xFFFFFFFFu) {
            if (lane == 0) out[w] = __uint_as_float(0x7FC00000u);
            continue;
        }
        const uint8_t *rowb[1] = {rows + (size_t)id * pitch};
        float d[1];
        Tile::run(rowb, qb, dim, lane, d);
        if (lane == 0) out[w] = d[0];
    }
}

// ------------------------------------------------------------------------------------------------
// reply formatting and shard merge
// ------------------------------------------------------------------------------------------------
__global__ void unpack_results_kernel(const uint64_t *__restrict__ comp, uint32_t total,
                                      const uint64_t *__restrict__ id_to_label, int64_t *__restrict__ labels,
                                      float *__restrict__ scores) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const uint64_t c = comp[i];
    if (c == kEmptySlot) {
        labels[i] = -1;
        scores[i] = __uint_as_float(0x7FC00000u);
    } else {
        const uint32_t id = (uint32_t)c;
        labels[i] = id_to_label ? (int64_t)id_to_label[id] : (int64_t)id;
        scores[i] = key_to_float((uint32_t)(c >> 32));
    }
}

// One CTA per query, rank sort of G*k (score,label) pairs; empty entries have label < 0.
// Shard g's arrays start score_stride floats / label_stride int64s after shard g-1's ([G][nq][k] arrays: nq*k; the
// packed exchange buffer of the shard group: one block of labels + scores per shard).
__global__ void merge_shards_kernel(const float *__restrict__ scores, const int64_t *__restrict__ labels, uint32_t G,
                                    uint32_t nq, uint32_t k, size_t score_stride, size_t label_stride,
                                    float *__restrict__ out_scores, int64_t *__restrict__ out_labels) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint32_t q = blockIdx.x, n = G * k;
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem);
    int64_t *labs = reinterpret_cast<int64_t *>(smem + (((size_t)n * 4 + 15) & ~(size_t)15));
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t g = i / k, p = i - g * k;
        const size_t src = (size_t)q * k + p;
        const int64_t l = labels[(size_t)g * label_stride + src];
        labs[i] = l;
        keys[i] = (l < 0) ? 0xFFFFFFFFu : orderable_key(scores[(size_t)g * score_stride + src]);
    }
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) {
        out_labels[(size_t)q * k + i] = -1;
        out_scores[(size_t)q * k + i] = __uint_as_float(0x7FC00000u);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint32_t ki = keys[i];
        const int64_t li = labs[i];
        if (li < 0) continue;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; j++) {
            const uint32_t kj = keys[j];
            const int64_t lj = labs[j];
            if (lj < 0) continue;
            // (score, label) lexicographic — VS/utils/query_result_utils.h:19-23; index breaks
            // exact duplicates (the same label cannot live on two shards).
            if (kj < ki || (kj == ki && (lj < li || (lj == li && j < i)))) rank++;
        }
        if (rank < k) {
            out_labels[(size_t)q * k + rank] = li;
            out_scores[(size_t)q * k + rank] = key_to_float(ki);
        }
    }
}

// ================================================================================================
// host side: dispatch
// ================================================================================================
static inline uint32_t round16(uint32_t v) { return (v + 15u) & ~15u; }

static uint32_t query_blob_bytes(const CorpusView &c) {
    switch (c.dtype) {
    case DT_F32: return c.dim * 4;
    case DT_F16:
    case DT_BF16: return c.dim * 2;
    default: return c.dim + (c.metric == MT_COS ? 4 : 0);
    }
}

template <typename K>
static cudaError_t ensure_smem(K kernel, size_t bytes) {
    if (bytes <= 48 * 1024) return cudaSuccess;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <typename K>
static int occupancy(K kernel, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kScanThreads, smem) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

constexpr size_t kMaxQuerySmem = 96 * 1024; // beyond this the batched scan reads queries through L1

ScanPlan plan_scan_topk(const CorpusView &c, uint32_t nq, uint32_t k) {
    ScanPlan p{};
    p.qt = (nq == 1) ? 1 : 8;
    uint32_t groups = (nq + p.qt - 1) / p.qt;
    p.wq = groups >= 8 ? 8 : groups >= 4 ? 4 : groups >= 2 ? 2 : 1;
    p.grid_y = (groups + p.wq - 1) / p.wq;
    const uint32_t wr = kScanWarps / p.wq;
    const uint32_t qsp = round16(query_blob_bytes(c));
    size_t qs = (size_t)p.wq * p.qt * qsp;
    if (qs > kMaxQuerySmem) qs = 0;
    p.smem_bytes = qs + (size_t)kScanWarps * p.qt * k * 8 + (size_t)kScanWarps * p.qt * 12;
    // persistent grid: resident CTAs only, split evenly over the query slices
    const int sms = device_sm_count();
    const uint32_t rt = 4;
    const uint32_t ntiles = (c.n_rows + rt - 1) / rt;
    uint32_t want = (ntiles + wr - 1) / wr;
    uint32_t resident = (uint32_t)sms * 2u; // refined at launch time from the occupancy API
    p.grid_x = std::max(1u, std::min(want, std::max(1u, resident / p.grid_y)));
    p.lists_per_query = p.grid_x * wr;
    p.cand_elems = (size_t)nq * p.lists_per_query * k;
    return p;
}

template <int DT, int MT, int RT, int QT, bool QSMEM>
static cudaError_t launch_scan_inst(const ScanArgs &a, const ScanPlan &plan, cudaStream_t s) {
    auto kern = scan_topk_kernel<DT, MT, RT, QT, QSMEM>;
    cudaError_t e = ensure_smem(kern, plan.smem_bytes);
    if (e != cudaSuccess) return e;
    kern<<<dim3(plan.grid_x, plan.grid_y), kScanThreads, plan.smem_bytes, s>>>(a);
    return cudaGetLastError();
}

template <int DT, int MT>
static cudaError_t launch_scan_dm(const ScanArgs &a, const ScanPlan &plan, bool qsmem, cudaStream_t s) {
    if (plan.qt == 1) return launch_scan_inst<DT, MT, 4, 1, true>(a, plan, s);
    if (qsmem) return launch_scan_inst<DT, MT, 4, 8, true>(a, plan, s);
    return launch_scan_inst<DT, MT, 4, 8, false>(a, plan, s);
}

#define RSB_DISPATCH_DM(dtype, metric, CALL)                                                         \
    switch (dtype) {                                                                                 \
    case DT_F32:                                                                                     \
        if ((metric) == MT_L2) { CALL(DT_F32, MT_L2); } else { CALL(DT_F32, MT_IP); }                \
        break;                                                                                       \
    case DT_F16:                                                                                     \
        if ((metric) == MT_L2) { CALL(DT_F16, MT_L2); } else { CALL(DT_F16, MT_IP); }                \
        break;                                                                                       \
    case DT_BF16:                                                                                    \
        if ((metric) == MT_L2) { CALL(DT_BF16, MT_L2); } else { CALL(DT_BF16, MT_IP); }              \
        break;                                                                                       \
    case DT_I8:                                                                                      \
        if ((metric) == MT_L2) { CALL(DT_I8, MT_L2); }                                               \
        else if ((metric) == MT_IP) { CALL(DT_I8, MT_IP); }                                          \
        else { CALL(DT_I8, MT_COS); }                                                                \
        break;                                                                                       \
    case DT_U8:                                                                                      \
        if ((metric) == MT_L2) { CALL(DT_U8, MT_L2); }                                               \
        else if ((metric) == MT_IP) { CALL(DT_U8, MT_IP); }                                          \
        else { CALL(DT_U8, MT_COS); }                                                                \
        break;                                                                                       \
    }

__global__ void blend_kernel(const uint32_t *__restrict__ ok, const uint64_t *__restrict__ a, const uint64_t *__restrict__ b,
                             uint32_t total, uint32_t k, uint64_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < total) out[i] = ok[i / k] ? a[i] : b[i];
}
cudaError_t launch_blend(const uint32_t *d_ok, const uint64_t *d_a, const uint64_t *d_b, uint32_t nq, uint32_t k, uint64_t *d_out,
                         cudaStream_t s, LaunchCounters *ctr) {
    const uint32_t total = nq * k;
    if (!total) return cudaSuccess;
    blend_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_ok, d_a, d_b, total, k, d_out);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

cudaError_t launch_scan_topk(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, uint32_t k,
                             const ScanPlan &plan, uint64_t *d_cand, cudaStream_t s, LaunchCounters *ctr,
                             const uint32_t *d_q_ok, const uint32_t *d_abort) {
    if (nq == 0 || k == 0 || k > (uint32_t)kMaxFusedK || c.n_rows == 0) return cudaErrorInvalidValue;
    ScanArgs a{};
    a.rows = static_cast<const uint8_t *>(c.rows);
    a.pitch = c.pitch;
    a.n_rows = c.n_rows;
    a.dim = c.dim;
    a.queries = static_cast<const uint8_t *>(d_queries);
    a.qpitch = qpitch;
    a.q_smem_pitch = round16(query_blob_bytes(c));
    a.nq = nq;
    a.k = k;
    a.wq = plan.wq;
    a.lists_per_query = plan.lists_per_query;
    a.cand = d_cand;
    a.q_ok = d_q_ok;
    a.abort = d_abort;
    a.poll_mask = nq >= 16 ? 15u : 255u; // a batched tile costs ~100x a single-query tile: keep the reaction time in milliseconds
    const bool qsmem = (size_t)plan.wq * plan.qt * a.q_smem_pitch <= kMaxQuerySmem;
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_SCAN(DT, MT) e = launch_scan_dm<DT, MT>(a, plan, qsmem, s)
    RSB_DISPATCH_DM(c.dtype, c.metric, CALL_SCAN)
#undef CALL_SCAN
    if (ctr) ctr->launches++;
    return e;
}

cudaError_t launch_final_select(const uint64_t *d_cand, uint32_t nq, uint32_t m_per_query, uint32_t k,
                                uint64_t *d_out, cudaStream_t s, LaunchCounters *ctr, const uint32_t *d_nq_dev) {
    if (k == 0 || k > (uint32_t)kMaxFusedK) return cudaErrorInvalidValue;
    const size_t smem = (size_t)next_pow2(kScanWarps * k) * 8 + kScanWarps * 12;
    final_select_kernel<<<nq, kScanThreads, smem, s>>>(d_cand, m_per_query, k, d_out, d_nq_dev);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

template <int DT, int MT>
static cudaError_t launch_scores_inst(const CorpusView &c, const void *d_query, float *d_scores, cudaStream_t s) {
    auto kern = scan_scores_kernel<DT, MT>;
    const uint32_t qsp = round16(query_blob_bytes(c));
    cudaError_t e = ensure_smem(kern, qsp);
    if (e != cudaSuccess) return e;
    const uint32_t ntiles = (c.n_rows + 3) / 4;
    const uint32_t want = (ntiles + kScanWarps - 1) / kScanWarps;
    const uint32_t grid = std::max(1u, std::min(want, (uint32_t)(device_sm_count() * occupancy(kern, qsp))));
    kern<<<grid, kScanThreads, qsp, s>>>(static_cast<const uint8_t *>(c.rows), c.pitch, c.n_rows, c.dim,
                                         static_cast<const uint8_t *>(d_query), qsp, d_scores);
    return cudaGetLastError();
}

cudaError_t launch_scan_scores(const CorpusView &c, const void *d_query, float *d_scores, cudaStream_t s,
                               LaunchCounters *ctr) {
    if (c.n_rows == 0) return cudaSuccess;
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_SCORES(DT, MT) e = launch_scores_inst<DT, MT>(c, d_query, d_scores, s)
    RSB_DISPATCH_DM(c.dtype, c.metric, CALL_SCORES)
#undef CALL_SCORES
    if (ctr) ctr->launches++;
    return e;
}

static uint32_t select_grid(uint32_t n) {
    const uint32_t want = (n + kScanThreads - 1) / kScanThreads;
    return std::max(1u, std::min(want, (uint32_t)device_sm_count() * 2u));
}
uint32_t plan_select_scores_lists(uint32_t n) { return select_grid(n) * kScanWarps; }

cudaError_t launch_select_scores(const float *d_scores, uint32_t n, const uint64_t *d_cursor, uint32_t k,
                                 uint64_t *d_cand, cudaStream_t s, LaunchCounters *ctr) {
    if (k == 0 || k > (uint32_t)kMaxFusedK) return cudaErrorInvalidValue;
    const size_t smem = (size_t)kScanWarps * k * 8 + kScanWarps * 12;
    select_scores_kernel<<<select_grid(n), kScanThreads, smem, s>>>(d_scores, n, d_cursor, k, d_cand);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

cudaError_t launch_range_compact(const float *d_scores, uint32_t n, float radius, uint64_t *d_out, uint32_t *d_count,
                                 cudaStream_t s, LaunchCounters *ctr) {
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    const uint32_t grid = std::max(1u, std::min((n + 255u) / 256u, (uint32_t)device_sm_count() * 8u));
    range_compact_kernel<<<grid, 256, 0, s>>>(d_scores, n, radius, d_out, d_count);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

template <int DT, int MT>
static cudaError_t launch_gather_inst(const CorpusView &c, const void *d_query, const uint32_t *d_ids, uint32_t count,
                                      float *d_out, cudaStream_t s) {
    auto kern = gather_kernel<DT, MT>;
    const uint32_t qsp = round16(query_blob_bytes(c));
    cudaError_t e = ensure_smem(kern, qsp);
    if (e != cudaSuccess) return e;
    const uint32_t want = (count + kScanWarps - 1) / kScanWarps;
    const uint32_t grid = std::max(1u, std::min(want, (uint32_t)(device_sm_count() * occupancy(kern, qsp))));
    kern<<<grid, kScanThreads, qsp, s>>>(static_cast<const uint8_t *>(c.rows), c.pitch, c.dim,
                                         static_cast<const uint8_t *>(d_query), qsp, d_ids, count, d_out);
    return cudaGetLastError();
}

// labels (docIds) -> internal row ids through a dense device table; absent / out of range -> 0xFFFFFFFF (-> NaN distance)
__global__ void __launch_bounds__(256) map_labels_kernel(const uint32_t *__restrict__ labels, uint32_t n, const uint32_t *__restrict__ table,
                                                         uint32_t table_size, uint32_t *__restrict__ ids) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t l = labels[i];
        ids[i] = l < table_size ? table[l] : 0xFFFFFFFFu;
    }
}
cudaError_t launch_map_labels(const uint32_t *d_labels, uint32_t n, const uint32_t *d_table, uint32_t table_size, uint32_t *d_ids,
                              cudaStream_t s, LaunchCounters *ctr) {
    if (n == 0) return cudaSuccess;
    const uint32_t grid = std::min<uint32_t>((n + 255) / 256, (uint32_t)device_sm_count() * 8);
    map_labels_kernel<<<grid, 256, 0, s>>>(d_labels, n, d_table, table_size, d_ids);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}
// out_labels[i] = labels[index part of comp[i]] (0 for empty slots)
__global__ void pick_labels_kernel(const uint64_t *__restrict__ comp, uint32_t k, const uint32_t *__restrict__ labels, uint32_t *__restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) out[i] = comp[i] == kEmptySlot ? 0u : labels[(uint32_t)comp[i]];
}
cudaError_t launch_pick_labels(const uint64_t *d_comp, uint32_t k, const uint32_t *d_labels, uint32_t *d_out, cudaStream_t s,
                               LaunchCounters *ctr) {
    if (k == 0) return cudaSuccess;
    pick_labels_kernel<<<(k + 127) / 128, 128, 0, s>>>(d_comp, k, d_labels, d_out);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

cudaError_t launch_gather_distances(const CorpusView &c, const void *d_query, const uint32_t *d_ids, uint32_t count,
                                    float *d_out, cudaStream_t s, LaunchCounters *ctr) {
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaErrorInvalidValue;
#define CALL_GATHER(DT, MT) e = launch_gather_inst<DT, MT>(c, d_query, d_ids, count, d_out, s)
    RSB_DISPATCH_DM(c.dtype, c.metric, CALL_GATHER)
#undef CALL_GATHER
    if (ctr) ctr->launches++;
    return e;
}

cudaError_t launch_unpack_results(const uint64_t *d_comp, uint32_t nq, uint32_t k, const uint64_t *d_id_to_label,
                                  int64_t *d_labels, float *d_scores, cudaStream_t s, LaunchCounters *ctr) {
    const uint32_t total = nq * k;
    if (total == 0) return cudaSuccess;
    unpack_results_kernel<<<(total + 255) / 256, 256, 0, s>>>(d_comp, total, d_id_to_label, d_labels, d_scores);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

cudaError_t launch_merge_shards(const float *d_scores, const int64_t *d_labels, uint32_t G, uint32_t nq, uint32_t k,
                                float *d_out_scores, int64_t *d_out_labels, cudaStream_t s, LaunchCounters *ctr,
                                size_t score_stride, size_t label_stride) {
    if (score_stride == 0) score_stride = (size_t)nq * k;
    if (label_stride == 0) label_stride = (size_t)nq * k;
    if (nq == 0 || k == 0 || G == 0) return cudaSuccess;
    const size_t n = (size_t)G * k;
    const size_t smem = ((n * 4 + 15) & ~(size_t)15) + n * 8;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = ensure_smem(merge_shards_kernel, smem);
    if (e != cudaSuccess) return e;
    merge_shards_kernel<<<nq, 256, smem, s>>>(d_scores, d_labels, G, nq, k, score_stride, label_stride, d_out_scores, d_out_labels);
    if (ctr) ctr->launches++;
    return cudaGetLastError();
}

} // namespace rsb200
