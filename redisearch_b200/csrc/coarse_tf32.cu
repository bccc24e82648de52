// Batched fp32 cosine KNN, stage 1: a tcgen05 `kind::tf32` coarse pass over the HBM-resident corpus
// that keeps, per CTA and per query, the 16 best rows by APPROXIMATE distance.  Stage 2
// (rescore_kernel) recomputes those candidates with the bit-exact arithmetic of distance_core.cuh and
// stage 3 (verify) proves per query that no discarded row can belong to the exact top-k — otherwise
// the query falls back to the exact scan.  Result: the exact answer of BruteForceIndex::topKQuery
// (VS/algorithms/brute_force/brute_force.h:243-291) at tensor-core speed.
//
// Why: 256 queries x 10M x 768 fp32 is 3.93 TFLOP per corpus pass — FMA-bound on CUDA cores
// (DESIGN.md §4); TF32 tensor cores bring the pass back to the HBM/L2 roofline.  TF32 truncates each
// operand to 10 mantissa bits, so |approx - exact| <= ~2^-9 for unit vectors; that is far too coarse for
// the reference's 1e-5 parity bar, hence coarse-then-exact instead of trusting the GEMM.
//
// Kernel shape (one CTA per SM, persistent over 128-row tiles of its row range):
//   warp 0   TMA producer: A tiles [128 rows x 32 floats] (128B swizzle) into a 6-stage ring
//   warp 1   MMA issuer (one elected lane): tcgen05.mma.cta_group::1.kind::tf32, M=128 N=32 K=8,
//            A and B from shared memory, accumulator in TMEM (2 stages x 32 columns)
//   warp 2   TMEM allocator
//   warps 4-7 epilogue: tcgen05.ld 32x32b.x32 (one row x 32 queries per thread), threshold test,
//            candidate lists in shared memory with lazy compaction
// The CTA's 32 queries stay resident in shared memory (24 swizzled K-blocks, 96 KB) for the whole pass.
#include "coarse_tf32.h"
#include "distance_core.cuh"
#include "topk_common.cuh"

#include <cuda.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace rsb200 {

constexpr int kTileM = 128;       // rows per tile (UMMA_M)
constexpr int kTileN = 32;        // queries per CTA (UMMA_N)
constexpr int kBlockK = 32;       // floats per K block = 128 bytes = one swizzle row
constexpr int kUmmaK = 8;         // tf32: 32 bytes per instruction
constexpr int kStages = 6;
constexpr int kAccStages = 2;
constexpr int kCoarseThreads = 256;
constexpr int kListCap = 64;      // per-query candidate buffer in shared memory
constexpr uint32_t kStageBytes = kTileM * kBlockK * 4; // 16 KB
constexpr uint32_t kQBlockBytes = kTileN * kBlockK * 4; // 4 KB

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// slice of a tile delivered to the same shared-memory offset of every CTA in `mask`; each destination's
// mbarrier (same offset) receives the byte count
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
// pull a box into L2 ahead of time (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap *map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version [46,48) = 1,
//  layout_type [61,64) = 2 (SWIZZLE_128B)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;          // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(1024 >> 4) << 32; // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;           // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;           // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format TF32 (2) [7,10)/[10,13), K-major both,
// N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
struct CoarseSmem {
    // offsets computed at runtime from the 1024-aligned base
};

__global__ void __launch_bounds__(kCoarseThreads, 1)
coarse_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_q, uint32_t n_rows,
                   uint32_t nq, uint32_t num_kb, uint32_t tiles_total, uint32_t keep, uint32_t csize, uint32_t pf_tiles,
                   uint64_t *__restrict__ cand_out) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem;                                   // num_kb x [32 x 128B]
    uint8_t *sA = sQ + (size_t)num_kb * kQBlockBytes;     // kStages x [128 x 128B]
    uint64_t *lists = reinterpret_cast<uint64_t *>(sA + (size_t)kStages * kStageBytes); // [32][kListCap]
    uint64_t *bars = lists + kTileN * kListCap;
    uint64_t *full = bars, *empty = bars + kStages, *tfull = bars + 2 * kStages, *tempty = tfull + kAccStages;
    uint64_t *qbar = tempty + kAccStages;
    uint32_t *counts = reinterpret_cast<uint32_t *>(qbar + 1); // [32]
    uint32_t *thresh = counts + kTileN;                        // [32] orderable keys
    uint32_t *tmem_slot = thresh + kTileN;
    uint32_t *pending_flag = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q_base = blockIdx.y * kTileN;
    // this CTA's tiles: t = blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t my_tiles = (tiles_total > blockIdx.x) ? (tiles_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], csize); // one arrival per CTA of the cluster (all read the multicast tile)
        }
        for (int a = 0; a < kAccStages; a++) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 128);
        }
        mbar_init(qbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        *pending_flag = 0;
    }
    if (threadIdx.x < kTileN) {
        counts[threadIdx.x] = 0;
        thresh[threadIdx.x] = 0xFFFFFFFFu;
    }
    if (warp == 2) tmem_alloc(tmem_slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // cluster mode: the csize CTAs that share blockIdx.x (one per query group) read the SAME row tiles;
    // every CTA fetches 1/csize of each tile from HBM/L2 and multicasts it into all csize shared memories
    const uint32_t crank = (csize > 1) ? cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    const uint32_t slice_rows = kTileM / csize;
    if (csize > 1) cluster_sync_all(); // remote CTAs must see initialised barriers before signalling them

    if (warp == 0 && lane == 0) {
        // ===== TMA producer =====
        mbar_expect_tx(qbar, num_kb * kQBlockBytes);
        for (uint32_t kb = 0; kb < num_kb; kb++) tma_load_2d(sQ + (size_t)kb * kQBlockBytes, &map_q, qbar, (int)(kb * kBlockK), (int)q_base);
        uint32_t it = 0;
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t tile = blockIdx.x + i * gridDim.x;
            // HBM -> L2 prefetch kPrefetchTiles tiles ahead, so the ring below refills at L2 latency;
            // the query-group CTAs of a row range share the tiles, one of them (rotating) prefetches
            if (pf_tiles && i + pf_tiles < my_tiles && (i % gridDim.y) == blockIdx.y) {
                const uint32_t ptile = blockIdx.x + (i + pf_tiles) * gridDim.x;
                for (uint32_t kb = 0; kb < num_kb; kb++)
                    for (uint32_t r = 0; r < csize; r++)
                        tma_prefetch_2d(&map_a, (int)(kb * kBlockK), (int)(ptile * kTileM + r * slice_rows));
            }
            for (uint32_t kb = 0; kb < num_kb; kb++, it++) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_expect_tx(&full[s], kStageBytes);
                if (csize > 1)
                    tma_load_2d_mc(sA + (size_t)s * kStageBytes + (size_t)crank * slice_rows * 128, &map_a, &full[s],
                                   (int)(kb * kBlockK), (int)(tile * kTileM + crank * slice_rows), cmask);
                else
                    tma_load_2d(sA + (size_t)s * kStageBytes, &map_a, &full[s], (int)(kb * kBlockK), (int)(tile * kTileM));
            }
        }
    } else if (warp == 1 && lane == 0) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc_tf32(kTileM, kTileN);
        mbar_wait(qbar, 0);
        uint32_t it = 0;
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t a = i % kAccStages, aph = (i / kAccStages) & 1;
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * kTileN;
            for (uint32_t kb = 0; kb < num_kb; kb++, it++) {
                const uint32_t s = it % kStages, ph = (it / kStages) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_smem_desc(smem_u32(sA + (size_t)s * kStageBytes));
                const uint64_t bdesc = make_smem_desc(smem_u32(sQ + (size_t)kb * kQBlockBytes));
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; k++) {
                    // advance 32 bytes along K inside the swizzled 128-byte row: +2 in 16-byte units
                    umma_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | (uint32_t)k) != 0);
                }
                if (csize > 1)
                    umma_commit_mc(&empty[s], cmask); // this CTA is done with stage s: tell every producer of the cluster
                else
                    umma_commit(&empty[s]); // frees the A stage when these MMAs retire
            }
            umma_commit(&tfull[a]); // accumulator of this tile complete
        }
    } else if (warp >= 4) {
        // ===== epilogue: TMEM -> registers -> candidate lists =====
        const int ew = warp - 4;                 // TMEM lane quadrant = warp % 4
        const int et = threadIdx.x - 128;        // 0..127
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t tile = blockIdx.x + i * gridDim.x;
            const uint32_t a = i % kAccStages, aph = (i / kAccStages) & 1;
            mbar_wait(&tfull[a], aph);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + a * kTileN, v);
            tc_fence_before();
            mbar_arrive(&tempty[a]); // accumulator stage may be overwritten
            const uint32_t row = tile * kTileM + ew * 32 + lane;
            uint32_t pend = 0; // bit j: candidate for query j still to be stored
            if (row < n_rows) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const float d = 1.0f - __uint_as_float(v[j]);
                    v[j] = orderable_key(d);
                    if (v[j] < thresh[j]) pend |= 1u << j;
                }
            }
            // insert with retry: lists are compacted (keep best `keep`) whenever they run full
            for (;;) {
                uint32_t still = 0;
                uint32_t p = pend;
                while (p) {
                    const int j = __ffs(p) - 1;
                    p &= p - 1;
                    uint32_t key = 0;
#pragma unroll
                    for (int jj = 0; jj < 32; jj++)
                        if (jj == j) key = v[jj];
                    if (key >= thresh[j]) continue; // threshold tightened meanwhile
                    const uint32_t slot = atomicAdd(&counts[j], 1u);
                    if (slot < (uint32_t)kListCap)
                        lists[j * kListCap + slot] = ((uint64_t)key << 32) | row;
                    else
                        still |= 1u << j;
                }
                pend = still;
                if (pend) *pending_flag = 1;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                // compaction: warp ew handles queries j = ew, ew+4, ...
                const uint32_t any_pending = *pending_flag;
                for (int j = ew; j < kTileN; j += 4) {
                    const uint32_t c = min(counts[j], (uint32_t)kListCap);
                    if (c < (uint32_t)(kListCap / 2) && !any_pending) continue;
                    if (c <= keep) {
                        if (lane == 0) counts[j] = c;
                        continue;
                    }
                    // each lane holds up to 2 entries; extract the `keep` smallest by repeated warp-min
                    uint64_t e0 = (lane < (int)c) ? lists[j * kListCap + lane] : kEmptySlot;
                    uint64_t e1 = (lane + 32 < (int)c) ? lists[j * kListCap + lane + 32] : kEmptySlot;
                    __syncwarp();
                    uint64_t last = 0;
                    for (uint32_t r = 0; r < keep; r++) {
                        uint64_t m = e0 < e1 ? e0 : e1;
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) {
                            const uint64_t o = shfl_xor_u64(m, sft);
                            m = o < m ? o : m;
                        }
                        if (e0 == m) e0 = kEmptySlot; else if (e1 == m) e1 = kEmptySlot;
                        if (lane == 0) lists[j * kListCap + r] = m;
                        last = m;
                    }
                    if (lane == 0) {
                        counts[j] = keep;
                        thresh[j] = (uint32_t)(last >> 32);
                    }
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (!any_pending) break;
                if (et == 0) *pending_flag = 0;
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
        }
        // final compaction + publish: cand_out[q][blockIdx.x][keep]
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int j = ew; j < kTileN; j += 4) {
            const uint32_t c = min(counts[j], (uint32_t)kListCap);
            uint64_t e0 = (lane < (int)c) ? lists[j * kListCap + lane] : kEmptySlot;
            uint64_t e1 = (lane + 32 < (int)c) ? lists[j * kListCap + lane + 32] : kEmptySlot;
            const uint32_t q = q_base + j;
            for (uint32_t r = 0; r < keep; r++) {
                uint64_t m = e0 < e1 ? e0 : e1;
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) {
                    const uint64_t o = shfl_xor_u64(m, sft);
                    m = o < m ? o : m;
                }
                if (e0 == m) e0 = kEmptySlot; else if (e1 == m) e1 = kEmptySlot;
                if (lane == 0 && q < nq) cand_out[((size_t)q * gridDim.x + blockIdx.x) * keep + r] = m;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all(); // no CTA may exit while peers can still write its smem / barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 64);
    }
}

// ------------------------------------------------------------------------------------------------
// stage 2: exact rescoring of the candidates (bit-exact arithmetic of distance_core.cuh)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rescore_kernel(const uint8_t *rows, size_t pitch, uint32_t dim, const uint8_t *queries,
                                                      size_t qpitch, uint32_t nq, uint32_t per_query,
                                                      const uint64_t *__restrict__ cand, uint64_t *__restrict__ exact) {
    using Tile = DistTile<DT_F32, MT_IP, 1, 1>;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t total = (size_t)nq * per_query;
    for (size_t w = (size_t)blockIdx.x * 8 + warp; w < total; w += (size_t)gridDim.x * 8) {
        const uint64_t c = cand[w];
        if (c == kEmptySlot) {
            if (lane == 0) exact[w] = kEmptySlot;
            continue;
        }
        const uint32_t row = (uint32_t)c;
        const uint32_t q = (uint32_t)(w / per_query);
        const uint8_t *rowb[1] = {rows + (size_t)row * pitch};
        const uint8_t *qb[1] = {queries + (size_t)q * qpitch};
        float d[1];
        Tile::run(rowb, qb, dim, lane, d);
        if (lane == 0) exact[w] = make_composite(d[0], row);
    }
}

// stage 3: per query, is the exact top-k provably complete?  A row that is NOT among the candidates of
// its list has approx >= the list's worst kept approx a_w, hence exact >= a_w - eps.  If
// a_w - eps > e_k (the k-th best exact distance found) for every FULL list, nothing was missed.
__global__ void verify_kernel(const uint64_t *__restrict__ cand, const uint64_t *__restrict__ topk, uint32_t nq,
                              uint32_t lists_per_query, uint32_t keep, uint32_t k, float eps, uint32_t *__restrict__ ok) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint64_t kth = topk[(size_t)q * k + (k - 1)];
    bool good = true;
    if (kth == kEmptySlot) {
        // fewer than k rows found: complete only if no list was truncated
        for (uint32_t l = 0; l < lists_per_query; l++)
            if (cand[((size_t)q * lists_per_query + l) * keep + (keep - 1)] != kEmptySlot) good = false;
    } else {
        const float ek = key_to_float((uint32_t)(kth >> 32));
        for (uint32_t l = 0; l < lists_per_query; l++) {
            const uint64_t worst = cand[((size_t)q * lists_per_query + l) * keep + (keep - 1)];
            if (worst == kEmptySlot) continue; // list not full: it holds every row of its range that passed
            const float aw = key_to_float((uint32_t)(worst >> 32));
            if (!(aw - eps > ek)) good = false;
        }
    }
    ok[q] = good ? 1u : 0u;
}

// ================================================================================================
// host
// ================================================================================================
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

static bool make_map(CUtensorMap *m, const void *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner,
                     uint32_t box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

void finalize_coarse_plan(CoarsePlan &p, uint32_t nq);

bool coarse_supported(const CorpusView &c, uint32_t nq, uint32_t k) {
    if (c.dtype != DT_F32 || c.metric != MT_IP) return false; // cosine on normalised rows (and raw IP, see eps)
    if (c.dim % 4 != 0 || c.dim < 32 || c.dim > 1024) return false;
    if (c.pitch % 16 != 0) return false;
    if (k > kCoarseMaxK || nq < 16) return false;
    if (c.n_rows < 65536) return false; // tiny corpora: the exact kernel is already fast
    return encode_fn() != nullptr;
}

CoarsePlan plan_coarse(const CorpusView &c, uint32_t nq) {
    CoarsePlan p{};
    p.num_kb = (c.dim + kBlockK - 1) / kBlockK;
    p.tiles = (c.n_rows + kTileM - 1) / kTileM;
    p.grid_y = (nq + kTileN - 1) / kTileN;
    const uint32_t sms = (uint32_t)device_sm_count();
    p.grid_x = std::max(1u, std::min(p.tiles, sms / p.grid_y));
    p.keep = kCoarseKeep;
    p.csize = 1;
    p.cand_elems = (size_t)nq * p.grid_x * p.keep;
    p.smem_bytes = 1024 + (size_t)p.num_kb * kQBlockBytes + (size_t)kStages * kStageBytes + (size_t)kTileN * kListCap * 8 +
                   (2 * kStages + 2 * kAccStages + 1) * 8 + kTileN * 8 + 64;
    finalize_coarse_plan(p, nq);
    return p;
}

// Cluster size for the multicast variant: the query groups (grid.y) of one row range form a cluster.
static uint32_t pick_cluster(uint32_t grid_y) {
    static int cap = -1; // VECSIM_B200_CLUSTER = largest cluster size to use (0/1 = no clusters); default 8
    if (cap < 0) {
        const char *e = getenv("VECSIM_B200_CLUSTER");
        cap = e ? atoi(e) : 8;
        if (cap < 1) cap = 1;
    }
    for (uint32_t cs = 8; cs > 1; cs >>= 1)
        if ((int)cs <= cap && grid_y % cs == 0) return cs;
    return 1;
}

static void fill_launch_cfg(cudaLaunchConfig_t &cfg, cudaLaunchAttribute *at, const CoarsePlan &p, cudaStream_t s) {
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(p.grid_x, p.grid_y, 1);
    cfg.blockDim = dim3(kCoarseThreads);
    cfg.dynamicSmemBytes = p.smem_bytes;
    cfg.stream = s;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1;
    at[0].val.clusterDim.y = p.csize;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
}

void finalize_coarse_plan(CoarsePlan &p, uint32_t nq) {
    p.csize = pick_cluster(p.grid_y);
    if (p.csize > 1) {
        cudaFuncSetAttribute(coarse_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
        cudaLaunchConfig_t cfg;
        cudaLaunchAttribute at[1];
        CoarsePlan probe = p;
        probe.grid_x = 1;
        fill_launch_cfg(cfg, at, probe, nullptr);
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, coarse_tf32_kernel, &cfg) != cudaSuccess || nclusters < 1) {
            cudaGetLastError();
            p.csize = 1;
        } else {
            // one wave of co-resident clusters: grid_x row ranges x (grid_y / csize) clusters each
            const uint32_t per_range = p.grid_y / p.csize;
            p.grid_x = std::max(1u, std::min(p.tiles, (uint32_t)nclusters / per_range));
        }
    }
    p.cand_elems = (size_t)nq * p.grid_x * p.keep;
}

cudaError_t launch_coarse(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, const CoarsePlan &p,
                          uint64_t *d_cand, cudaStream_t s) {
    CUtensorMap ma, mq;
    if (!make_map(&ma, c.rows, c.dim, c.n_rows, c.pitch, kBlockK, kTileM / p.csize)) return cudaErrorInvalidValue;
    if (!make_map(&mq, d_queries, c.dim, nq, qpitch, kBlockK, kTileN)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(coarse_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    fill_launch_cfg(cfg, at, p, s);
    static int pf = -1; // VECSIM_B200_PREFETCH = tiles of L2 prefetch distance (default 4, 0 = off)
    if (pf < 0) {
        const char *e = getenv("VECSIM_B200_PREFETCH");
        pf = e ? atoi(e) : 4;
    }
    return cudaLaunchKernelEx(&cfg, coarse_tf32_kernel, ma, mq, c.n_rows, nq, p.num_kb, p.tiles, p.keep, p.csize, (uint32_t)pf,
                              d_cand);
}

cudaError_t launch_rescore(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, uint32_t per_query,
                           const uint64_t *d_cand, uint64_t *d_exact, cudaStream_t s) {
    const size_t total = (size_t)nq * per_query;
    const uint32_t grid = (uint32_t)std::max<size_t>(1, std::min<size_t>((total + 7) / 8, (size_t)device_sm_count() * 8));
    rescore_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(c.rows), c.pitch, c.dim, static_cast<const uint8_t *>(d_queries),
                                        qpitch, nq, per_query, d_cand, d_exact);
    return cudaGetLastError();
}

cudaError_t launch_verify(const uint64_t *d_cand, const uint64_t *d_topk, uint32_t nq, uint32_t lists_per_query, uint32_t keep,
                          uint32_t k, float eps, uint32_t *d_ok, cudaStream_t s) {
    verify_kernel<<<(nq + 127) / 128, 128, 0, s>>>(d_cand, d_topk, nq, lists_per_query, keep, k, eps, d_ok);
    return cudaGetLastError();
}

} // namespace rsb200
