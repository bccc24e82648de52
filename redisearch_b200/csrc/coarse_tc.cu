// Batched KNN on the tensor cores (tcgen05, sm_100a).  Replaces, for query batches, the B independent passes of
// BruteForceIndex::topKQuery (VS/algorithms/brute_force/brute_force.h:243-291) + the distance kernels of VS/spaces/.
//
// fp32 cosine corpora — coarse-then-exact: 256 queries x 10M x 768 is 3.93 TFLOP per corpus pass,
// FMA-bound on CUDA cores, and no tensor-core operand type reproduces the reference's fp32 bits.  So
//   stage 1  a tcgen05 GEMM with APPROXIMATE distances keeps, per CTA row range and query, the kCoarseKeep best rows
//            (coarse_qtmem_kernel<false,...> over an fp16 shadow copy, or coarse_kernel<CfgTF32> over the fp32 rows);
//   stage 2  rescore_kernel recomputes those candidates from the fp32 rows with the bit-exact arithmetic of
//            distance_core.cuh;
//   stage 3  verify_kernel proves per query that no discarded row can belong to the exact top-k — otherwise the query
//            falls back to the exact scan, on the device.
// fp16 / bf16 corpora (coarse_qtmem_kernel<true,*,0>) and int8 / uint8 corpora (<true,*,1|2>, kind::i8): the
// tensor-core result IS the distance (16-bit: fp32-accumulated exact products, bar 1e-2; 8-bit: exact integer dot
// products + the reference's float expression, bit-exact), each CTA keeps its exact top-k.
//
// coarse_qtmem_kernel (one CTA per SM, persistent over the 128-row tiles of its row range):
//   warp 0    producer: row tiles into an n-stage shared-memory ring (contiguous bulk copies from the tiled shadow, or
//             128B-swizzle tensor-map boxes from a row-major 16/8-bit corpus), multicast across the query-group cluster
//   warp 1    MMA issuer (elect.sync lane): tcgen05.mma.cta_group::1, M = 128 queries, N = 128 rows, 32 bytes of K per
//             instruction; the queries are the A operand from TENSOR MEMORY (first 8 K blocks) / shared memory (rest)
//   warp 2    TMEM allocator
//   warps 4-7 epilogue: a thread owns one query; tcgen05.ld drains an accumulator stage to registers, raw dot products
//             are tested against a register threshold, survivors go to a per-query list in global memory that a
//             warp-wide radix select cuts back to the best `keep`
// coarse_kernel<CfgTF32>: the earlier SS-mode shape (32 queries per CTA in shared memory, rows = A operand), kept for
// fp32 corpora without shadow memory (mode 2) and rows wider than 896.
#include "coarse_tc.h"
#include "distance_core.cuh"
#include "topk_common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace rsb200 {

constexpr int kTileM = 128;       // rows per tile (UMMA_M)
constexpr int kMaxStages = 8;     // ring depth is chosen at plan time from the shared memory left over
constexpr int kAccStages = 2;
constexpr int kCoarseThreads = 256;
constexpr int kListCap = 64;      // per-query candidate buffer in shared memory
constexpr uint32_t kStageBytes = kTileM * 128; // one K block of a row tile: 128 rows x 128 bytes = 16 KB

struct CfgTF32 {
    static constexpr int kTileN = 32;   // queries per CTA (UMMA_N)
    static constexpr int kBlockK = 32;  // elements per K block = 128 bytes = one swizzle row
    static constexpr int kElem = 4;
    static constexpr uint32_t kFmt = 2; // UMMA a/b format TF32
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one lane of a converged warp; keeps the surrounding code warp-uniform so that descriptors and barrier
// addresses stay in uniform registers (a `lane == 0` branch makes the compiler re-broadcast them per MMA,
// which left the issuing thread — not the tensor pipe — as the bottleneck: ~57 clk per instruction)
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning — the pass runs at the 1000 W power cap, so idle polling costs clock
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
    } while (!done);
}
// pure spin (no suspend hint): for the two single-warp roles whose wake-up latency is on the critical path (the
// producer and the MMA issuer) — a sleeping try_wait is resumed through NANOSLEEP.SYNCS and adds to every stage hand-over
__device__ __forceinline__ void mbar_wait_spin(uint64_t *bar, uint32_t parity) {
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
// slice of a tile delivered to the same shared-memory offset of every CTA in `mask`
__device__ __forceinline__ void tma_load_2d_mc(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// contiguous global -> shared copy through the TMA engine (no tensor map), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// the same, delivered to the same shared-memory offset of every CTA in `mask`; each destination's mbarrier (same
// offset) receives the byte count
__device__ __forceinline__ void bulk_load_1d_mc(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// ---- CTA pair (cta_group::2) plumbing --------------------------------------------------------------------------------
// address of `local_smem_addr` in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// arrive on an mbarrier of another CTA of the cluster (release at cluster scope: what this thread has observed is published)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with acquire at cluster scope (the arrivals come from the peer CTA)
__device__ __forceinline__ void mbar_wait_spin_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 across the pair: A = each CTA's own 128 queries (tensor memory / shared memory),
// B = 128 rows, 64 in each CTA's shared memory at the same offset.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_ts_f16_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ss_f16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the barrier at this offset in EVERY CTA of `mask` when all previously issued pair MMAs have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
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
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, fp32 accumulate; 32 bytes of K per instruction
template <class Cfg>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (Cfg::kElem == 4) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(d_tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// D[tmem] (+)= A[tmem] * B[smem]^T: A (the resident queries) is read from tensor memory, only B streams
// through shared memory
__device__ __forceinline__ void umma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ss_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 8-bit integer operands (kind::i8: 32 elements per instruction, int32 accumulate) — exact dot products
__device__ __forceinline__ void umma_ts_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ss_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
template <int kOp>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kOp == 0 || kOp == 3)
        umma_ts_f16(d_tmem, a_tmem, bdesc, idesc, accumulate);
    else
        umma_ts_i8(d_tmem, a_tmem, bdesc, idesc, accumulate);
}
template <int kOp>
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if constexpr (kOp == 0 || kOp == 3)
        umma_ss_f16(d_tmem, adesc, bdesc, idesc, accumulate);
    else
        umma_ss_i8(d_tmem, adesc, bdesc, idesc, accumulate);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
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
// cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6), a/b format [7,10)/[10,13) (F16 0, TF32 2), K-major both,
// N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ constexpr uint32_t make_idesc(uint32_t fmt, int m, int n) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
template <class Cfg>
__global__ void __launch_bounds__(kCoarseThreads, 1)
coarse_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_q, uint32_t n_rows, uint32_t nq,
              uint32_t num_kb, uint32_t tiles_total, uint32_t keep, uint32_t nstages, uint64_t *__restrict__ cand_out) {
    constexpr int kTileN = Cfg::kTileN;
    constexpr int kHalves = kTileN / 32;
    constexpr uint32_t kQBlockBytes = kTileN * 128;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the 128B swizzle atoms
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem;                                   // num_kb x [kTileN x 128B]
    uint8_t *sA = sQ + (size_t)num_kb * kQBlockBytes;     // nstages x [128 x 128B]
    uint64_t *lists = reinterpret_cast<uint64_t *>(sA + (size_t)nstages * kStageBytes); // [kTileN][kListCap]
    uint64_t *bars = lists + kTileN * kListCap;
    uint64_t *full = bars, *empty = bars + kMaxStages, *tfull = bars + 2 * kMaxStages, *tempty = tfull + kAccStages;
    uint64_t *qbar = tempty + kAccStages;
    uint32_t *counts = reinterpret_cast<uint32_t *>(qbar + 1); // [kTileN]
    uint32_t *thresh = counts + kTileN;                        // [kTileN] orderable keys
    uint32_t *tmem_slot = thresh + kTileN;
    uint32_t *pending_flag = tmem_slot + 1;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q_base = blockIdx.y * kTileN;
    // this CTA's tiles: t = blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint32_t my_tiles = (tiles_total > blockIdx.x) ? (tiles_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nstages; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
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
    if (warp == 2) tmem_alloc(tmem_slot, kAccStages * kTileN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            mbar_expect_tx(qbar, num_kb * kQBlockBytes);
            for (uint32_t kb = 0; kb < num_kb; kb++)
                tma_load_2d(sQ + (size_t)kb * kQBlockBytes, &map_q, qbar, (int)(kb * Cfg::kBlockK), (int)q_base);
        }
        __syncwarp();
        uint32_t s = 0, ph = 0;
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t tile = blockIdx.x + i * gridDim.x;
            for (uint32_t kb = 0; kb < num_kb; kb++) {
                mbar_wait(&empty[s], ph ^ 1);
                if (elect_one_sync()) {
                    mbar_expect_tx(&full[s], kStageBytes);
                    tma_load_2d(sA + (size_t)s * kStageBytes, &map_a, &full[s], (int)(kb * Cfg::kBlockK), (int)(tile * kTileM));
                }
                __syncwarp();
                if (++s == nstages) s = 0, ph ^= 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = make_idesc(Cfg::kFmt, kTileM, kTileN);
        mbar_wait(qbar, 0);
        uint32_t s = 0, ph = 0;
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t a = i % kAccStages, aph = (i / kAccStages) & 1;
            mbar_wait(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * kTileN;
            for (uint32_t kb = 0; kb < num_kb; kb++) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint64_t adesc = make_smem_desc(smem_u32(sA + (size_t)s * kStageBytes));
                const uint64_t bdesc = make_smem_desc(smem_u32(sQ + (size_t)kb * kQBlockBytes));
                if (elect_one_sync()) {
                    // advance 32 bytes along K inside the swizzled 128-byte row: +2 in 16-byte units
                    umma_ss<Cfg>(d_tmem, adesc, bdesc, idesc, kb != 0);
                    umma_ss<Cfg>(d_tmem, adesc + 2, bdesc + 2, idesc, 1);
                    umma_ss<Cfg>(d_tmem, adesc + 4, bdesc + 4, idesc, 1);
                    umma_ss<Cfg>(d_tmem, adesc + 6, bdesc + 6, idesc, 1);
                    umma_commit(&empty[s]); // frees the A stage when these MMAs retire
                }
                __syncwarp();
                if (++s == nstages) s = 0, ph ^= 1;
            }
            if (elect_one_sync()) umma_commit(&tfull[a]); // accumulator of this tile complete
            __syncwarp();
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
            uint32_t v[kHalves][32];
#pragma unroll
            for (int h = 0; h < kHalves; h++) tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + a * kTileN + h * 32, v[h]);
            tc_fence_before();
            mbar_arrive(&tempty[a]); // accumulator stage may be overwritten
            const uint32_t row = tile * kTileM + ew * 32 + lane;
            uint32_t pend[kHalves]; // bit j: candidate for query h*32+j still to be stored
#pragma unroll
            for (int h = 0; h < kHalves; h++) {
                pend[h] = 0;
                if (row < n_rows) {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const float d = 1.0f - __uint_as_float(v[h][j]);
                        v[h][j] = orderable_key(d);
                        if (v[h][j] < thresh[h * 32 + j]) pend[h] |= 1u << j;
                    }
                }
            }
            // insert with retry: lists are compacted (keep best `keep`) whenever they run full
            for (;;) {
                uint32_t any_left = 0;
#pragma unroll
                for (int h = 0; h < kHalves; h++) {
                    uint32_t still = 0;
                    uint32_t p = pend[h];
                    while (p) {
                        const int jj = __ffs(p) - 1;
                        p &= p - 1;
                        uint32_t key = 0;
#pragma unroll
                        for (int x = 0; x < 32; x++)
                            if (x == jj) key = v[h][x];
                        const int j = h * 32 + jj;
                        if (key >= thresh[j]) continue; // threshold tightened meanwhile
                        const uint32_t slot = atomicAdd(&counts[j], 1u);
                        if (slot < (uint32_t)kListCap)
                            lists[j * kListCap + slot] = ((uint64_t)key << 32) | row;
                        else
                            still |= 1u << jj;
                    }
                    pend[h] = still;
                    any_left |= still;
                }
                if (any_left) *pending_flag = 1;
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
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kAccStages * kTileN);
    }
}

// ------------------------------------------------------------------------------------------------
// fp16 variant with the QUERIES resident in tensor memory (roles swapped: M = 128 queries, N = 64 rows).
// Measured on the SS kernel above: with both operands in shared memory every MMA re-reads its query
// block, and the pass is bound by shared-memory operand bandwidth (skipping the MMAs halves the time
// while the tensor pipe is 12% busy).  Here the 128 queries of a CTA are written once into TMEM
// (lane = query, 2 fp16 per 32-bit column, 384 columns at dim 768) and the MMA reads them from there;
// shared memory only carries the streaming row tiles (2 KB per MMA instead of 6), a CTA serves 128
// queries instead of 64 (half the L2->SM ingest), and in the epilogue a thread owns one QUERY: its
// threshold lives in a register and its candidate list needs no atomics.
//   TMEM columns: [0,256) two accumulator stages of a 128-row tile; [256,512) the first 8 K blocks (512 dims) of
//   the queries.  K blocks beyond that live in shared memory (swizzled like the row tiles) and are multiplied in
//   SS mode — 107 instead of 76 clk per MMA, but with a single accumulator the tensor pipe idled a third of the
//   time while the epilogue drained it (ncu: 33% of the issuer's samples on the tempty barrier).
//   The candidate lists live in global memory (L2): appends are rare once the thresholds have settled, and shared
//   memory goes to the row-tile ring.
// ------------------------------------------------------------------------------------------------
constexpr int kQM = 128;            // queries per CTA
constexpr int kQN = 128;            // rows per tile = N of one MMA (measured: N=64 58 clk, N=128 76 clk per instruction)
constexpr int kQMaxStages = 16; // (a 32 KB stage fits 7 times; the CTA-pair build packs 16 KB half-stages)
constexpr int kQKbPerStage = 2;     // K blocks per pipeline stage: amortises the barrier round trip over 8 MMAs
// candidate lists: kEpl entries per lane of the compacting warp -> kEpl*32 slots per query; a list is cut back to
// `keep` when it holds more than (slots - 32) entries after a 32-row chunk.  kEpl = 3 (96 slots) serves keep <= 32,
// kEpl = 8 (256 slots) keep <= 128.
constexpr int kQListStride = 129;   // lists[slot * stride + query]: conflict-free appends
constexpr uint32_t kQBlockBytes = kQN * 128;                    // one K block of a row tile: 128 rows x 128 bytes
constexpr uint32_t kQStageBytes = kQKbPerStage * kQBlockBytes; // 32 KB
constexpr uint32_t kQAccCols = 2 * kQN; // two accumulator stages
constexpr uint32_t kQTmemKb = (512 - kQAccCols) / 32; // K blocks of the queries that fit in tensor memory (8 = 512 dims)

// Cut list `q` (c entries, keep < c <= 32 * kEpl) back to its `keep` smallest, unordered, in slots [0, keep); returns the
// key of the worst kept entry (the new admission threshold).  Warp-wide radix select on the 32-bit key — 32 rounds
// of ballots — instead of `keep` rounds of warp-min extraction (measured: 21K clk per call, a third of the
// epilogue's time and, worse, a stall of the accumulator hand-back).  Among entries tied with the threshold key the
// lowest row ids are kept, so the kept set is exactly the `keep` smallest composites (key, row).
template <int kEpl>
__device__ __forceinline__ uint32_t select_keep(uint64_t *lists, int q, uint32_t c, uint32_t keep, int lane) {
    __syncwarp(); // the list was appended to by one lane: order its (global-memory) writes before the warp's reads
    uint64_t e[kEpl];
    uint32_t k[kEpl];
    bool v[kEpl];
#pragma unroll
    for (int t = 0; t < kEpl; t++) {
        const uint32_t idx = lane + 32 * t;
        v[t] = idx < c;
        e[t] = v[t] ? lists[idx * kQListStride + q] : kEmptySlot;
        k[t] = (uint32_t)(e[t] >> 32);
    }
    __syncwarp();
    uint32_t prefix = 0, remaining = keep;
#pragma unroll 1
    for (int bit = 31; bit >= 0; bit--) {
        const uint32_t hi_mask = bit == 31 ? 0u : ~((2u << bit) - 1u); // the bits already decided
        uint32_t zeros = 0;
#pragma unroll
        for (int t = 0; t < kEpl; t++) {
            const bool z = v[t] && ((k[t] ^ prefix) & hi_mask) == 0 && !((k[t] >> bit) & 1u);
            zeros += __popc(__ballot_sync(0xFFFFFFFFu, z));
        }
        if (zeros < remaining) {
            remaining -= zeros;
            prefix |= 1u << bit;
        }
    }
    // prefix = keep-th smallest key; `remaining` (>= 1) entries equal to it are still needed
    const uint32_t lt = (1u << lane) - 1u;
    uint32_t base = 0;
#pragma unroll
    for (int t = 0; t < kEpl; t++) {
        const bool less = v[t] && k[t] < prefix;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, less);
        if (less) lists[(base + __popc(m & lt)) * kQListStride + q] = e[t];
        base += __popc(m);
    }
    // entries tied with the threshold key: the `remaining` LOWEST row ids stay (the exact scan admits rows in id order
    // with a strict `<`), found by a second radix select on the low word when there are more ties than places
    uint32_t n_eq = 0;
#pragma unroll
    for (int t = 0; t < kEpl; t++) n_eq += __popc(__ballot_sync(0xFFFFFFFFu, v[t] && k[t] == prefix));
    uint32_t row_cut = 0xFFFFFFFFu; // keep tied entries with row <= row_cut
    if (n_eq > remaining) {
        uint32_t rp = 0, rem = remaining;
#pragma unroll 1
        for (int bit = 31; bit >= 0; bit--) {
            const uint32_t hi_mask = bit == 31 ? 0u : ~((2u << bit) - 1u);
            uint32_t zeros = 0;
#pragma unroll
            for (int t = 0; t < kEpl; t++) {
                const uint32_t row = (uint32_t)e[t];
                const bool z = v[t] && k[t] == prefix && ((row ^ rp) & hi_mask) == 0 && !((row >> bit) & 1u);
                zeros += __popc(__ballot_sync(0xFFFFFFFFu, z));
            }
            if (zeros < rem) {
                rem -= zeros;
                rp |= 1u << bit;
            }
        }
        row_cut = rp; // row ids are unique: exactly `remaining` tied entries have row <= rp
    }
#pragma unroll
    for (int t = 0; t < kEpl; t++) {
        const bool eq = v[t] && k[t] == prefix && (uint32_t)e[t] <= row_cut;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, eq);
        if (eq) lists[(base + __popc(m & lt)) * kQListStride + q] = e[t];
        base += __popc(m);
    }
    __syncwarp();
    return prefix;
}

// kDirect = false: fp32 corpus, rows come from the tiled fp16 shadow (bulk copies), output = sorted candidate lists for
//                  the exact rescoring + proof.
// kDirect = true : fp16 / bf16 corpus, rows come straight from the row-major corpus through a 128B-swizzle tensor
//                  map; the fp32-accumulated products of the stored 16-bit values ARE the distances (the reference's
//                  own tiers differ by more: SURVEY.md finding 5, bar 1e-2), so each CTA keeps its exact top-`keep`
//                  and the lists go to final_select unsorted.
//   kOp = 0  16-bit float operands, fp32 accumulators: distance 1 - dot
//   kOp = 1  int8 / uint8 operands (kind::i8), int32 accumulators, inner product: (float)(1 - dot)  (IP.cpp:248-252)
//   kOp = 2  the same, cosine: 1 - (float)dot / (norm_row * norm_query), norms = the fp32 stored after the payload
//            (IP.cpp:264-271).  The integer dot products are exact, so kOp 1/2 reproduce the reference bit for bit.
//   kOp = 3  16-bit float operands, squared L2 from the GEMM: (|q|^2 + |row|^2) - 2 dot, with the squared norms of
//            the fp32 rows / queries in row_norm2 / q_norm2 (coarse stage of the fp32 L2 route)
//   kFixed   the admission threshold of every query is FIXED for the whole pass (thr_fixed[q], a distance, from the sample
//            pass): every row with approximate distance < thr_fixed[q] is kept — no running threshold, no list compaction;
//            a list that runs full sets overflow[q] (the query goes to the next tier).  fp32 route only (kOp 0 / 3).
//   kSample  the sample pass: no lists at all — every thread keeps the smallest approximate distance it has seen in each of
//            8 interleaved slices of its row range (chunk of the tile x tile parity) and publishes those 8 values as
//            keep = 8 composites per (query, row range).  The k-th smallest of a query's lists x 8 minima is an upper
//            bound of its k-th best distance over the sample (k distinct rows are at or below it), which is all the
//            fixed bound needs.  (An earlier version ran the adaptive lists here: with ~10 tiles per CTA they never left
//            their warm-up — 435 us for 1 % of the rows, ncu launch list profiles/r2c_launches.md.)
//   tile_stride > 1: only every tile_stride-th row tile is visited (the sample pass)
//   kPair    the two query groups of a row range (cluster of 2) run as a CTA pair: ONE tcgen05.mma.cta_group::2 stream issued
//            by the leader computes both groups (M = 256), each CTA streams only ITS half of every row tile (rows 64 * rank ..)
//            into its own shared memory — no multicast, half the L2 -> SM and shared-memory operand traffic per SM.  The
//            kernel runs at the board power limit (ncu: 1.37 GHz inside a launch), so operand traffic is clock.  fp32 main pass.
template <bool kDirect, int kEpl, int kOp, int kMode, bool kPair = false>
__global__ void __launch_bounds__(kCoarseThreads, 1)
coarse_qtmem_kernel(const __grid_constant__ CUtensorMap map_rows, const uint8_t *__restrict__ shadow, size_t row_pitch,
                    const uint8_t *__restrict__ q16, size_t q16_pitch, const float *__restrict__ row_norm2,
                    const float *__restrict__ q_norm2, uint32_t n_rows, uint32_t nq, uint32_t dim, uint32_t row_bytes,
                    uint32_t num_kb, uint32_t tiles_total, uint32_t keep, uint32_t nstages, uint32_t csize, uint32_t nacc, uint32_t idesc,
                    uint64_t *__restrict__ list_scratch, uint64_t *__restrict__ cand_out, const uint32_t *__restrict__ nq_dev,
                    uint32_t tile_stride, const float *__restrict__ thr_fixed, uint32_t *__restrict__ overflow) {
    constexpr bool kFixed = kMode == 1, kSample = kMode == 2;
    // logical coordinates: bx = row range, by = query group.  A CTA pair must be two CTAs adjacent along x of the cluster (the
    // hardware pairs cluster ranks 2i / 2i + 1 of the x dimension: a (1, 2, 1) cluster fails to launch with "cluster
    // misconfiguration"), so the pair build is launched as grid (2, row ranges) with cluster (2, 1, 1) and the roles of x / y swap.
    const uint32_t bx = kPair ? blockIdx.y : blockIdx.x, by = kPair ? blockIdx.x : blockIdx.y;
    const uint32_t gx = kPair ? gridDim.y : gridDim.x;
    static_assert(kMode == 0 || (!kDirect && (kOp == 0 || kOp == 3)) || (kDirect && kOp == 0),
                  "fixed bound / sample pass: the fp32 route and 16-bit corpora (inner product / cosine)");
    static_assert(!kPair || (!kDirect && kMode == 1), "CTA pairs: the fp32 main pass");
    constexpr int kSliceSets = (int)kCoarseSampleSlices / (kQN / 32); // tiles i, i + kSliceSets, ... share a slice set
    constexpr int kQListCap = kEpl * 32;
    if (nq_dev) { // second tier: the number of live queries is only known on the device; nothing to do = every CTA leaves
        nq = min(nq, *nq_dev);
        if (nq == 0) return;
    }
    constexpr uint32_t kQTrigger = kQListCap - 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // K blocks of the queries in tensor memory (all 512 columns minus the nacc accumulator stages); the rest in sQ
    const uint32_t kb_tmem = min(num_kb, (512u - nacc * kQN) / 32u);
    uint8_t *sB = smem;                                            // nstages x kQKbPerStage x [128 x 128B]
    // a pair member holds 64 rows of every K block: half-size blocks and stages, twice as many of them in the ring (the stage
    // hand-over of a pair is longer — the peer's arrival is relayed to the leader — so the ring has to look further ahead)
    constexpr uint32_t kBlkB = kPair ? kQBlockBytes / 2 : kQBlockBytes, kStgB = kPair ? kQStageBytes / 2 : kQStageBytes;
    uint8_t *sQ = sB + (size_t)nstages * kStgB;                    // (num_kb - kb_tmem) x [128 queries x 128B], swizzled
    uint64_t *bars = reinterpret_cast<uint64_t *>(sQ + (size_t)(num_kb - kb_tmem) * kQBlockBytes);
    uint64_t *full = bars, *empty = bars + kQMaxStages, *tfull = bars + 2 * kQMaxStages, *tempty = tfull + kAccStages;
    // this CTA's candidate lists [kQListCap][kQListStride]
    uint64_t *lists = list_scratch + (size_t)(by * gx + bx) * (kQListCap * kQListStride);
    uint64_t *qbar = tempty + kAccStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(qbar + 1);
    uint64_t *pfull = reinterpret_cast<uint64_t *>(tmem_slot + 2); // kPair, leader: "the peer's half of stage s has landed"

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t q_base = by * kQM;
    const uint32_t my_tiles = (tiles_total > bx) ? (tiles_total - bx + gx - 1) / gx : 0;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nstages; s++) {
            mbar_init(&full[s], 1);
            // multicast clusters: one commit per CTA (all of them read the stage); a pair: ONE commit, multicast to both CTAs
            mbar_init(&empty[s], kPair ? 1u : csize);
            if constexpr (kPair) mbar_init(&pfull[s], 1);
        }
        for (int a = 0; a < kAccStages; a++) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], kPair ? 128 + 4 : 128); // pair: + one arrival per epilogue warp of the peer
        }
        mbar_init(qbar, kPair ? 128 + 4 : 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (kPair)
            tmem_alloc_pair(tmem_slot, 512);
        else
            tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_q = tmem_base + nacc * kQN;
    // cluster mode: the csize CTAs that share blockIdx.x (one per group of 128 queries) walk the SAME row tiles.
    // Each fetches 1/csize of every stage and multicasts it into all csize shared memories, so a tile leaves
    // HBM/L2 once per cluster (measured without it: the query groups drift apart and every tile is read from HBM
    // once per group, L2 hit rate 10%).
    const uint32_t crank = (csize > 1) ? cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    if (csize > 1) cluster_sync_all(); // peers must see initialised barriers before anything is signalled remotely

    if (warp == 0) {
        // ===== producer: row tiles [64 rows x 64 halves] per K block, kQKbPerStage K blocks per stage =====
        uint32_t s = 0, ph = 0;
        const uint32_t slice_full = kQStageBytes / csize, slice_tail = ((num_kb % kQKbPerStage) * kQBlockBytes) / csize;
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t tile = (bx + i * gx) * tile_stride;
            for (uint32_t kb0 = 0; kb0 < num_kb; kb0 += kQKbPerStage) {
                const uint32_t kbn = min((uint32_t)kQKbPerStage, num_kb - kb0);
                mbar_wait_spin(&empty[s], ph ^ 1);
                if (elect_one_sync()) {
                    const uint32_t bytes = kPair ? kbn * (kQBlockBytes / 2) : kbn * kQBlockBytes; // a pair member loads half tiles
                    mbar_expect_tx(&full[s], bytes);
                    if constexpr (kDirect) {
                        // row-major 16-bit corpus: one 128B-swizzle box of (128 / csize) rows x 64 elements per K block
                        const uint32_t slice_rows = kQN / csize;
                        for (uint32_t j = 0; j < kbn; j++) {
                            uint8_t *dst = sB + (size_t)s * kQStageBytes + j * kQBlockBytes + crank * slice_rows * 128;
                            const int c0 = (int)((kb0 + j) * ((kOp == 0 || kOp == 3) ? 64 : 128)), c1 = (int)(tile * kQN + crank * slice_rows);
                            if (csize > 1)
                                tma_load_2d_mc(dst, &map_rows, &full[s], c0, c1, cmask);
                            else
                                tma_load_2d(dst, &map_rows, &full[s], c0, c1);
                        }
                    } else if constexpr (kPair) {
                        // this CTA's half of the tile: rows 64 * rank .. 64 * rank + 63 of every K block (8 KB each, a whole
                        // number of 8-row swizzle atoms), at the START of the K block's slot — the pair MMA reads 64 rows of B
                        // from each CTA at the same shared-memory offset
                        const uint8_t *src = shadow + ((size_t)tile * num_kb + kb0) * kQBlockBytes + crank * (kQBlockBytes / 2);
                        for (uint32_t j = 0; j < kbn; j++)
                            bulk_load_1d(sB + (size_t)s * kStgB + j * kBlkB, src + (size_t)j * kQBlockBytes, kQBlockBytes / 2, &full[s]);
                    } else {
                        // the shadow copy is stored tile by tile in the swizzled shared-memory image (to_f16_tiled_kernel):
                        // the K blocks of a stage are one contiguous run in HBM
                        const uint8_t *src = shadow + ((size_t)tile * num_kb + kb0) * kQBlockBytes;
                        if (csize > 1) {
                            const uint32_t slice = kbn == (uint32_t)kQKbPerStage ? slice_full : slice_tail; // bytes / csize, multiple of 16
                            bulk_load_1d_mc(sB + (size_t)s * kQStageBytes + crank * slice, src + crank * slice, slice, &full[s], cmask);
                        } else {
                            bulk_load_1d(sB + (size_t)s * kQStageBytes, src, bytes, &full[s]);
                        }
                    }
                }
                __syncwarp();
                if (++s == nstages) s = 0, ph ^= 1;
            }
        }
#define UMMA_TS(d, a, b, i, acc)                          \
    do {                                                   \
        if constexpr (kPair)                               \
            umma_ts_f16_pair(d, a, b, i, acc);             \
        else                                               \
            umma_ts<kOp>(d, a, b, i, acc);                 \
    } while (0)
#define UMMA_SS(d, a, b, i, acc)                          \
    do {                                                   \
        if constexpr (kPair)                               \
            umma_ss_f16_pair(d, a, b, i, acc);             \
        else                                               \
            umma_ss2<kOp>(d, a, b, i, acc);                \
    } while (0)
    } else if (warp == 1) {
        // ===== MMA issuer: D[128 queries x 128 rows] += Q[tmem] * rows[smem]^T =====
        // Everything that does not change from stage to stage is hoisted, and a stage takes one of two branch-free
        // bodies (all K blocks from tensor memory / all from shared memory): ncu showed this thread spending half its
        // time on descriptor arithmetic and per-MMA branches while the tensor pipe sat at 53 %.
        if (kPair && crank != 0) {
            // ===== peer of a CTA pair: no MMAs to issue — relay "my half of stage s has landed" to the leader =====
            uint32_t s = 0, ph = 0;
            const uint32_t leader_pfull = mapa_u32(smem_u32(pfull), 0);
            for (uint32_t i = 0; i < my_tiles; i++)
                for (uint32_t kb0 = 0; kb0 < num_kb; kb0 += kQKbPerStage) {
                    mbar_wait_spin(&full[s], ph);
                    if (elect_one_sync()) mbar_arrive_remote(leader_pfull + s * 8);
                    __syncwarp();
                    if (++s == nstages) s = 0, ph ^= 1;
                }
        } else {
        mbar_wait(qbar, 0); // pair: both CTAs' queries are in place (the peer's warps arrive remotely)
        tc_fence_after();
        const uint32_t idesc_mma = kPair ? ((idesc & ~(0x1Fu << 24)) | (16u << 24)) : idesc; // M = 256 across the pair
        uint32_t s = 0, ph = 0;
        const uint64_t bdesc_first = make_smem_desc(smem_u32(sB));
        const uint64_t adesc_first = make_smem_desc(smem_u32(sQ));
        constexpr uint64_t kStageStep = kStgB >> 4, kBlockStep = kBlkB >> 4; // B descriptors; the query blocks in sQ keep 16 KB
        constexpr uint64_t kQBlockStep = kQBlockBytes >> 4;
        uint64_t bdesc_s = bdesc_first; // descriptor of ring stage s
        const bool multi = csize > 1;
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t a = (nacc == 2) ? (i & 1u) : 0u, aph = (nacc == 2) ? ((i >> 1) & 1u) : (i & 1u);
            mbar_wait_spin(&tempty[a], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + a * kQN;
            for (uint32_t kb0 = 0; kb0 < num_kb; kb0 += kQKbPerStage) {
                mbar_wait_spin(&full[s], ph);
                // ... and the peer's half.  A plain (CTA-scope) wait, as CUTLASS's ClusterBarrier does for remotely signalled
                // barriers: the first pair build polled with .acquire.cluster and every poll cost ~1 us (5.86 ms per pass)
                if constexpr (kPair) mbar_wait_spin(&pfull[s], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    if (kb0 + kQKbPerStage <= kb_tmem) {
                        // both K blocks of the stage: queries from tensor memory (8 columns per instruction), 16 halves
                        // (32 bytes of the swizzled row tile) per instruction
                        const uint32_t at = tmem_q + kb0 * 32;
                        UMMA_TS(d_tmem, at, bdesc_s, idesc_mma, kb0 != 0);
                        UMMA_TS(d_tmem, at + 8, bdesc_s + 2, idesc_mma, 1);
                        UMMA_TS(d_tmem, at + 16, bdesc_s + 4, idesc_mma, 1);
                        UMMA_TS(d_tmem, at + 24, bdesc_s + 6, idesc_mma, 1);
                        UMMA_TS(d_tmem, at + 32, bdesc_s + kBlockStep, idesc_mma, 1);
                        UMMA_TS(d_tmem, at + 40, bdesc_s + kBlockStep + 2, idesc_mma, 1);
                        UMMA_TS(d_tmem, at + 48, bdesc_s + kBlockStep + 4, idesc_mma, 1);
                        UMMA_TS(d_tmem, at + 56, bdesc_s + kBlockStep + 6, idesc_mma, 1);
                    } else if (kb0 >= kb_tmem && kb0 + kQKbPerStage <= num_kb) {
                        // both K blocks: queries from shared memory
                        const uint64_t ad = adesc_first + (uint64_t)(kb0 - kb_tmem) * kQBlockStep;
                        UMMA_SS(d_tmem, ad, bdesc_s, idesc_mma, kb0 != 0);
                        UMMA_SS(d_tmem, ad + 2, bdesc_s + 2, idesc_mma, 1);
                        UMMA_SS(d_tmem, ad + 4, bdesc_s + 4, idesc_mma, 1);
                        UMMA_SS(d_tmem, ad + 6, bdesc_s + 6, idesc_mma, 1);
                        UMMA_SS(d_tmem, ad + kQBlockStep, bdesc_s + kBlockStep, idesc_mma, 1);
                        UMMA_SS(d_tmem, ad + kQBlockStep + 2, bdesc_s + kBlockStep + 2, idesc_mma, 1);
                        UMMA_SS(d_tmem, ad + kQBlockStep + 4, bdesc_s + kBlockStep + 4, idesc_mma, 1);
                        UMMA_SS(d_tmem, ad + kQBlockStep + 6, bdesc_s + kBlockStep + 6, idesc_mma, 1);
                    } else {
                        // a stage that straddles the tensor-memory / shared-memory split, or the odd last K block
                        const uint32_t kbn = min((uint32_t)kQKbPerStage, num_kb - kb0);
                        for (uint32_t j = 0; j < kbn; j++) {
                            const uint32_t kb = kb0 + j;
                            const uint64_t bd = bdesc_s + (uint64_t)j * kBlockStep;
                            if (kb < kb_tmem) {
                                const uint32_t at = tmem_q + kb * 32;
                                UMMA_TS(d_tmem, at, bd, idesc_mma, kb != 0);
                                UMMA_TS(d_tmem, at + 8, bd + 2, idesc_mma, 1);
                                UMMA_TS(d_tmem, at + 16, bd + 4, idesc_mma, 1);
                                UMMA_TS(d_tmem, at + 24, bd + 6, idesc_mma, 1);
                            } else {
                                const uint64_t ad = adesc_first + (uint64_t)(kb - kb_tmem) * kQBlockStep;
                                UMMA_SS(d_tmem, ad, bd, idesc_mma, kb != 0);
                                UMMA_SS(d_tmem, ad + 2, bd + 2, idesc_mma, 1);
                                UMMA_SS(d_tmem, ad + 4, bd + 4, idesc_mma, 1);
                                UMMA_SS(d_tmem, ad + 6, bd + 6, idesc_mma, 1);
                            }
                        }
                    }
                    if constexpr (kPair)
                        umma_commit_pair(&empty[s], 0b11); // both halves of stage s are free: tell both producers
                    else if (multi)
                        umma_commit_mc(&empty[s], cmask); // this CTA is done with stage s: tell every producer of the cluster
                    else
                        umma_commit(&empty[s]);
                }
                __syncwarp();
                bdesc_s += kStageStep;
                if (++s == nstages) s = 0, ph ^= 1, bdesc_s = bdesc_first;
            }
            if (elect_one_sync()) {
                if constexpr (kPair)
                    umma_commit_pair(&tfull[a], 0b11); // the accumulators of BOTH CTAs are complete
                else
                    umma_commit(&tfull[a]);
            }
            __syncwarp();
        }
        } // leader / single CTA
#undef UMMA_TS
#undef UMMA_SS
    } else if (warp >= 4) {
        const int ew = warp - 4;          // TMEM lane quadrant
        const int et = threadIdx.x - 128; // 0..127 = query slot = TMEM lane
        const uint32_t q = q_base + et;
        const uint32_t lane_addr = (uint32_t)(ew * 32) << 16;
        // ===== queries -> TMEM (zero padded to num_kb * 64 halves and to 128 queries) =====
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(q16 + (size_t)q * q16_pitch);
            for (uint32_t kb = 0; kb < num_kb; kb++) {
                uint4 x[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    x[u] = make_uint4(0, 0, 0, 0);
                    if (q < nq && kb * 128 + u * 16 < row_bytes) x[u] = src[kb * 8 + u];
                }
                if (kb < kb_tmem) {
                    uint32_t w[32];
#pragma unroll
                    for (int u = 0; u < 8; u++) w[u * 4 + 0] = x[u].x, w[u * 4 + 1] = x[u].y, w[u * 4 + 2] = x[u].z, w[u * 4 + 3] = x[u].w;
                    tmem_st32(tmem_q + lane_addr + kb * 32, w);
                } else { // row `et` of a K-major 128B-swizzled operand tile: chunk c at et*128 + ((c ^ (et & 7)) * 16)
                    uint8_t *row = sQ + (size_t)(kb - kb_tmem) * kQBlockBytes + et * 128;
#pragma unroll
                    for (int u = 0; u < 8; u++) *reinterpret_cast<uint4 *>(row + ((u ^ (et & 7)) * 16)) = x[u];
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes of sQ -> tensor-core reads
            tmem_wait_st();
            tc_fence_before();
            if (kPair && crank != 0) { // the MMAs are issued by the leader: one arrival per warp on ITS barrier
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(qbar), 0));
            } else {
                mbar_arrive(qbar);
            }
        }
        // ===== epilogue: this thread's query against 128 rows per tile =====
        uint32_t thr = 0xFFFFFFFFu, cnt = 0;
        // pre-test bound in the raw accumulator domain; -inf: everything passes until the first compaction
        float thr_dot = -__int_as_float(0x7f800000);
        float nq_norm = 1.0f;
        if constexpr (kOp == 2) nq_norm = q < nq ? *reinterpret_cast<const float *>(q16 + (size_t)q * q16_pitch + dim) : 1.0f;
        if constexpr (kOp == 3) nq_norm = q < nq ? q_norm2[q] : 0.0f; // |q|^2
        bool ovf = false;
        float smax[kSample ? kSliceSets : 1][kQN / 32]; // kSample: running maxima of the pre-test value per slice (largest = smallest distance)
#pragma unroll
        for (int x = 0; x < (kSample ? kSliceSets : 1); x++)
#pragma unroll
            for (int h = 0; h < kQN / 32; h++) smax[x][h] = -__int_as_float(0x7f800000);
        if constexpr (kFixed) {
            // fixed admission bound (a distance): keep every row with approximate distance < T
            const float T = q < nq ? thr_fixed[q] : -__int_as_float(0x7f800000);
            thr = orderable_key(T);
            if constexpr (kOp == 0) { // d < T  <=>  dot > 1 - T; slack: the rounding of the two subtractions
                const float t = 1.0f - T;
                thr_dot = t - (4e-7f + 2.4e-7f * fabsf(t));
            } else {
                thr_dot = 0.5f * (nq_norm - T) - 2e-6f * (fabsf(nq_norm) + fabsf(T));
            }
            if (!(T == T)) thr_dot = -__int_as_float(0x7f800000), thr = 0xFFFFFFFFu; // NaN bound: keep everything (-> overflow -> next tier)
        }
        if (q >= nq) thr_dot = __int_as_float(0x7f800000); // padding lanes of a partial query group: nothing ever passes
        for (uint32_t i = 0; i < my_tiles; i++) {
            const uint32_t tile = (bx + i * gx) * tile_stride;
            const uint32_t a = (nacc == 2) ? (i & 1u) : 0u, aph = (nacc == 2) ? ((i >> 1) & 1u) : (i & 1u);
            float nrm[kQN / 32]; // kOp 2 / 3: lane l holds the norm / squared norm of rows h*32 + l of the tile
            if constexpr (kOp == 3) {
#pragma unroll
                for (int h = 0; h < kQN / 32; h++) {
                    const uint32_t r = tile * kQN + h * 32 + lane;
                    nrm[h] = r < n_rows ? __ldg(row_norm2 + r) : 0.0f;
                }
            }
            if constexpr (kOp == 2) {
#pragma unroll
                for (int h = 0; h < kQN / 32; h++) {
                    const uint32_t r = tile * kQN + h * 32 + lane;
                    nrm[h] = r < n_rows ? __ldg(reinterpret_cast<const float *>(shadow + (size_t)r * row_pitch + dim)) : 1.0f;
                }
            }
            mbar_wait(&tfull[a], aph);
            tc_fence_after();
            // drain the whole tile to registers and hand the accumulator stage back before looking at a value
            uint32_t v[kQN / 32][32];
#pragma unroll
            for (int h = 0; h < kQN / 32; h++) tmem_ld32_nowait(tmem_base + lane_addr + a * kQN + h * 32, v[h]);
            tmem_wait_ld();
            tc_fence_before();
            if (kPair && crank != 0) { // hand the accumulator stage back to the leader's issuer
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty[a]), 0));
            } else {
                mbar_arrive(&tempty[a]);
            }
#pragma unroll
            for (int h = 0; h < kQN / 32; h++) {
                const uint32_t row0 = tile * kQN + h * 32;
                uint32_t pass_pre = 0;
                if constexpr (kSample) {
                    // largest dot (cosine / inner product) or largest dot - |row|^2 / 2 (squared L2) of the chunk
                    float mx = -__int_as_float(0x7f800000);
                    const bool tail = row0 + 32 > n_rows; // rows past the end (zero fill / stale shadow bytes) must not count
                    if constexpr (kOp == 0) {
                        if (!tail) {
                            float m8[8];
#pragma unroll
                            for (int g = 0; g < 8; g++)
                                m8[g] = fmaxf(fmaxf(__uint_as_float(v[h][4 * g]), __uint_as_float(v[h][4 * g + 1])),
                                              fmaxf(__uint_as_float(v[h][4 * g + 2]), __uint_as_float(v[h][4 * g + 3])));
                            mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (row0 + j < n_rows) mx = fmaxf(mx, __uint_as_float(v[h][j]));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            const float u = fmaf(__shfl_sync(0xFFFFFFFFu, nrm[h], j), -0.5f, __uint_as_float(v[h][j]));
                            if (!tail || row0 + j < n_rows) mx = fmaxf(mx, u);
                        }
                    }
                    // slice set = tile counter mod kSliceSets: a uniform switch keeps the register indices static
#pragma unroll
                    for (int x = 0; x < kSliceSets; x++)
                        if ((int)(i % (uint32_t)kSliceSets) == x) smax[kSample ? x : 0][h] = fmaxf(smax[kSample ? x : 0][h], mx);
                    continue;
                }
                if constexpr (kFixed) {
                    // With a fixed bound only ~k * (rows / sample rows) rows of the whole corpus pass: one 3-input max tree
                    // over the 32 values (16 instructions) and ONE compare decide the common case
                    if constexpr (kOp == 0) {
                        float mx;
                        float m8[8];
#pragma unroll
                        for (int g = 0; g < 8; g++)
                            m8[g] = fmaxf(fmaxf(__uint_as_float(v[h][4 * g]), __uint_as_float(v[h][4 * g + 1])),
                                          fmaxf(__uint_as_float(v[h][4 * g + 2]), __uint_as_float(v[h][4 * g + 3])));
                        mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                        if (!(mx > thr_dot)) continue;
                    } else { // squared L2: the bound depends on the row; the shuffles need the whole warp, so the pass mask
                             // itself is built here, before the lanes diverge
#pragma unroll
                        for (int j = 0; j < 32; j++)
                            if (__uint_as_float(v[h][j]) > fmaf(__shfl_sync(0xFFFFFFFFu, nrm[h], j), 0.499999f, thr_dot)) pass_pre |= 1u << j;
                        if (pass_pre == 0) continue;
                    }
                }
                // pre-test on the raw dot product against a slightly loose bound (a few instructions per value);
                // the few survivors take the exact distance and key comparison below
                uint32_t pass = pass_pre;
#pragma unroll
                for (int j = 0; j < ((kFixed && kOp == 3) ? 0 : 32); j++) {
                    bool p;
                    if constexpr (kOp == 0)
                        p = __uint_as_float(v[h][j]) > thr_dot;
                    else if constexpr (kOp == 1)
                        p = (float)(int)v[h][j] > thr_dot;
                    else if constexpr (kOp == 2)
                        p = (float)(int)v[h][j] > __shfl_sync(0xFFFFFFFFu, nrm[h], j) * thr_dot - 0.01f;
                    else // d < d_thr  <=>  dot > |row|^2 / 2 + (|q|^2 - d_thr) / 2, loosened for the rounding of both sides
                        p = __uint_as_float(v[h][j]) > fmaf(__shfl_sync(0xFFFFFFFFu, nrm[h], j), 0.499999f, thr_dot);
                    if (p) pass |= 1u << j;
                }
                if (row0 + 32 > n_rows) pass &= (n_rows > row0) ? ((1u << (n_rows - row0)) - 1u) : 0u; // TMA zero fill past the end
                while (pass) {
                    const int j = __ffs(pass) - 1;
                    pass &= pass - 1;
                    uint32_t raw = 0;
#pragma unroll
                    for (int x = 0; x < 32; x++)
                        if (x == j) raw = v[h][x];
                    float d;
                    if constexpr (kOp == 0)
                        d = 1.0f - __uint_as_float(raw);
                    else if constexpr (kOp == 1)
                        d = (float)(1 - (int)raw);
                    else if constexpr (kOp == 2) {
                        const uint32_t r = row0 + j; // rare path: re-read the norm (L1/L2 hit) instead of a divergent shuffle
                        const float nr = __ldg(reinterpret_cast<const float *>(shadow + (size_t)r * row_pitch + dim));
                        d = __fsub_rn(1.0f, __fdiv_rn((float)(int)raw, __fmul_rn(nr, nq_norm)));
                    } else {
                        d = __fsub_rn(__fadd_rn(nq_norm, __ldg(row_norm2 + row0 + j)), __fmul_rn(2.0f, __uint_as_float(raw)));
                    }
                    const uint32_t key = orderable_key(d);
                    if (key < thr) {
                        if constexpr (kFixed) {
                            if (cnt < (uint32_t)kQListCap) {
                                lists[cnt * kQListStride + et] = ((uint64_t)key << 32) | (row0 + j);
                                cnt++;
                            } else {
                                ovf = true; // more rows below the bound than the list holds: the query goes to the next tier
                            }
                        } else {
                            lists[cnt * kQListStride + et] = ((uint64_t)key << 32) | (row0 + j);
                            cnt++;
                        }
                    }
                }
                if constexpr (kFixed) continue;
                // lists that ran past the trigger are cut back to the best `keep` by the whole warp
                uint32_t m = __ballot_sync(0xFFFFFFFFu, cnt > kQTrigger);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t c = __shfl_sync(0xFFFFFFFFu, cnt, src);
                    const uint32_t worst = select_keep<kEpl>(lists, ew * 32 + src, c, keep, lane);
                    if (lane == src) {
                        cnt = keep;
                        thr = worst;
                        // d < d_thr needs dot > t = 1 - d_thr (times the norms for the integer cosine); the slack covers
                        // the rounding of the subtractions, of int -> float and of the norm product
                        const float t = 1.0f - key_to_float(thr);
                        if constexpr (kOp == 0)
                            thr_dot = t - 4e-7f;
                        else if constexpr (kOp == 1)
                            thr_dot = t - (fabsf(t) * 1e-6f + 2.0f);
                        else if constexpr (kOp == 2) {
                            const float tq = t * nq_norm;
                            thr_dot = tq - fabsf(tq) * 4e-6f;
                        } else {
                            const float dthr = key_to_float(thr);
                            thr_dot = 0.5f * (nq_norm - dthr) - 2e-6f * (fabsf(nq_norm) + fabsf(dthr));
                        }
                    }
                }
            }
        }
        // publish: cand_out[q][blockIdx.x][keep], kEmptySlot padded
        __syncwarp();
        if constexpr (kFixed) {
            if (ovf && q < nq) overflow[q] = 1u;
        }
        if constexpr (kSample) { // keep == 8: the slice minima as composites (the row id is not needed by threshold_kernel)
            if (q < nq) {
                uint64_t *dst = cand_out + ((size_t)q * gx + bx) * keep;
#pragma unroll
                for (int x = 0; x < kSliceSets; x++)
#pragma unroll
                    for (int h = 0; h < kQN / 32; h++) {
                        const float m = smax[kSample ? x : 0][h];
                        uint64_t c = kEmptySlot;
                        if (m > -__int_as_float(0x7f800000)) {
                            const float d = kOp == 0 ? 1.0f - m : __fsub_rn(nq_norm, __fmul_rn(2.0f, m));
                            c = (uint64_t)orderable_key(d) << 32;
                        }
                        if ((uint32_t)(x * (kQN / 32) + h) < keep) dst[x * (kQN / 32) + h] = c;
                    }
            }
        }
        for (int src = 0; src < (kSample ? 0 : 32); src++) {
            uint32_t c = __shfl_sync(0xFFFFFFFFu, cnt, src);
            const uint32_t qq = q_base + ew * 32 + src;
            if (c > keep) {
                select_keep<kEpl>(lists, ew * 32 + src, c, keep, lane);
                c = keep;
            }
            if (qq < nq) {
                // unordered: final_select (direct routes) and refine_kernel (fp32 route) take the lists in any order
                uint64_t *dst = cand_out + ((size_t)qq * gx + bx) * keep;
                __syncwarp();
                for (uint32_t r = lane; r < keep; r += 32) dst[r] = r < c ? lists[r * kQListStride + ew * 32 + src] : kEmptySlot;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (csize > 1) cluster_sync_all(); // no CTA may exit while peers can still write its shared memory / barriers
    if (warp == 2) {
        tc_fence_after();
        if constexpr (kPair)
            tmem_dealloc_pair(tmem_base, 512);
        else
            tmem_dealloc(tmem_base, 512);
    }
}

// fp32 rows -> fp16 (round to nearest even) shadow rows; 8 elements per thread, dim % 8 == 0
__global__ void __launch_bounds__(256) to_f16_kernel(const uint8_t *__restrict__ src, size_t spitch, uint32_t dim, uint32_t first,
                                                     uint32_t n, uint8_t *__restrict__ dst, size_t dpitch) {
    const uint32_t per_row = dim / 8;
    const size_t total = (size_t)n * per_row;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = first + (uint32_t)(i / per_row), c = (uint32_t)(i % per_row);
        const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)r * spitch) + 2 * c;
        const float4 x = p[0], y = p[1];
        __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
        __half2 h2 = __floats2half2_rn(y.x, y.y), h3 = __floats2half2_rn(y.z, y.w);
        uint4 o;
        o.x = *reinterpret_cast<uint32_t *>(&h0);
        o.y = *reinterpret_cast<uint32_t *>(&h1);
        o.z = *reinterpret_cast<uint32_t *>(&h2);
        o.w = *reinterpret_cast<uint32_t *>(&h3);
        reinterpret_cast<uint4 *>(dst + (size_t)r * dpitch)[c] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// stages 2 + 3 in one kernel, one CTA per query: exact rescoring of the candidates that can still matter, the exact
// top-k, and the completeness proof.
//
//   * a_k = the k-th smallest APPROXIMATE distance over all candidate lists of the query (radix select on the keys).
//     With |approx - exact| <= eps, the k best-approx candidates all have exact <= a_k + eps, so the k-th exact distance
//     e_k <= a_k + eps; a candidate with approx > a_k + 2 eps has exact > a_k + eps >= e_k and cannot be among the k
//     best: only candidates with approx <= a_k + 2 eps are re-read from the fp32 corpus (a few dozen rows instead of
//     lists x keep = 1,776) and scored with the bit-exact arithmetic of distance_core.cuh.
//   * proof (as before): a row that is NOT among the candidates of its list has approx >= the list's worst kept
//     approx a_w, hence exact >= a_w - eps.  If a_w - eps > e_k for every FULL list, nothing was missed.
//   eps: unit vectors (cosine): the constant `eps`.  Otherwise (q_norm2 != NULL) it scales with the norms — fp16 RN
//   operands give |dot error| <= eps |a| |q| (Cauchy-Schwarz; `eps` already holds the accumulation slack)
//   + 2^-24 sqrt(D) (|a| + |q|) for elements below the fp16 normal range; |a| <= max_norm for every row.  L2: twice
//   that (the -2 dot term) + the fp32 rounding of the squared norms, (D + 4) 2^-23 (max_norm^2 + |q|^2).
//   q_index (nullable): second tier — CTA i works on query q_index[i] of the original batch (its lists are stored at
//   position i); nq_dev (nullable) = number of live CTAs.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kRefineMaxSurv = 2048;

// |approx - exact| bound of one query (see above); q_norm2 == NULL: unit vectors
__device__ __forceinline__ float query_eps(float eps, const float *q_norm2, uint32_t pos, float max_norm, uint32_t dim, bool l2) {
    if (!q_norm2) return eps;
    const float qn2 = q_norm2[pos], qn = sqrtf(qn2);
    float e = eps * max_norm * qn + 5.97e-8f * sqrtf((float)dim) * (max_norm + qn);
    if (l2) e = 2.0f * e + (float)(dim + 4) * 1.2e-7f * (max_norm * max_norm + qn2);
    return e * 1.0001f;
}
// k-th smallest key (high word) among the non-empty composites of mine[0, total) — block-wide radix select, 8 bits per
// pass; 0xFFFFFFFF when there are fewer than k.  256 threads; hist[256] and ctl[4] in shared memory.
__device__ __forceinline__ uint32_t block_kth_key(const uint64_t *mine, uint32_t total, uint32_t k, uint32_t *hist, uint32_t *ctl) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) ctl[0] = 0, ctl[1] = k, ctl[2] = 0;
    __syncthreads();
    uint32_t cnt = 0;
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) cnt += mine[i] != kEmptySlot;
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if (lane == 0 && cnt) atomicAdd(&ctl[2], cnt);
    __syncthreads();
    if (ctl[2] < k) return 0xFFFFFFFFu;
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t prefix = ctl[0];
        const uint32_t hi_mask = shift == 24 ? 0u : ~((1u << (shift + 8)) - 1u);
        for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
            const uint64_t c = mine[i];
            if (c == kEmptySlot) continue;
            const uint32_t key = (uint32_t)(c >> 32);
            if (((key ^ prefix) & hi_mask) == 0) atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t rem = ctl[1], b = 0;
            for (; b < 255; b++) {
                if (hist[b] >= rem) break;
                rem -= hist[b];
            }
            ctl[0] = prefix | (b << shift);
            ctl[1] = rem;
        }
        __syncthreads();
    }
    return ctl[0];
}

// Admission bound of the main pass from the SAMPLE pass (every tile_stride-th row tile): T[q] = (k-th smallest approximate
// distance among the sample's candidates) + 2 eps.  The k-th exact distance over the whole corpus e_k is at most the
// sample's, which is <= a_k + eps, so T - eps > e_k: a row the main pass drops (approx >= T) has exact >= T - eps > e_k.
// Also clears the overflow flags.  One CTA per query.
__global__ void __launch_bounds__(256) threshold_kernel(const uint64_t *__restrict__ cand, uint32_t nq, uint32_t lists_per_query,
                                                        uint32_t keep, uint32_t k, float eps, const float *__restrict__ q_norm2,
                                                        float max_norm, uint32_t dim, int l2, float *__restrict__ thr_out,
                                                        uint32_t *__restrict__ overflow) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t ctl[4];
    const uint32_t q = blockIdx.x;
    if (q >= nq) return;
    const uint32_t ak = block_kth_key(cand + (size_t)q * lists_per_query * keep, lists_per_query * keep, k, hist, ctl);
    if (threadIdx.x == 0) {
        const float e = query_eps(eps, q_norm2, q, max_norm, dim, l2 != 0);
        float T;
        if (ak == 0xFFFFFFFFu)
            T = __int_as_float(0x7f800000);
        else if (e == 0.0f) // 16-bit corpora: the GEMM result IS the distance — keep d <= a_k (the main pass tests d < T)
            T = key_to_float(ak + 1u);
        else
            T = key_to_float(ak) + (2.0f * e) * 1.001f + 1e-30f;
        thr_out[q] = T;
        overflow[q] = 0;
    }
}

// thr_T == NULL: adaptive lists (a full list's worst kept approximate distance bounds what its row range dropped).
// thr_T != NULL: lists of the fixed-bound pass — every list holds ALL rows of its range with approx < thr_T[pos] unless
//                overflow[pos] is set; what was dropped has approx >= thr_T[pos].
template <int MT>
__global__ void __launch_bounds__(256) refine_kernel(const uint8_t *rows, size_t pitch, uint32_t dim, const uint8_t *queries,
                                                     size_t qpitch, uint32_t nq, uint32_t lists_per_query, uint32_t keep, uint32_t k,
                                                     const uint64_t *__restrict__ cand, float eps, const float *__restrict__ q_norm2,
                                                     float max_norm, uint32_t *__restrict__ ok, uint32_t ok_value, uint64_t *__restrict__ out,
                                                     const uint32_t *__restrict__ q_index, const uint32_t *__restrict__ nq_dev,
                                                     const float *__restrict__ thr_T, const uint32_t *__restrict__ overflow,
                                                     uint32_t smem_cap) {
    using Tile = DistTile<DT_F32, MT, 1, 1>;
    __shared__ uint64_t surv[kRefineMaxSurv];
    __shared__ uint32_t hist[256];
    __shared__ uint32_t ctl[4];
    __shared__ uint32_t s_nsurv, s_bad, s_ncomp;
    if (nq_dev) nq = min(nq, *nq_dev);
    if (blockIdx.x >= nq) return;
    const uint32_t q = q_index ? q_index[blockIdx.x] : blockIdx.x; // query of the original batch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint64_t *mine = cand + (size_t)blockIdx.x * lists_per_query * keep;
    uint32_t total = lists_per_query * keep;
    eps = query_eps(eps, q_norm2, blockIdx.x, max_norm, dim, MT == MT_L2);
    if (threadIdx.x == 0) s_nsurv = 0, s_bad = 0, s_ncomp = 0;
    __syncthreads();
    // the lists are mostly empty slots (fixed-bound pass: ~15 of 96 per list): pack the real candidates into shared memory
    // once, so that the selection passes below do not re-read 57 KB of global memory five times per query
    extern __shared__ uint64_t s_cand[];
    for (uint32_t i0 = 0; i0 < total; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const uint64_t c = i < total ? mine[i] : kEmptySlot;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, c != kEmptySlot);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_ncomp, (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (c != kEmptySlot) {
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < smem_cap) s_cand[pos] = c;
        }
    }
    __syncthreads();
    const bool packed = s_ncomp <= smem_cap;
    const uint64_t *mine_all = mine; // the proof of the adaptive tier walks the lists as published
    const uint32_t total_all = total;
    if (packed) {
        mine = s_cand;
        total = s_ncomp;
    }
    // a_k: k-th smallest approximate key; fewer than k candidates: everything survives
    const uint32_t ak_key = block_kth_key(mine, total, k, hist, ctl);
    // survivors: approx <= a_k + 2 eps (rounded up)
    const float cut = ak_key == 0xFFFFFFFFu ? __int_as_float(0x7f800000) : key_to_float(ak_key) + (2.0f * eps) * 1.001f + 1e-30f;
    for (uint32_t i0 = 0; i0 < total; i0 += blockDim.x) {
        const uint32_t i = i0 + threadIdx.x;
        const uint64_t c = i < total ? mine[i] : kEmptySlot;
        const bool take = c != kEmptySlot && (ak_key == 0xFFFFFFFFu || key_to_float((uint32_t)(c >> 32)) <= cut);
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, take);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(&s_nsurv, (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (take) {
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < kRefineMaxSurv) surv[pos] = c;
        }
    }
    __syncthreads();
    const uint32_t n_all = s_nsurv, n_surv = min(n_all, kRefineMaxSurv);
    // exact distances of the survivors (one warp per row)
    const uint8_t *qb[1] = {queries + (size_t)q * qpitch};
    for (uint32_t i = warp; i < n_surv; i += 8) {
        const uint32_t row = (uint32_t)surv[i];
        const uint8_t *rowb[1] = {rows + (size_t)row * pitch};
        float d[1];
        Tile::run(rowb, qb, dim, lane, d);
        __syncwarp();
        if (lane == 0) surv[i] = make_composite(d[0], row);
    }
    const uint32_t n_sort = max(32u, next_pow2(n_surv));
    __syncthreads();
    for (uint32_t i = n_surv + threadIdx.x; i < n_sort; i += blockDim.x) surv[i] = kEmptySlot;
    bitonic_sort_smem(surv, n_sort);
    for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) out[(size_t)q * k + i] = i < n_surv ? surv[i] : kEmptySlot;
    // proof
    bool bad = n_all > kRefineMaxSurv; // more candidates within 2 eps of the k-th than the buffer holds: the next tier answers
    const bool have_k = n_surv >= k;
    const float ek = have_k ? key_to_float((uint32_t)(surv[k - 1] >> 32)) : 0.0f;
    if (thr_T) {
        if (threadIdx.x == 0 && (overflow[blockIdx.x] != 0 || !have_k || !(thr_T[blockIdx.x] - eps > ek))) bad = true;
    } else {
        (void)total_all;
        for (uint32_t l = threadIdx.x; l < lists_per_query; l += blockDim.x) {
            const uint64_t *lst = mine_all + (size_t)l * keep;
            uint32_t worst = 0;
            bool full = true;
            for (uint32_t j = 0; j < keep; j++) {
                const uint64_t c = lst[j];
                if (c == kEmptySlot)
                    full = false;
                else
                    worst = max(worst, (uint32_t)(c >> 32));
            }
            if (!full) continue; // the list holds every row of its range that passed the kernel's threshold
            if (!have_k)
                bad = true; // fewer than k rows found although a list was truncated
            else if (!(key_to_float(worst) - eps > ek))
                bad = true;
        }
    }
    if (bad) atomicOr(&s_bad, 1u);
    __syncthreads();
    if (threadIdx.x == 0) ok[q] = s_bad ? 0u : ok_value; // 1 = first tier, 2 = second tier (any non-zero = proven)
}

// indices of the queries the first tier left unproven, densely packed: idx[0, *count)
__global__ void compact_unproven_kernel(const uint32_t *__restrict__ ok, uint32_t nq, uint32_t *__restrict__ idx,
                                        uint32_t *__restrict__ count) {
    __shared__ uint32_t n;
    if (threadIdx.x == 0) n = 0;
    __syncthreads();
    for (uint32_t q0 = 0; q0 < nq; q0 += blockDim.x) { // one CTA, ascending order kept chunk by chunk
        const uint32_t q = q0 + threadIdx.x;
        const bool un = q < nq && ok[q] == 0;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, un);
        __shared__ uint32_t wbase[32];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) wbase[warp] = __popc(m);
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; w++) before += wbase[w];
        uint32_t all = 0;
        for (uint32_t w = 0; w < blockDim.x / 32; w++) all += wbase[w];
        if (un) idx[n + before + __popc(m & ((1u << lane) - 1u))] = q;
        __syncthreads();
        if (threadIdx.x == 0) n += all;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = n;
}
// dst row i = src row idx[i] for i < *count (fp16 query rows and, optionally, their squared norms)
__global__ void gather_queries_kernel(const uint8_t *__restrict__ src, size_t pitch, const float *__restrict__ src_n2,
                                      const uint32_t *__restrict__ idx, const uint32_t *__restrict__ count,
                                      uint8_t *__restrict__ dst, float *__restrict__ dst_n2) {
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src + (size_t)idx[i] * pitch);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst + (size_t)i * pitch);
        for (uint32_t c = threadIdx.x; c < pitch / 16; c += blockDim.x) d4[c] = s4[c];
        if (src_n2 && threadIdx.x == 0) dst_n2[i] = src_n2[idx[i]];
    }
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

static bool make_map(CUtensorMap *m, CUtensorMapDataType dt, const void *base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                     uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    return fn(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}


static constexpr size_t kSmemLimit = 232448; // 227 KB opt-in maximum per CTA on sm_100

// accumulator stages: 2 (default) leaves 8 query K blocks in tensor memory and puts the rest in shared memory
// (SS-mode MMAs); 1 keeps up to 12 there but serialises the accumulator drain.  VECSIM_B200_ACC overrides.
static uint32_t qtmem_nacc() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("VECSIM_B200_ACC");
        v = (e && atoi(e) == 1) ? 1 : 2;
    }
    return (uint32_t)v;
}
static const void *qtmem_kernel_fn(CoarseKind kind, uint32_t epl, bool int_cos, int mode = 0, bool pair = false) {
    if (pair) // the fp32 main pass as a CTA pair (cta_group::2)
        return int_cos ? (const void *)coarse_qtmem_kernel<false, 3, 3, 1, true> : (const void *)coarse_qtmem_kernel<false, 3, 0, 1, true>;
    if (kind == CoarseDirect16) {
        if (mode == 1) return (const void *)coarse_qtmem_kernel<true, 8, 0, 1>; // fixed bound, lists of 256
        if (mode == 2) return (const void *)coarse_qtmem_kernel<true, 3, 0, 2>; // sample pass
        return epl == 3 ? (const void *)coarse_qtmem_kernel<true, 3, 0, 0> : (const void *)coarse_qtmem_kernel<true, 8, 0, 0>;
    }
    if (kind == CoarseDirect8) {
        if (int_cos) return epl == 3 ? (const void *)coarse_qtmem_kernel<true, 3, 2, 0> : (const void *)coarse_qtmem_kernel<true, 8, 2, 0>;
        return epl == 3 ? (const void *)coarse_qtmem_kernel<true, 3, 1, 0> : (const void *)coarse_qtmem_kernel<true, 8, 1, 0>;
    }
    // fp32 route (shadow rows): the flag selects the squared-L2 epilogue; epl 8 = lists of up to 128 (second tier);
    // mode 1 = fixed admission bound (lists of 96, no compaction), mode 2 = the sample pass (slice minima only)
    if (mode == 1) return int_cos ? (const void *)coarse_qtmem_kernel<false, 3, 3, 1> : (const void *)coarse_qtmem_kernel<false, 3, 0, 1>;
    if (mode == 2) return int_cos ? (const void *)coarse_qtmem_kernel<false, 3, 3, 2> : (const void *)coarse_qtmem_kernel<false, 3, 0, 2>;
    if (int_cos) return epl == 3 ? (const void *)coarse_qtmem_kernel<false, 3, 3, 0> : (const void *)coarse_qtmem_kernel<false, 8, 3, 0>;
    return epl == 3 ? (const void *)coarse_qtmem_kernel<false, 3, 0, 0> : (const void *)coarse_qtmem_kernel<false, 8, 0, 0>;
}
static size_t qtmem_fixed_smem(uint32_t num_kb) {
    const uint32_t kb_t = (512u - qtmem_nacc() * kQN) / 32u;
    const uint32_t kb_smem = num_kb > kb_t ? num_kb - kb_t : 0; // query K blocks that do not fit tensor memory
    return 1024 + (size_t)kb_smem * kQBlockBytes + (2 * kQMaxStages + 2 * kAccStages + 1) * 8 + 64 + kQMaxStages * 8 /* pfull (CTA pairs) */;
}
// fp16 with the queries in tensor memory (first 512 dims) + shared memory (the rest): keep >= 4 ring stages
static bool qtmem_fits_bytes(uint32_t row_bytes) { return qtmem_fixed_smem((row_bytes + 127) / 128) + 4 * (size_t)kQStageBytes <= kSmemLimit; }
static bool qtmem_fits(uint32_t dim) { return qtmem_fits_bytes(dim * 2); }

static size_t fixed_smem(uint32_t num_kb) {
    const uint32_t tn = CfgTF32::kTileN;
    return 1024 + (size_t)num_kb * tn * 128 + (size_t)tn * kListCap * 8 + (2 * kMaxStages + 2 * kAccStages + 1) * 8 + tn * 8 + 64;
}

bool coarse_supported(const CorpusView &c, uint32_t nq, uint32_t k, CoarseKind kind) {
    if (kind == CoarseDirect16) { // fp16 / bf16 corpora, inner product or cosine (normalised rows): tensor-core results are final
        if ((c.dtype != DT_F16 && c.dtype != DT_BF16) || c.metric != MT_IP) return false;
        if (c.dim % 8 != 0 || c.dim < 32 || c.pitch % 16 != 0 || !qtmem_fits(c.dim)) return false;
        if (k > 128 || nq < 1 || c.n_rows < 65536) return false;
        return encode_fn() != nullptr;
    }
    if (kind == CoarseDirect8) { // int8 / uint8 corpora, inner product or cosine: exact integer dot products on kind::i8
        if ((c.dtype != DT_I8 && c.dtype != DT_U8) || (c.metric != MT_IP && c.metric != MT_COS)) return false;
        if (c.dim % 16 != 0 || c.dim < 32 || c.pitch % 16 != 0 || !qtmem_fits_bytes(c.dim)) return false;
        if (k > 128 || nq < 1 || c.n_rows < 65536) return false;
        return encode_fn() != nullptr;
    }
    // fp32: cosine / inner product (distance 1 - dot) or squared L2; the caller supplies the error bound (unit vectors or norms)
    if (c.dtype != DT_F32 || (c.metric != MT_IP && c.metric != MT_L2)) return false;
    if (kind == CoarseTF32 && c.metric != MT_IP) return false;
    if (c.dim % 8 != 0 || c.dim < 32 || c.dim > 1024) return false;
    if (c.pitch % 16 != 0) return false;
    if (k > kCoarseMaxK || nq < 1) return false; // batch_scan decides whether a small batch is worth the route
    if (c.n_rows < 65536) return false; // tiny corpora: the exact kernel is already fast
    if (kind == CoarseF16 && !qtmem_fits(c.dim)) return false; // wider rows: the TF32 variant (queries in shared memory)
    if (kind == CoarseTF32 && fixed_smem((c.dim + CfgTF32::kBlockK - 1) / CfgTF32::kBlockK) + 3 * kStageBytes > kSmemLimit) return false;
    return encode_fn() != nullptr;
}

CoarsePlan plan_coarse(const CorpusView &c, uint32_t nq, CoarseKind kind, uint32_t k, uint32_t keep_override, uint32_t tile_stride,
                       int mode) {
    CoarsePlan p{};
    p.kind = kind;
    p.tile_stride = std::max(1u, tile_stride);
    p.mode = (kind == CoarseF16 || (kind == CoarseDirect16 && c.metric == MT_IP)) ? mode : 0;
    if (kind == CoarseF16 || kind == CoarseDirect16 || kind == CoarseDirect8) {
        p.num_kb = kind == CoarseDirect8 ? (c.dim + 127) / 128 : (c.dim + 63) / 64;
        p.tiles = ((c.n_rows + kQN - 1) / kQN + p.tile_stride - 1) / p.tile_stride; // row tiles this pass visits
        p.grid_y = (nq + kQM - 1) / kQM;
        const uint32_t sms = (uint32_t)device_sm_count();
        p.grid_x = std::max(1u, std::min(p.tiles, sms / p.grid_y));
        // direct routes: the CTA's exact top-k of its rows.  fp32 route: candidates per (row range, query) — kCoarseKeep
        // for k <= 16, 128 for larger k and for the second tier
        p.keep = kind == CoarseF16 ? (keep_override ? keep_override : (k <= kCoarseTier1MaxK ? kCoarseKeep : kCoarseKeepWide))
                                   : (k <= 32 ? 32u : 128u);
        p.epl = p.keep <= 32 ? 3 : 8;
        if (p.mode == 1 && kind == CoarseF16) p.keep = kCoarseFixedCap, p.epl = 3; // every row below the bound, up to the list capacity
        if (p.mode == 1 && kind == CoarseDirect16) p.keep = kCoarseFixedCapDirect, p.epl = 8;
        if (p.mode == 2) p.keep = kCoarseSampleSlices, p.epl = 3; // the slice minima
        p.stages = (uint32_t)std::min<size_t>(kQMaxStages, (kSmemLimit - qtmem_fixed_smem(p.num_kb)) / kQStageBytes);
        p.smem_bytes = qtmem_fixed_smem(p.num_kb) + (size_t)p.stages * kQStageBytes;
        // the query groups of a row range form a thread-block cluster (multicast of the row tiles)
        p.csize = 1;
        static int ccap = -1; // VECSIM_B200_CLUSTER caps the cluster size (1 = no clusters)
        if (ccap < 0) {
            const char *e = getenv("VECSIM_B200_CLUSTER");
            ccap = e ? std::max(1, atoi(e)) : 4;
        }
        for (uint32_t cs = 4; cs > 1; cs >>= 1)
            if ((int)cs <= ccap && p.grid_y % cs == 0 && (kQStageBytes / kQKbPerStage) % (16 * cs) == 0) {
                p.csize = cs;
                break;
            }
        // CTA pairs: the fp32 main pass with exactly two query groups per row range (the BASELINE batch of 256)
        static int pair_on = -1; // VECSIM_B200_PAIR=1 enables
        if (pair_on < 0) {
            const char *e = getenv("VECSIM_B200_PAIR");
            pair_on = e ? atoi(e) : 0;
        }
        p.pair = pair_on != 0 && kind == CoarseF16 && p.mode == 1 && p.csize == 2 && p.grid_y == 2;
        if (p.pair) { // half-size stages, twice as many
            p.stages = (uint32_t)std::min<size_t>(kQMaxStages, (kSmemLimit - qtmem_fixed_smem(p.num_kb)) / (kQStageBytes / 2));
            p.smem_bytes = qtmem_fixed_smem(p.num_kb) + (size_t)p.stages * (kQStageBytes / 2);
        }
        const void *kfn = qtmem_kernel_fn(kind, p.epl, kind == CoarseF16 ? c.metric == MT_L2 : c.metric == MT_COS, p.mode, p.pair);
        if (p.csize > 1) {
            cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
            cudaLaunchConfig_t cfg{};
            cudaLaunchAttribute at[1];
            cfg.gridDim = p.pair ? dim3(2, 1, 1) : dim3(1, p.grid_y, 1);
            cfg.blockDim = dim3(kCoarseThreads);
            cfg.dynamicSmemBytes = p.smem_bytes;
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = p.pair ? 2 : 1, at[0].val.clusterDim.y = p.pair ? 1 : p.csize, at[0].val.clusterDim.z = 1;
            cfg.attrs = at, cfg.numAttrs = 1;
            int nclusters = 0;
            if (cudaOccupancyMaxActiveClusters(&nclusters, kfn, &cfg) != cudaSuccess || nclusters < 1) {
                cudaGetLastError();
                p.csize = 1;
            } else {
                // one wave of co-resident clusters: grid_x row ranges x (grid_y / csize) clusters each
                p.grid_x = std::max(1u, std::min(p.grid_x, (uint32_t)nclusters / (p.grid_y / p.csize)));
            }
        }
        p.cand_elems = (size_t)nq * p.grid_x * p.keep;
        p.scratch_elems = (size_t)p.grid_x * p.grid_y * (p.epl * 32) * kQListStride;
        return p;
    }
    const uint32_t bk = CfgTF32::kBlockK, tn = CfgTF32::kTileN;
    p.num_kb = (c.dim + bk - 1) / bk;
    p.tiles = (c.n_rows + kTileM - 1) / kTileM;
    p.grid_y = (nq + tn - 1) / tn;
    const uint32_t sms = (uint32_t)device_sm_count();
    p.grid_x = std::max(1u, std::min(p.tiles, sms / p.grid_y));
    p.keep = kCoarseKeep;
    const size_t fixed_bytes = fixed_smem(p.num_kb);
    p.stages = (uint32_t)std::min<size_t>(kMaxStages, (kSmemLimit - fixed_bytes) / kStageBytes);
    p.cand_elems = (size_t)nq * p.grid_x * p.keep;
    p.smem_bytes = fixed_bytes + (size_t)p.stages * kStageBytes;
    return p;
}

template <class Cfg>
static cudaError_t launch_coarse_t(const void *rows, size_t pitch, uint32_t n_rows, uint32_t dim, const void *d_queries, size_t qpitch,
                                   uint32_t nq, const CoarsePlan &p, uint64_t *d_cand, cudaStream_t s) {
    const CUtensorMapDataType dt = Cfg::kElem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap ma, mq;
    if (!make_map(&ma, dt, rows, dim, n_rows, pitch, Cfg::kBlockK, kTileM)) return cudaErrorInvalidValue;
    if (!make_map(&mq, dt, d_queries, dim, nq, qpitch, Cfg::kBlockK, Cfg::kTileN)) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(coarse_kernel<Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    if (e != cudaSuccess) return e;
    coarse_kernel<Cfg><<<dim3(p.grid_x, p.grid_y), kCoarseThreads, p.smem_bytes, s>>>(ma, mq, n_rows, nq, p.num_kb, p.tiles, p.keep,
                                                                                     p.stages, d_cand);
    return cudaGetLastError();
}

cudaError_t launch_coarse(const CoarseOperands &o, uint32_t n_rows, uint32_t dim, uint32_t nq, const CoarsePlan &p, uint64_t *d_cand,
                          uint64_t *d_scratch, cudaStream_t s, const uint32_t *d_nq_dev, const float *d_thr_fixed, uint32_t *d_overflow) {
    if (p.kind == CoarseF16 || p.kind == CoarseDirect16 || p.kind == CoarseDirect8) {
        if (p.mode == 1 && (!d_thr_fixed || !d_overflow)) return cudaErrorInvalidValue;
        const void *kfn = qtmem_kernel_fn(p.kind, p.epl, o.int_cosine != 0, p.mode, p.pair); // (CoarseF16: the flag selects the L2 epilogue)
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
        if (e != cudaSuccess) {
            fprintf(stderr, "vecsim_b200: coarse pass: %zu bytes of shared memory refused: %s\n", p.smem_bytes, cudaGetErrorString(e));
            return e;
        }
        CUtensorMap mr{};
        // UMMA instruction descriptor: c_format [4,6) (F32 = 1, S32 = 2), a/b format [7,10)/[10,13), N>>3 [17,23), M>>4 [24,29)
        uint32_t cfmt = 1, fmt = 0; // kind::f16: F16 = 0, BF16 = 1; kind::i8: UINT8 = 0, INT8 = 1
        uint32_t row_bytes = dim * 2;
        if (p.kind == CoarseDirect16) {
            fmt = o.elem_variant ? 1u : 0u;
            if (!make_map(&mr, fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, o.rows, dim, n_rows, o.pitch, 64,
                          kQN / p.csize))
                return cudaErrorInvalidValue;
        } else if (p.kind == CoarseDirect8) {
            cfmt = 2;
            fmt = o.elem_variant ? 1u : 0u; // signed?
            row_bytes = dim;
            if (!make_map(&mr, CU_TENSOR_MAP_DATA_TYPE_UINT8, o.rows, dim, n_rows, o.pitch, 128, kQN / p.csize)) return cudaErrorInvalidValue;
        }
        const uint32_t idesc = (cfmt << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kQN >> 3) << 17) | ((uint32_t)(kQM >> 4) << 24);
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute at[1];
        cfg.gridDim = p.pair ? dim3(2, p.grid_x, 1) : dim3(p.grid_x, p.grid_y, 1); // a CTA pair lies along x (see the kernel)
        cfg.blockDim = dim3(kCoarseThreads);
        cfg.dynamicSmemBytes = p.smem_bytes;
        cfg.stream = s;
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = p.pair ? 2 : 1, at[0].val.clusterDim.y = p.pair ? 1 : p.csize, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        const uint8_t *rows = static_cast<const uint8_t *>(o.rows), *qs = static_cast<const uint8_t *>(o.queries);
        size_t rp = o.pitch, qp = o.qpitch;
        const float *rn2 = o.row_norm2, *qn2 = o.q_norm2;
        uint32_t a_nrows = n_rows, a_nq = nq, a_dim = dim, a_rb = row_bytes, a_kb = p.num_kb, a_tiles = p.tiles, a_keep = p.keep,
                 a_st = p.stages, a_cs = p.csize, a_nacc = qtmem_nacc(), a_idesc = idesc, a_stride = p.tile_stride;
        void *args[] = {&mr,    &rows,   &rp,   &qs,   &qp,     &rn2, &qn2, &a_nrows, &a_nq,      &a_dim, &a_rb,
                        &a_kb,  &a_tiles, &a_keep, &a_st, &a_cs, &a_nacc,  &a_idesc, &d_scratch, &d_cand, &d_nq_dev,
                        &a_stride, &d_thr_fixed, &d_overflow};
        const cudaError_t le = cudaLaunchKernelExC(&cfg, kfn, args);
        if (le != cudaSuccess)
            fprintf(stderr, "vecsim_b200: coarse pass launch failed (kind %d mode %d pair %d csize %u grid %u x %u smem %zu): %s\n", (int)p.kind,
                    p.mode, (int)p.pair, p.csize, p.grid_x, p.grid_y, p.smem_bytes, cudaGetErrorString(le));
        return le;
    }
    return launch_coarse_t<CfgTF32>(o.rows, o.pitch, n_rows, dim, o.queries, o.qpitch, nq, p, d_cand, s);
}

// fp32 rows -> the tiled fp16 shadow: [tile of 128 rows][K block of 64 halves][128 rows x 128 B, 128B-swizzled],
// i.e. exactly the bytes a SWIZZLE_128B tensor-map load would have produced in shared memory, so that
// coarse_qtmem_kernel can stream it with contiguous bulk copies.  One 16-byte chunk (8 halves) per thread;
// chunks past `dim` are zero.
__global__ void __launch_bounds__(256) to_f16_tiled_kernel(const uint8_t *__restrict__ src, size_t spitch, uint32_t dim, uint32_t first,
                                                           uint32_t n, uint8_t *__restrict__ dst, uint32_t num_kb) {
    const uint32_t per_row = num_kb * 8;
    const size_t total = (size_t)n * per_row;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t r = first + (uint32_t)(i / per_row), ci = (uint32_t)(i % per_row);
        uint4 o = make_uint4(0, 0, 0, 0);
        if (ci * 8 < dim) {
            const float4 *p = reinterpret_cast<const float4 *>(src + (size_t)r * spitch) + 2 * ci;
            const float4 x = p[0], y = p[1];
            __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
            __half2 h2 = __floats2half2_rn(y.x, y.y), h3 = __floats2half2_rn(y.z, y.w);
            o.x = *reinterpret_cast<uint32_t *>(&h0);
            o.y = *reinterpret_cast<uint32_t *>(&h1);
            o.z = *reinterpret_cast<uint32_t *>(&h2);
            o.w = *reinterpret_cast<uint32_t *>(&h3);
        }
        const uint32_t tile = r / kQN, rr = r % kQN, kb = ci / 8, c = ci % 8;
        uint8_t *blk = dst + ((size_t)tile * num_kb + kb) * kQBlockBytes;
        *reinterpret_cast<uint4 *>(blk + rr * 128 + ((c ^ (rr & 7)) * 16)) = o;
    }
}

size_t coarse_shadow_bytes(uint32_t rows, uint32_t dim) {
    return (size_t)((rows + kQN - 1) / kQN) * ((dim + 63) / 64) * kQBlockBytes;
}

cudaError_t launch_to_f16_tiled(const void *src, size_t spitch, uint32_t dim, uint32_t first, uint32_t n, void *dst, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t num_kb = (dim + 63) / 64;
    const size_t total = (size_t)n * num_kb * 8;
    const uint32_t grid = (uint32_t)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)device_sm_count() * 16));
    to_f16_tiled_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(src), spitch, dim, first, n, static_cast<uint8_t *>(dst), num_kb);
    return cudaGetLastError();
}

// squared norm of fp32 rows [first, first+n) -> norm2[first + r]; running maxima (as float bits: the values are >= 0,
// a NaN compares as huge and disables the route) of the squared norm and of |x| into stats[0], stats[1]
__global__ void __launch_bounds__(256) row_stats_kernel(const uint8_t *__restrict__ rows, size_t pitch, uint32_t dim, uint32_t first,
                                                        uint32_t n, float *__restrict__ norm2, uint32_t *__restrict__ stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t r = blockIdx.x * 8 + warp; r < n; r += gridDim.x * 8) {
        const float *x = reinterpret_cast<const float *>(rows + (size_t)(first + r) * pitch);
        float s = 0.0f, m = 0.0f;
        for (uint32_t i = lane; i < dim; i += 32) {
            const float v = x[i];
            s = fmaf(v, v, s);
            m = fmaxf(m, fabsf(v));
            if (v != v) m = __int_as_float(0x7f800000);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
            m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
        }
        if (lane == 0) {
            norm2[first + r] = s;
            if (stats) {
                atomicMax(&stats[0], __float_as_uint(s));
                atomicMax(&stats[1], __float_as_uint(m));
            }
        }
    }
}
cudaError_t launch_row_stats(const void *rows, size_t pitch, uint32_t dim, uint32_t first, uint32_t n, float *d_norm2, uint32_t *d_stats,
                             cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const uint32_t grid = std::max(1u, std::min((n + 7) / 8, (uint32_t)device_sm_count() * 8));
    row_stats_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(rows), pitch, dim, first, n, d_norm2, d_stats);
    return cudaGetLastError();
}

cudaError_t launch_to_f16(const void *src, size_t spitch, uint32_t dim, uint32_t first, uint32_t n, void *dst, size_t dpitch,
                          cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const size_t total = (size_t)n * (dim / 8);
    const uint32_t grid = (uint32_t)std::max<size_t>(1, std::min<size_t>((total + 255) / 256, (size_t)device_sm_count() * 16));
    to_f16_kernel<<<grid, 256, 0, s>>>(static_cast<const uint8_t *>(src), spitch, dim, first, n, static_cast<uint8_t *>(dst), dpitch);
    return cudaGetLastError();
}

cudaError_t launch_refine(const CorpusView &c, const void *d_queries, size_t qpitch, uint32_t nq, uint32_t lists_per_query, uint32_t keep,
                          uint32_t k, const uint64_t *d_cand, float eps, const float *d_q_norm2, float max_norm, uint32_t *d_ok,
                          uint64_t *d_out, const uint32_t *d_q_index, const uint32_t *d_nq_dev, cudaStream_t s, const float *d_thr_T,
                          const uint32_t *d_overflow) {
    if (nq == 0) return cudaSuccess;
    const uint32_t okv = d_q_index ? 2u : 1u;
    const uint8_t *rows = static_cast<const uint8_t *>(c.rows), *qs = static_cast<const uint8_t *>(d_queries);
    // candidates packed in shared memory: as many slots as the lists have, up to 20 KB worth
    const uint32_t smem_cap = std::min<uint32_t>(lists_per_query * keep, 2560);
    const size_t smem = (size_t)smem_cap * 8;
    if (c.metric == MT_L2)
        refine_kernel<MT_L2><<<nq, 256, smem, s>>>(rows, c.pitch, c.dim, qs, qpitch, nq, lists_per_query, keep, k, d_cand, eps, d_q_norm2,
                                                   max_norm, d_ok, okv, d_out, d_q_index, d_nq_dev, d_thr_T, d_overflow, smem_cap);
    else
        refine_kernel<MT_IP><<<nq, 256, smem, s>>>(rows, c.pitch, c.dim, qs, qpitch, nq, lists_per_query, keep, k, d_cand, eps, d_q_norm2,
                                                   max_norm, d_ok, okv, d_out, d_q_index, d_nq_dev, d_thr_T, d_overflow, smem_cap);
    return cudaGetLastError();
}

cudaError_t launch_threshold(const uint64_t *d_cand, uint32_t nq, uint32_t lists_per_query, uint32_t keep, uint32_t k, float eps,
                             const float *d_q_norm2, float max_norm, uint32_t dim, int l2, float *d_thr, uint32_t *d_overflow, cudaStream_t s) {
    if (nq == 0) return cudaSuccess;
    threshold_kernel<<<nq, 256, 0, s>>>(d_cand, nq, lists_per_query, keep, k, eps, d_q_norm2, max_norm, dim, l2, d_thr, d_overflow);
    return cudaGetLastError();
}

__global__ void flags_from_overflow_kernel(const uint32_t *__restrict__ ovf, uint32_t nq, uint32_t *__restrict__ ok) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) ok[q] = ovf[q] ? 0u : 1u;
}
cudaError_t launch_flags_from_overflow(const uint32_t *d_overflow, uint32_t nq, uint32_t *d_ok, cudaStream_t s) {
    if (nq == 0) return cudaSuccess;
    flags_from_overflow_kernel<<<(nq + 255) / 256, 256, 0, s>>>(d_overflow, nq, d_ok);
    return cudaGetLastError();
}
__global__ void scatter_rows_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ count,
                                    uint32_t k, uint64_t *__restrict__ dst, uint32_t *__restrict__ ok) {
    const uint32_t n = *count;
    for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t q = idx[i];
        for (uint32_t t = threadIdx.x; t < k; t += blockDim.x) dst[(size_t)q * k + t] = src[(size_t)i * k + t];
        if (threadIdx.x == 0) ok[q] = 2u;
    }
}
cudaError_t launch_scatter_rows(const uint64_t *d_src, const uint32_t *d_idx, const uint32_t *d_count, uint32_t max_n, uint32_t k,
                                uint64_t *d_dst, uint32_t *d_ok, cudaStream_t s) {
    if (max_n == 0) return cudaSuccess;
    scatter_rows_kernel<<<std::min(max_n, 256u), 128, 0, s>>>(d_src, d_idx, d_count, k, d_dst, d_ok);
    return cudaGetLastError();
}

cudaError_t launch_compact_unproven(const uint32_t *d_ok, uint32_t nq, uint32_t *d_idx, uint32_t *d_count, cudaStream_t s) {
    compact_unproven_kernel<<<1, 256, 0, s>>>(d_ok, nq, d_idx, d_count);
    return cudaGetLastError();
}
cudaError_t launch_gather_queries(const void *d_src, size_t pitch, const float *d_src_n2, const uint32_t *d_idx, const uint32_t *d_count,
                                  uint32_t max_n, void *d_dst, float *d_dst_n2, cudaStream_t s) {
    if (max_n == 0) return cudaSuccess;
    gather_queries_kernel<<<std::min(max_n, 256u), 128, 0, s>>>(static_cast<const uint8_t *>(d_src), pitch, d_src_n2, d_idx, d_count,
                                                               static_cast<uint8_t *>(d_dst), d_dst_n2);
    return cudaGetLastError();
}

} // namespace rsb200
