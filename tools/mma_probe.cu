// Microbenchmark: issue rate / latency of tcgen05.mma (kind::f16) as a function of N, operand source
// and accumulator dependency.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.b32 %0, 1, 0, P;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t a) {
    uint64_t d = 0;
    d |= (uint64_t)((a & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}

// mode bit0: TS (A in TMEM) else SS; nacc: number of accumulators cycled through
template <int N>
__global__ void probe(int iters, int ts, int nacc, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *sm = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (threadIdx.x < 32) {
        const uint64_t bdesc = make_smem_desc(smem_u32(sm));
        const uint64_t adesc = make_smem_desc(smem_u32(sm + 32768));
        long long t0 = clock64();
        if (elect_one_sync()) {
            for (int i = 0; i < iters; i++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t d = tb + ((i * 4 + k) % nacc) * N;
                    if (ts)
                        mma_ts(d, tb + 384 + k * 8, bdesc + 2 * k, idesc, 1);
                    else
                        mma_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        }
        __syncwarp();
        long long t1 = clock64();
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar))
                         : "memory");
        } while (!done);
        long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb));
    }
}

// The same for a CTA PAIR: tcgen05.mma.cta_group::2 with M = 256 across the two CTAs of a (2,1,1) cluster, issued by the leader
// only; B = N rows, N / 2 in each CTA's shared memory.  Both CTAs allocate (cta_group::2) and wait for the multicast commit.
__device__ __forceinline__ void mma_ts2(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ss2(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
template <int N>
__global__ void probe_pair(int iters, int ts, int nacc, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *sm = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t crank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t *)sm)[i] = 0;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tb = slot;
    constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
    if (threadIdx.x < 32) {
        const uint64_t bdesc = make_smem_desc(smem_u32(sm));
        const uint64_t adesc = make_smem_desc(smem_u32(sm + 32768));
        long long t0 = clock64();
        if (crank == 0 && elect_one_sync()) {
            for (int i = 0; i < iters; i++) {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t d = tb + ((i * 4 + k) % nacc) * N;
                    if (ts)
                        mma_ts2(d, tb + 384 + k * 8, bdesc + 2 * k, idesc, 1);
                    else
                        mma_ss2(d, adesc + 2 * k, bdesc + 2 * k, idesc, 1);
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)),
                         "h"((uint16_t)3)
                         : "memory");
        }
        __syncwarp();
        long long t1 = clock64();
        uint32_t done;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done)
                         : "r"(smem_u32(&bar))
                         : "memory");
        } while (!done);
        long long t2 = clock64();
        if (threadIdx.x == 0 && blockIdx.x == 0) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tb));
    }
}
template <int N>
void run_pair(int ts, int nacc) {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int iters = 2000;
    cudaFuncSetAttribute(probe_pair<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute at[1];
    cfg.gridDim = dim3(148, 1, 1);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 100 * 1024;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
    cfg.attrs = at, cfg.numAttrs = 1;
    int it_arg = iters;
    void *args[] = {&it_arg, &ts, &nacc, &d};
    for (int rep = 0; rep < 2; rep++) {
        cudaError_t e = cudaLaunchKernelExC(&cfg, (const void *)probe_pair<N>, args);
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("pair N=%d ts=%d nacc=%d: %s\n", N, ts, nacc, cudaGetErrorString(e));
            return;
        }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("PAIR (M=256) N=%3d %s nacc=%d: issue %.1f clk/MMA, complete %.1f clk/MMA\n", N, ts ? "TS" : "SS", nacc, (double)h[0] / (iters * 4),
           (double)h[1] / (iters * 4));
    cudaFree(d);
}

template <int N>
void run(int ts, int nacc) {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    const int iters = 2000;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; rep++) {
        probe<N><<<148, 128, 100 * 1024>>>(iters, ts, nacc, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("N=%d ts=%d nacc=%d: %s\n", N, ts, nacc, cudaGetErrorString(e));
            return;
        }
    }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d %s nacc=%d: issue %.1f clk/MMA, complete %.1f clk/MMA\n", N, ts ? "TS" : "SS", nacc, (double)h[0] / (iters * 4),
           (double)h[1] / (iters * 4));
    cudaFree(d);
}

int main() {
    for (int ts = 0; ts < 2; ts++) {
        run<32>(ts, 1);
        run<64>(ts, 1);
        run<64>(ts, 2);
        run<128>(ts, 1);
        run<128>(ts, 2);
        run<256>(ts, 1);
    }
    for (int ts = 0; ts < 2; ts++) {
        run_pair<64>(ts, 1);
        run_pair<128>(ts, 1);
        run_pair<128>(ts, 2);
        run_pair<256>(ts, 1);
    }
    return 0;
}
