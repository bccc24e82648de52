// Warp-level distance arithmetic for the FLAT scan: one warp computes an RT x QT tile of
// (stored row, query) distances, every lane owning a strided slice of the vector dimension.
//
// fp32 (the parity-critical type) reproduces the summation ORDER of the reference's AVX-512 tier,
// VS/spaces/L2/L2_AVX512F_FP32.h:21-59 and VS/spaces/IP/IP_AVX512F_FP32.h:19-56, bit for bit:
//   * that kernel keeps two 16-lane accumulators sum0/sum1 fed by alternating 16-float steps, i.e.
//     32 independent FMA chains; chain c = 16*h + j sees elements residual + 32*u + c, u = 0,1,..
//     -> exactly one chain per GPU lane, and a warp reads 128 contiguous bytes per step;
//   * dim % 32 leftovers go first: the dim%16 head is a masked MULTIPLY into sum0 (lanes j < dim%16),
//     a remaining full 16-step is an FMA into sum1 (lanes 16..31);
//   * sum0+sum1 then _mm512_reduce_add_ps == butterfly over lane^16, ^8, ^4, ^2, ^1 (float add is
//     commutative, so the xor-butterfly reproduces the tree exactly);
//   * dim < 8 takes the scalar baseline (L2.cpp:76-86, IP.cpp:185-194): sequential, unfused
//     multiply-then-add (L2_space.cpp:213-217, IP_space.cpp:448-452).
// tests/test_vecsim_parity.py checks this against oracle/_ref (the reference's own code) with
// bit-equality on AVX-512F hosts.
//
// fp16 / bf16: fp32 accumulation of exactly-converted inputs (the reference's own tiers disagree
// with each other beyond 1e-3 here; tolerance 1e-2), 16-byte vector loads.
// int8 / uint8: exact int32 accumulation with dp4a — bit-exact whatever the order
// (VS/spaces/IP/IP.cpp:248-285, VS/spaces/L2/L2.cpp:150-174).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include "vecsim_kernels.h"

namespace rsb200 {

// ---------------------------------------------------------------------------------------------
// Transposing warp reduction of V per-lane partials (V a power of two).  After the call:
//   V <= 32: v[0] is the full sum of value index (lane >> (5 - log2 V)); lanes sharing an index all
//            hold it.
//   V  > 32: v[t], t < V/32, is the full sum of value index lane * (V/32) + t.
// Add order per value: ((p[l] + p[l^16]) + (.. ^8)) ... — the _mm512_reduce_add_ps tree.
// ---------------------------------------------------------------------------------------------
template <typename A, int V>
__device__ __forceinline__ void warp_transpose_reduce(A (&v)[V], int lane) {
    int c = V;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        if (c > 1) {
            const int half = c >> 1;
            const bool upper = (lane & m) != 0;
#pragma unroll
            for (int i = 0; i < V / 2; i++) {
                if (i < half) {
                    A send = upper ? v[i] : v[i + half];
                    A keep = upper ? v[i + half] : v[i];
                    A recv = __shfl_xor_sync(0xffffffffu, send, m);
                    v[i] = keep + recv;
                }
            }
            c = half;
        } else {
            v[0] = v[0] + __shfl_xor_sync(0xffffffffu, v[0], m);
        }
    }
}

template <int V>
struct TileMap { // which (row i, query j) of the RT x QT tile a lane ends up holding
    static constexpr int kLog = (V >= 64) ? 6 : (V >= 32) ? 5 : (V >= 16) ? 4 : (V >= 8) ? 3 : (V >= 4) ? 2 : (V >= 2) ? 1 : 0;
    static constexpr int kPerLane = (V > 32) ? V / 32 : 1;
    __device__ static __forceinline__ int value_index(int lane, int t) {
        return (V > 32) ? lane * kPerLane + t : (lane >> (5 - kLog));
    }
    __device__ static __forceinline__ bool primary(int lane) {
        return (V >= 32) ? true : ((lane & ((1 << (5 - kLog)) - 1)) == 0);
    }
};

__device__ __forceinline__ float load_f32_unaligned(const uint8_t *p) {
    uint32_t u = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    return __uint_as_float(u);
}

// ---------------------------------------------------------------------------------------------
// DistTile<DT, MT, RT, QT>::run — all lanes call with RT row pointers / QT query pointers
// (bytes, stored form).  Result: out[t] for t < kPerLane, tile position via TileMap<RT*QT>.
// ---------------------------------------------------------------------------------------------
template <int DT, int MT, int RT, int QT>
struct DistTile;

// ----------------------------------------- fp32 ------------------------------------------------
template <int MT, int RT, int QT>
struct DistTile<DT_F32, MT, RT, QT> {
    static constexpr int V = RT * QT;
    using Map = TileMap<V>;
    __device__ static __forceinline__ float term_mul(float x, float y) {
        if (MT == MT_L2) {
            float d = __fsub_rn(x, y);
            return __fmul_rn(d, d);
        }
        return __fmul_rn(x, y);
    }
    __device__ static __forceinline__ float term_fma(float x, float y, float acc) {
        if (MT == MT_L2) {
            float d = __fsub_rn(x, y);
            return __fmaf_rn(d, d, acc);
        }
        return __fmaf_rn(x, y, acc);
    }
    __device__ static __forceinline__ void run(const uint8_t *const (&rowb)[RT], const uint8_t *const (&qb)[QT],
                                               uint32_t dim, int lane, float (&out)[Map::kPerLane]) {
        const float *rp[RT];
        const float *qp[QT];
#pragma unroll
        for (int i = 0; i < RT; i++) rp[i] = reinterpret_cast<const float *>(rowb[i]);
#pragma unroll
        for (int j = 0; j < QT; j++) qp[j] = reinterpret_cast<const float *>(qb[j]);
        float acc[V];
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = 0.0f;

        if (dim < 8) {
            // scalar baseline: res += t*t, one rounding per operation, element order 0..dim-1.
            // Lane 0 carries the value, the others carry +0 so the butterfly is an identity.
            if (lane == 0) {
                for (uint32_t e = 0; e < dim; e++) {
                    float x[RT], y[QT];
#pragma unroll
                    for (int i = 0; i < RT; i++) x[i] = rp[i][e];
#pragma unroll
                    for (int j = 0; j < QT; j++) y[j] = qp[j][e];
#pragma unroll
                    for (int i = 0; i < RT; i++)
#pragma unroll
                        for (int j = 0; j < QT; j++) acc[i * QT + j] = __fadd_rn(acc[i * QT + j], term_mul(x[i], y[j]));
                }
            }
        } else {
            const uint32_t res = dim & 31u, r16 = res & 15u;
            if (r16 != 0 && (uint32_t)lane < r16) { // masked multiply into sum0
                float x[RT], y[QT];
#pragma unroll
                for (int i = 0; i < RT; i++) x[i] = rp[i][lane];
#pragma unroll
                for (int j = 0; j < QT; j++) y[j] = qp[j][lane];
#pragma unroll
                for (int i = 0; i < RT; i++)
#pragma unroll
                    for (int j = 0; j < QT; j++) acc[i * QT + j] = term_mul(x[i], y[j]);
            }
            if (res >= 16 && lane >= 16) { // the odd full 16-step goes to sum1
                const uint32_t e = r16 + (uint32_t)lane - 16u;
                float x[RT], y[QT];
#pragma unroll
                for (int i = 0; i < RT; i++) x[i] = rp[i][e];
#pragma unroll
                for (int j = 0; j < QT; j++) y[j] = qp[j][e];
#pragma unroll
                for (int i = 0; i < RT; i++)
#pragma unroll
                    for (int j = 0; j < QT; j++) acc[i * QT + j] = term_fma(x[i], y[j], acc[i * QT + j]);
            }
            const uint32_t nchunks = (dim - res) >> 5;
            const uint32_t off = res + (uint32_t)lane;
            constexpr int U = (V <= 8) ? 8 : (V <= 16 ? 4 : 2);
            uint32_t u = 0;
            for (; u + U <= nchunks; u += U) {
                float x[U][RT], y[U][QT];
#pragma unroll
                for (int s = 0; s < U; s++) {
#pragma unroll
                    for (int i = 0; i < RT; i++) x[s][i] = __ldg(rp[i] + off + ((u + s) << 5));
#pragma unroll
                    for (int j = 0; j < QT; j++) y[s][j] = qp[j][off + ((u + s) << 5)];
                }
#pragma unroll
                for (int s = 0; s < U; s++)
#pragma unroll
                    for (int i = 0; i < RT; i++)
#pragma unroll
                        for (int j = 0; j < QT; j++) acc[i * QT + j] = term_fma(x[s][i], y[s][j], acc[i * QT + j]);
            }
            for (; u < nchunks; u++) {
                float x[RT], y[QT];
#pragma unroll
                for (int i = 0; i < RT; i++) x[i] = __ldg(rp[i] + off + (u << 5));
#pragma unroll
                for (int j = 0; j < QT; j++) y[j] = qp[j][off + (u << 5)];
#pragma unroll
                for (int i = 0; i < RT; i++)
#pragma unroll
                    for (int j = 0; j < QT; j++) acc[i * QT + j] = term_fma(x[i], y[j], acc[i * QT + j]);
            }
        }
        warp_transpose_reduce<float, V>(acc, lane);
#pragma unroll
        for (int t = 0; t < Map::kPerLane; t++) out[t] = (MT == MT_L2) ? acc[t] : __fsub_rn(1.0f, acc[t]);
    }
};

// ------------------------------------- fp16 / bf16 ----------------------------------------------
template <int DT>
__device__ __forceinline__ void unpack8(const uint4 &raw, float (&f)[8]) {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (DT == DT_F16) {
            float2 p = __half22float2(*reinterpret_cast<const __half2 *>(&w[i]));
            f[2 * i] = p.x;
            f[2 * i + 1] = p.y;
        } else { // bf16 -> fp32 is a 16-bit shift (VS/types/bfloat16.h:31-38)
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
}
template <int DT>
__device__ __forceinline__ float load16(const uint8_t *p, uint32_t e) {
    uint16_t bits = reinterpret_cast<const uint16_t *>(p)[e];
    if (DT == DT_F16) return __half2float(__ushort_as_half(bits));
    return __uint_as_float((uint32_t)bits << 16);
}

template <int DT, int MT, int RT, int QT>
struct DistTile16 {
    static constexpr int V = RT * QT;
    using Map = TileMap<V>;
    __device__ static __forceinline__ void run(const uint8_t *const (&rowb)[RT], const uint8_t *const (&qb)[QT],
                                               uint32_t dim, int lane, float (&out)[Map::kPerLane]) {
        float acc[V];
#pragma unroll
        for (int v = 0; v < V; v++) acc[v] = 0.0f;
        const uint32_t nvec = dim >> 3; // 8 elements = 16 bytes per lane per step
        for (uint32_t vi = lane; vi < nvec; vi += 32) {
            uint4 xr[RT], yr[QT];
#pragma unroll
            for (int i = 0; i < RT; i++) xr[i] = __ldg(reinterpret_cast<const uint4 *>(rowb[i]) + vi);
#pragma unroll
            for (int j = 0; j < QT; j++) yr[j] = reinterpret_cast<const uint4 *>(qb[j])[vi];
            float yf[QT][8];
#pragma unroll
            for (int j = 0; j < QT; j++) unpack8<DT>(yr[j], yf[j]);
#pragma unroll
            for (int i = 0; i < RT; i++) {
                float xf[8];
                unpack8<DT>(xr[i], xf);
#pragma unroll
                for (int j = 0; j < QT; j++)
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        if (MT == MT_L2) {
                            float d = xf[e] - yf[j][e];
                            acc[i * QT + j] = fmaf(d, d, acc[i * QT + j]);
                        } else {
                            acc[i * QT + j] = fmaf(xf[e], yf[j][e], acc[i * QT + j]);
                        }
                    }
            }
        }
        for (uint32_t e = (nvec << 3) + lane; e < dim; e += 32) {
#pragma unroll
            for (int i = 0; i < RT; i++) {
                float x = load16<DT>(rowb[i], e);
#pragma unroll
                for (int j = 0; j < QT; j++) {
                    float y = load16<DT>(qb[j], e);
                    if (MT == MT_L2) {
                        float d = x - y;
                        acc[i * QT + j] = fmaf(d, d, acc[i * QT + j]);
                    } else {
                        acc[i * QT + j] = fmaf(x, y, acc[i * QT + j]);
                    }
                }
            }
        }
        warp_transpose_reduce<float, V>(acc, lane);
#pragma unroll
        for (int t = 0; t < Map::kPerLane; t++) out[t] = (MT == MT_L2) ? acc[t] : 1.0f - acc[t];
    }
};
template <int MT, int RT, int QT>
struct DistTile<DT_F16, MT, RT, QT> : DistTile16<DT_F16, MT, RT, QT> {};
template <int MT, int RT, int QT>
struct DistTile<DT_BF16, MT, RT, QT> : DistTile16<DT_BF16, MT, RT, QT> {};

// ------------------------------------- int8 / uint8 ---------------------------------------------
template <int DT>
__device__ __forceinline__ int dot4(uint32_t a, uint32_t b, int c) {
    if (DT == DT_I8) return __dp4a((int)a, (int)b, c);
    return (int)__dp4a(a, b, (uint32_t)c);
}
template <int DT>
__device__ __forceinline__ int load8(const uint8_t *p, uint32_t e) {
    if (DT == DT_I8) return (int)reinterpret_cast<const int8_t *>(p)[e];
    return (int)p[e];
}

template <int DT, int MT, int RT, int QT>
struct DistTile8 {
    static constexpr int V = RT * QT;
    using Map = TileMap<V>;
    __device__ static __forceinline__ void run(const uint8_t *const (&rowb)[RT], const uint8_t *const (&qb)[QT],
                                               uint32_t dim, int lane, float (&out)[Map::kPerLane]) {
        int ip[V];
        int aa[RT], bb[QT]; // only for L2: sum a^2, sum b^2
#pragma unroll
        for (int v = 0; v < V; v++) ip[v] = 0;
#pragma unroll
        for (int i = 0; i < RT; i++) aa[i] = 0;
#pragma unroll
        for (int j = 0; j < QT; j++) bb[j] = 0;
        const uint32_t nvec = dim >> 4; // 16 bytes per lane per step
        for (uint32_t vi = lane; vi < nvec; vi += 32) {
            uint4 xr[RT], yr[QT];
#pragma unroll
            for (int i = 0; i < RT; i++) xr[i] = __ldg(reinterpret_cast<const uint4 *>(rowb[i]) + vi);
#pragma unroll
            for (int j = 0; j < QT; j++) yr[j] = reinterpret_cast<const uint4 *>(qb[j])[vi];
#pragma unroll
            for (int i = 0; i < RT; i++) {
                const uint32_t xa[4] = {xr[i].x, xr[i].y, xr[i].z, xr[i].w};
#pragma unroll
                for (int j = 0; j < QT; j++) {
                    const uint32_t ya[4] = {yr[j].x, yr[j].y, yr[j].z, yr[j].w};
#pragma unroll
                    for (int w = 0; w < 4; w++) ip[i * QT + j] = dot4<DT>(xa[w], ya[w], ip[i * QT + j]);
                }
                if (MT == MT_L2) {
#pragma unroll
                    for (int w = 0; w < 4; w++) aa[i] = dot4<DT>(xa[w], xa[w], aa[i]);
                }
            }
            if (MT == MT_L2) {
#pragma unroll
                for (int j = 0; j < QT; j++) {
                    const uint32_t ya[4] = {yr[j].x, yr[j].y, yr[j].z, yr[j].w};
#pragma unroll
                    for (int w = 0; w < 4; w++) bb[j] = dot4<DT>(ya[w], ya[w], bb[j]);
                }
            }
        }
        for (uint32_t e = (nvec << 4) + lane; e < dim; e += 32) {
            int x[RT], y[QT];
#pragma unroll
            for (int i = 0; i < RT; i++) x[i] = load8<DT>(rowb[i], e);
#pragma unroll
            for (int j = 0; j < QT; j++) y[j] = load8<DT>(qb[j], e);
#pragma unroll
            for (int i = 0; i < RT; i++) {
#pragma unroll
                for (int j = 0; j < QT; j++) ip[i * QT + j] += x[i] * y[j];
                if (MT == MT_L2) aa[i] += x[i] * x[i];
            }
            if (MT == MT_L2) {
#pragma unroll
                for (int j = 0; j < QT; j++) bb[j] += y[j] * y[j];
            }
        }
        if (MT == MT_L2) {
            // sum (a-b)^2 = sum a^2 + sum b^2 - 2 sum ab, exact in int32 like the reference's int
            // accumulator (L2.cpp:134-174); combine per lane, then reduce once.
#pragma unroll
            for (int i = 0; i < RT; i++)
#pragma unroll
                for (int j = 0; j < QT; j++) ip[i * QT + j] = aa[i] + bb[j] - 2 * ip[i * QT + j];
        }
        warp_transpose_reduce<int, V>(ip, lane);
        float nrs[RT], nqs[QT];
        if (MT == MT_COS) { // fp32 norm stored right after the dim payload bytes (IP.cpp:264-268)
#pragma unroll
            for (int i = 0; i < RT; i++) nrs[i] = load_f32_unaligned(rowb[i] + dim);
#pragma unroll
            for (int j = 0; j < QT; j++) nqs[j] = load_f32_unaligned(qb[j] + dim);
        }
#pragma unroll
        for (int t = 0; t < Map::kPerLane; t++) {
            if (MT == MT_L2) {
                out[t] = (float)ip[t];
            } else if (MT == MT_IP) {
                out[t] = (float)(1 - ip[t]); // IP.cpp:248-252: `1 - int`, then int->float
            } else {
                const int idx = Map::value_index(lane, t);
                float nr = 0.0f, nq = 0.0f;
#pragma unroll
                for (int i = 0; i < RT; i++)
                    if (idx / QT == i) nr = nrs[i];
#pragma unroll
                for (int j = 0; j < QT; j++)
                    if (idx % QT == j) nq = nqs[j];
                out[t] = __fsub_rn(1.0f, __fdiv_rn((float)ip[t], __fmul_rn(nr, nq))); // IP.cpp:264-271
            }
        }
    }
};
template <int MT, int RT, int QT>
struct DistTile<DT_I8, MT, RT, QT> : DistTile8<DT_I8, MT, RT, QT> {};
template <int MT, int RT, int QT>
struct DistTile<DT_U8, MT, RT, QT> : DistTile8<DT_U8, MT, RT, QT> {};

} // namespace rsb200
