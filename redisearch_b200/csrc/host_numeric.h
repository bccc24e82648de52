// Host-side scalar arithmetic that must be BIT-IDENTICAL to the reference because its results are
// stored (normalised vectors, appended norms) and then compared by the kernels:
//   fp16 <-> fp32   VS/types/float16.h:33-117   (fp32->fp16 is the reference's "round half up on the
//                                                 13th bit via a float multiply" scheme, not IEEE RNE)
//   fp32  -> bf16   VS/types/bfloat16.h:22-29   (RNE by adding 0x7FFF + lsb; no NaN special case)
//   normalisers     VS/spaces/normalize/normalize_naive.h:23-88, compute_norm.h:17-28
// Per-vector work at ingest / per query; never on the scan path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace rsb200 {

inline float bits_to_f32(uint32_t u) {
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline uint32_t f32_to_bits(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}

inline float half_bits_to_f32(uint16_t h) {
    const uint32_t exp_mask = 0x7c00u << 13;
    uint32_t mag = ((uint32_t)(h & 0x7fffu)) << 13;
    const uint32_t exp = mag & exp_mask;
    mag += (127u - 15u) << 23;
    uint32_t out;
    if (exp == exp_mask) {
        out = mag + ((128u - 16u) << 23); // inf / nan
    } else if (exp == 0) {
        // zero / subnormal: renormalise through a float subtraction
        out = f32_to_bits(bits_to_f32(mag + (1u << 23)) - bits_to_f32(113u << 23));
    } else {
        out = mag;
    }
    return bits_to_f32(out | (((uint32_t)(h & 0x8000u)) << 16));
}

inline uint16_t f32_to_half_bits(float x) {
    uint32_t u = f32_to_bits(x);
    const uint32_t sign = u & 0x80000000u;
    u ^= sign;
    const uint32_t inf = 255u << 23;
    uint32_t o = (u > inf) ? 0x7e00u : 0x7c00u;
    const uint32_t keep = ~0xfffu;
    float scaled = bits_to_f32(u & keep) * bits_to_f32(15u << 23);
    const float cap = bits_to_f32((31u << 23) - 0x1000u);
    if (cap < scaled) scaled = cap; // std::min(fscale, cap) incl. its NaN behaviour
    const int32_t shifted = (int32_t)f32_to_bits(scaled) - (int32_t)keep;
    if (u < inf) o = (uint32_t)(shifted >> 13);
    return (uint16_t)(o | (sign >> 16));
}

inline float bf16_bits_to_f32(uint16_t b) { return bits_to_f32((uint32_t)b << 16); }
inline uint16_t f32_to_bf16_bits(float x) {
    uint32_t u = f32_to_bits(x);
    u += ((u >> 16) & 1u) + 0x7FFFu;
    return (uint16_t)(u >> 16);
}

// fp32: sum of squares in double, norm rounded to float, float division.
inline void normalize_f32(float *v, size_t dim) {
    double s = 0;
    for (size_t i = 0; i < dim; i++) s += (double)v[i] * (double)v[i];
    const float norm = (float)std::sqrt(s);
    for (size_t i = 0; i < dim; i++) v[i] = v[i] / norm;
}
// fp16 / bf16: sum in float (sequential, unfused), norm = (float)sqrt((double)sum), divide in float,
// round back with the conversions above.
inline void normalize_f16(uint16_t *v, size_t dim) {
    std::vector<float> t(dim);
    volatile float s = 0; // volatile: keep every partial sum rounded to fp32, no contraction
    for (size_t i = 0; i < dim; i++) {
        t[i] = half_bits_to_f32(v[i]);
        const float sq = t[i] * t[i];
        s = s + sq;
    }
    const float norm = (float)std::sqrt((double)s);
    for (size_t i = 0; i < dim; i++) v[i] = f32_to_half_bits(t[i] / norm);
}
inline void normalize_bf16(uint16_t *v, size_t dim) {
    std::vector<float> t(dim);
    volatile float s = 0;
    for (size_t i = 0; i < dim; i++) {
        t[i] = bf16_bits_to_f32(v[i]);
        const float sq = t[i] * t[i];
        s = s + sq;
    }
    const float norm = (float)std::sqrt((double)s);
    for (size_t i = 0; i < dim; i++) v[i] = f32_to_bf16_bits(t[i] / norm);
}
// int8 / uint8: norm = (float)sqrt((double)int_sum) appended after the dim payload bytes.
template <typename T>
inline void append_int_norm(T *v, size_t dim) {
    int s = 0;
    for (size_t i = 0; i < dim; i++) s += (int)v[i] * (int)v[i];
    const float norm = (float)std::sqrt((double)s);
    std::memcpy(reinterpret_cast<uint8_t *>(v) + dim, &norm, sizeof(norm));
}

} // namespace rsb200
