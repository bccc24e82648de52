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
