// synth_kernels.cuh — the reference's bench corpus generators on the device, plus the f32->f16
// RNE encoder used at index build (crates/frankensearch-index/src/simd.rs:2245-2304).
//
// Generator: crates/frankensearch-index/benches/fsvi_int8_two_pass.rs:199-231.  One thread per
// row replays the sequential f32 arithmetic of the reference (xorshift64 -> value, optional
// centroid + noise*value, sequential sum of squares, sqrt, per-element division), so the f16
// slab is bit-identical to the oracle's fso_synth_rows for any row range.
#pragma once

#include "fsgpu_common.cuh"

namespace fsgpu {

__device__ __forceinline__ uint64_t xorshift64(uint64_t s) {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return s;
}
__device__ __forceinline__ float xs_value(uint64_t s) {
    return __fsub_rn(__fdiv_rn((float)(s >> 40), 8388608.0f), 1.0f);
}

// centroid c = normalize(raw_vector(0xc000_0000 + c)); one thread per centroid.
__global__ void synth_centroids_kernel(uint32_t n_centroids, uint32_t dim, float* __restrict__ out) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_centroids) return;
    float* v = out + (size_t)c * dim;
    uint64_t s = (0xc0000000ull + c) | 1ull;
    float acc = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) {
        s = xorshift64(s);
        const float x = xs_value(s);
        v[d] = x;
        acc = add_rn(acc, mul_rn(x, x));
    }
    const float norm = __fsqrt_rn(acc);
    if (norm > 1e-12f)
        for (uint32_t d = 0; d < dim; ++d) v[d] = __fdiv_rn(v[d], norm);
}

// Two passes over the generator per row (norm, then emit) keep the thread state in registers.
__global__ void __launch_bounds__(128)
synth_rows_kernel(int kind, uint64_t seed_base, uint64_t row_start, uint64_t n_rows, uint32_t dim,
                  uint32_t n_centroids, float noise, const float* __restrict__ centroids,
                  uint16_t* __restrict__ out) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const uint64_t i = row_start + r;
    const float* cen = kind == 1 ? centroids + (size_t)(i % n_centroids) * dim : nullptr;
    const uint64_t seed = (seed_base + i) | 1ull;
    uint64_t s = seed;
    float acc = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) {
        s = xorshift64(s);
        float x = xs_value(s);
        if (cen) x = add_rn(cen[d], mul_rn(noise, x));
        acc = add_rn(acc, mul_rn(x, x));
    }
    const float norm = __fsqrt_rn(acc);
    const bool scale = norm > 1e-12f;
    s = seed;
    uint16_t* o = out + r * dim;
    for (uint32_t d = 0; d < dim; ++d) {
        s = xorshift64(s);
        float x = xs_value(s);
        if (cen) x = add_rn(cen[d], mul_rn(noise, x));
        if (scale) x = __fdiv_rn(x, norm);
        o[d] = __half_as_ushort(__float2half_rn(x));
    }
}

// encode_f32_to_f16_extend (simd.rs:2245-2304): IEEE round-to-nearest-even.
__global__ void encode_f16_kernel(const float* __restrict__ src, uint64_t n,
                                  uint16_t* __restrict__ dst) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        dst[i] = __half_as_ushort(__float2half_rn(src[i]));
}

}  // namespace fsgpu
