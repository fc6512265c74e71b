// fsgpu_common.cuh — shared device helpers: IEEE-exact scalar ops, order keys, the generic
// warp-cooperative exact dot, CTA-wide bitonic sort and the bounded top-k candidate buffer.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fsgpu.h"

namespace fsgpu {

typedef ::fsgpu_hit fsgpu_hit_t;
typedef ::fsgpu_fused_hit fsgpu_fused_hit_t;

constexpr uint32_t kNegInfOrdered = 0x007FFFFFu;  // ascending total-order image of -inf

// ─── exact arithmetic (never contracted into FMA) ───────────────────────────────────────────
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }

// Final 8-lane horizontal add in the configured `wide::f32x8::reduce_add` lane order
// (include/fsgpu.h fsgpu_reduce_order; crates/frankensearch-index/src/simd.rs:439).
__device__ __forceinline__ float reduce8(const float v[8], int order) {
    switch (order) {
        case 1:
            return add_rn(add_rn(add_rn(v[0], v[4]), add_rn(v[2], v[6])),
                          add_rn(add_rn(v[1], v[5]), add_rn(v[3], v[7])));
        case 2:
            return add_rn(add_rn(add_rn(add_rn(v[0], v[1]), v[2]), v[3]),
                          add_rn(add_rn(add_rn(v[4], v[5]), v[6]), v[7]));
        case 3:
            return add_rn(add_rn(add_rn(v[0], v[2]), add_rn(v[1], v[3])),
                          add_rn(add_rn(v[4], v[6]), add_rn(v[5], v[7])));
        case 4:
            return add_rn(
                add_rn(add_rn(add_rn(add_rn(add_rn(add_rn(v[0], v[1]), v[2]), v[3]), v[4]), v[5]),
                       v[6]),
                v[7]);
        default:
            return add_rn(add_rn(add_rn(v[0], v[1]), add_rn(v[2], v[3])),
                          add_rn(add_rn(v[4], v[5]), add_rn(v[6], v[7])));
    }
}

__device__ __forceinline__ float h2f(uint16_t bits) {
    return __half2float(__ushort_as_half(bits));  // exact widening (simd.rs:67-81)
}

// ─── order keys (search.rs:1655-1686) ───────────────────────────────
// score_key: NaN -> -inf.  Ascending u32 image of f32::total_cmp, then
// key = (ordered << 32) | ~global_row  so that a LARGER u64 is a BETTER hit (higher score, then
// lower row).  Key 0 is impossible for a real hit (ordered >= 0x007FFFFF) and marks "empty".
__device__ __forceinline__ uint32_t ordered_score(float s) {
    uint32_t u = __float_as_uint(s);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) u = 0xFF800000u;  // NaN -> -inf
    return (u & 0x80000000u) ? ~u : (u ^ 0x80000000u);
}
__device__ __forceinline__ float unordered_score(uint32_t o) {
    const uint32_t u = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t global_row) {
    return ((uint64_t)ordered_score(score) << 32) | (uint32_t)(~global_row);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key) { return ~(uint32_t)key; }
__device__ __forceinline__ float key_score(uint64_t key) {
    return unordered_score((uint32_t)(key >> 32));
}

__device__ __forceinline__ bool tombstoned(const uint8_t* __restrict__ bitmap, uint64_t local_row) {
    return bitmap != nullptr && ((__ldg(bitmap + (local_row >> 3)) >> (local_row & 7)) & 1u);
}

// ─── generic exact dot: one warp per row, any dim ───────────────────────────────────────────
// Lane t = 8a + l owns chain (accumulator a, SIMD lane l) of the reference kernel
// (crates/frankensearch-index/src/simd.rs:418-444): whole groups of four 8-element chunks go to
// accumulators 0..3, left-over chunks to accumulator 0, `(s0+s1)+(s2+s3)`, 8-lane reduce, then
// the scalar tail (`mul_add` when tail_fma, else mul then add).  All 32 lanes must call; every
// lane returns the score.  `q` may be global or shared.
__device__ __forceinline__ float warp_exact_dot(const uint16_t* __restrict__ row,
                                                const float* __restrict__ q, uint32_t dim,
                                                int reduce_order, int tail_fma) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t chunks = dim >> 3;
    const uint32_t groups = chunks >> 2;
    float acc = 0.0f;
    // the adds stay in reference order; unrolling only lets the loads of four groups be in flight
    // together (each is an L2 round trip when the row comes from a gather)
#pragma unroll 4
    for (uint32_t g = 0; g < groups; ++g) {
        const uint32_t e = g * 32u + lane;
        acc = add_rn(acc, mul_rn(h2f(row[e]), q[e]));
    }
    for (uint32_t c = groups * 4u; c < chunks; ++c) {
        if (lane < 8u) {
            const uint32_t e = c * 8u + lane;
            acc = add_rn(acc, mul_rn(h2f(row[e]), q[e]));
        }
    }
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));   // s0+s1 | s2+s3
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));  // (s0+s1)+(s2+s3)
    float v[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) v[l] = __shfl_sync(0xffffffffu, acc, l);
    float result = reduce8(v, reduce_order);
    for (uint32_t e = chunks * 8u; e < dim; ++e) {
        const float val = h2f(row[e]);
        result = tail_fma ? __fmaf_rn(val, q[e], result) : add_rn(result, mul_rn(val, q[e]));
    }
    return result;
}

// ─── exact dot of EIGHT gathered rows per warp (dim % 32 == 0) ──────────────────────────────
// Lane 4r + a owns accumulator a of row r — the lane mapping of scan_topk_fast_kernel — so a warp
// re-scores eight candidate rows with 16-byte loads all in flight at once instead of eight dependent
// one-row gathers (the refine kernel's exact dots were latency-bound: 37 % of its stall samples,
// profiles/r02_refine_hot_lines.txt).  `row` points at THIS lane's row (lanes of a group agree), `q`
// at the query in shared memory (16-byte aligned).  Chunk 4j + a goes to accumulator a in order of j
// starting from 0.0, `(s0+s1)+(s2+s3)` is two xor-shuffles, then the configured 8-lane reduce: bit for
// bit simd.rs:418-439.  Every lane of group r returns row r's score.
__device__ __forceinline__ float warp_exact_dot8(const uint16_t* __restrict__ row, const float* __restrict__ q,
                                                 uint32_t dim, int reduce_order) {
    const uint32_t a = threadIdx.x & 3u;
    const uint4* p = reinterpret_cast<const uint4*>(row) + a;
    const uint32_t nj = dim >> 5;
    float v[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) v[l] = 0.0f;
#pragma unroll 4
    for (uint32_t j = 0; j < nj; ++j) {
        uint4 x;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w)
                     : "l"(p + 4u * j));
        const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&x.x));
        const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&x.y));
        const float2 x2 = __half22float2(*reinterpret_cast<const __half2*>(&x.z));
        const float2 x3 = __half22float2(*reinterpret_cast<const __half2*>(&x.w));
        const float4* qp = reinterpret_cast<const float4*>(q + (4u * j + a) * 8u);
        const float4 q0 = qp[0], q1 = qp[1];
        v[0] = add_rn(v[0], mul_rn(x0.x, q0.x));
        v[1] = add_rn(v[1], mul_rn(x0.y, q0.y));
        v[2] = add_rn(v[2], mul_rn(x1.x, q0.z));
        v[3] = add_rn(v[3], mul_rn(x1.y, q0.w));
        v[4] = add_rn(v[4], mul_rn(x2.x, q1.x));
        v[5] = add_rn(v[5], mul_rn(x2.y, q1.y));
        v[6] = add_rn(v[6], mul_rn(x3.x, q1.z));
        v[7] = add_rn(v[7], mul_rn(x3.y, q1.w));
    }
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        float sacc = v[l];
        sacc = add_rn(sacc, __shfl_xor_sync(0xffffffffu, sacc, 1));  // s0+s1 | s2+s3
        sacc = add_rn(sacc, __shfl_xor_sync(0xffffffffu, sacc, 2));  // (s0+s1)+(s2+s3)
        v[l] = sacc;
    }
    return reduce8(v, reduce_order);
}

// ─── exact f32·f32 dot for resident WAL rows: one warp per row, any dim ─────────────────────
// dot_product_f32_f32 (crates/frankensearch-index/src/simd.rs:161-222 AVX2, :1559-1587 generic):
// four 8-lane accumulators over whole groups of 32 elements, `(acc0+acc1)+(acc2+acc3)` FIRST, then
// the left-over 8-element chunks are added to that combined vector (the f16 kernels add them to
// accumulator 0 before combining), 8-lane reduce, scalar tail `result += a*b` (mul, then add).
__device__ __forceinline__ float warp_exact_dot_f32(const float* __restrict__ row,
                                                    const float* __restrict__ q, uint32_t dim,
                                                    int reduce_order, int tail_fma = 0) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t chunks = dim >> 3;
    const uint32_t groups = chunks >> 2;
    float acc = 0.0f;
#pragma unroll 4
    for (uint32_t g = 0; g < groups; ++g) {
        const uint32_t e = g * 32u + lane;
        acc = add_rn(acc, mul_rn(row[e], q[e]));
    }
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 8));   // acc0+acc1 | acc2+acc3
    acc = add_rn(acc, __shfl_xor_sync(0xffffffffu, acc, 16));  // (acc0+acc1)+(acc2+acc3)
    for (uint32_t c = groups * 4u; c < chunks; ++c) {
        if (lane < 8u) {
            const uint32_t e = c * 8u + lane;
            acc = add_rn(acc, mul_rn(row[e], q[e]));
        }
    }
    float v[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) v[l] = __shfl_sync(0xffffffffu, acc, l);
    float result = reduce8(v, reduce_order);
    // scalar tail: `result += a*b` in the slice kernel (simd.rs:218-221), `mul_add` in the bytes kernel of
    // an f32-quantised FSVI slab (dot_product_f32_bytes_f32, simd.rs:581-760)
    for (uint32_t e = chunks * 8u; e < dim; ++e)
        result = tail_fma ? __fmaf_rn(row[e], q[e], result) : add_rn(result, mul_rn(row[e], q[e]));
    return result;
}

// Exact reference score of one slab row, whichever quantisation the slab has (f16: simd.rs:398-446;
// f32: simd.rs:581-760).  `slab` points at the start of the slab, `local_row` indexes it.
__device__ __forceinline__ float warp_exact_row(const void* __restrict__ slab, int slab_is_f32, uint64_t local_row,
                                                const float* __restrict__ q, uint32_t dim, int reduce_order,
                                                int tail_fma) {
    return slab_is_f32 ? warp_exact_dot_f32(static_cast<const float*>(slab) + local_row * dim, q, dim, reduce_order, 1)
                       : warp_exact_dot(static_cast<const uint16_t*>(slab) + local_row * dim, q, dim, reduce_order, tail_fma);
}

// ─── CTA-wide bitonic sort, descending, n a power of two, keys in shared memory ─────────────
// Every thread owns compare-exchange PAIRS (t -> i with bit j clear, i | j), so no thread idles, and
// the stages with j <= 16 only touch the 64 consecutive keys a warp owns (for every t of that warp,
// in every such stage): they need a warp barrier, not a CTA barrier — 15 CTA barriers instead of 55
// for 1024 keys.  All threads of the CTA must call.
__device__ __forceinline__ void cta_sort_desc(uint64_t* keys, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        if ((k >> 1) > 16u) __syncthreads();  // the warp-local stages of the previous phase are complete
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
                const uint32_t ixj = i | j;
                const uint64_t a = keys[i], b = keys[ixj];
                const bool desc = (i & k) == 0;
                if (desc ? (a < b) : (a > b)) {
                    keys[i] = b;
                    keys[ixj] = a;
                }
            }
            if (j > 16u) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
}

__device__ __forceinline__ uint32_t next_pow2(uint32_t x) {
    return x <= 1 ? 1u : 1u << (32 - __clz(x - 1));
}

// ─── bounded candidate buffer (one per query per CTA) ───────────────────────────────────────
// Threads append keys that beat `tau` (the k-th best key seen so far, 0 until k keys are held);
// `compact` (CTA-collective) sorts, keeps the best k and raises tau.  Equivalent to the
// reference's bounded heap + cutoff (search.rs:1285-1295, :1688-1702): what survives is exactly
// the k largest keys pushed, independent of push order.
struct CandBuf {
    uint64_t* keys;  // [cap] shared
    uint32_t* cnt;   // shared
    uint64_t* tau;   // shared
};

__device__ __forceinline__ bool cand_push(const CandBuf& b, uint32_t cap, uint64_t key) {
    const uint32_t pos = atomicAdd(b.cnt, 1u);
    if (pos < cap) {
        b.keys[pos] = key;
        return true;
    }
    return false;  // capacity contract violated (caller reports)
}

// CTA-collective.  All threads must call with identical arguments.
__device__ __forceinline__ void cand_compact(const CandBuf& b, uint32_t cap, uint32_t k) {
    __syncthreads();
    const uint32_t n = min(*b.cnt, cap);
    const uint32_t n2 = min(next_pow2(n), cap);
    for (uint32_t i = n + threadIdx.x; i < n2; i += blockDim.x) b.keys[i] = 0ull;
    __syncthreads();
    cta_sort_desc(b.keys, n2);
    if (threadIdx.x == 0) {
        const uint32_t kept = min(n, k);
        *b.cnt = kept;
        *b.tau = kept >= k ? b.keys[k - 1] : 0ull;
    }
    __syncthreads();
}

}  // namespace fsgpu
