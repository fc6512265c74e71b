// two_pass_kernels.cuh — the reference's quantised two-pass searches with THEIR semantics:
// VectorIndex::search_top_k_int8_two_pass / search_top_k_4bit_two_pass
// (crates/frankensearch-index/src/search.rs:514-650, :876-946): pass 1 ranks every live row by an INTEGER dot of
// corpus-wide-scaled codes (int8: simd.rs:1842-1859; signed 4-bit nibbles, two dims per byte: simd.rs:2201-2233)
// against the query's own codes (search.rs:1610-1655), keeps the best k * candidate_multiplier by
// (score, lower row), and pass 2 re-scores exactly those rows with the f16 kernel and keeps the top k.
// Unlike the library's own int8 forms (exact by a proven bound) this is the reference's
// recall-vs-multiplier trade: a row the integer ranking places outside the candidate set is lost, exactly as it is
// there.  Every quantity of pass 1 is an exact integer, so the candidate set — and with it the result — is
// bit-identical to the reference's for any multiplier (tests/test_gpu_two_pass.py against the oracle restatement).
// Pass 1 writes one 64-bit order key per row; the grid-wide radix select of select_kernels.cuh does the rest.
#pragma once

#include "fsgpu_common.cuh"

namespace fsgpu {

// nibble_of_4bit (simd.rs:1892-1896): clamp(round_half_away(x * scale), -7, 7) as 4-bit two's complement
__device__ __forceinline__ uint32_t nibble_of(float x, float scale) {
    const float c = fminf(fmaxf(roundf(__fmul_rn(x, scale)), -7.0f), 7.0f);
    return (uint32_t)(int)c & 0xFu;
}

// pack_f16_slab_to_4bit_generic (simd.rs:2201-2233): byte j of a row = dims 2j (low nibble) and 2j + 1 (high)
__global__ void __launch_bounds__(256)
pack_slab_4bit_kernel(const uint16_t* __restrict__ slab, uint64_t n_rows, uint32_t dim, float scale,
                      uint8_t* __restrict__ out) {
    const uint32_t bpv = (dim + 1u) / 2u;
    const uint64_t total = n_rows * bpv;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = i / bpv;
        const uint32_t d = (uint32_t)(i % bpv) * 2u;
        uint32_t b = nibble_of(h2f(slab[row * dim + d]), scale);
        if (d + 1u < dim) b |= nibble_of(h2f(slab[row * dim + d + 1u]), scale) << 4;
        out[i] = (uint8_t)b;
    }
}

// quantize_f16_slab_to_i8_generic (simd.rs:1842-1859) for any dim (the index's resident codes need dim % 128 == 0)
__global__ void __launch_bounds__(256)
quantize_slab_i8_any_kernel(const uint16_t* __restrict__ slab, uint64_t n_elems, float scale, int8_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_elems; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = (int8_t)(int)fminf(fmaxf(roundf(__fmul_rn(h2f(slab[i]), scale)), -127.0f), 127.0f);
}

// Pass 1: key[row] = (ordered(f32(integer dot)), ~global row), 0 for tombstoned rows.  Eight lanes per row, 16-byte
// loads when the rows allow it.  BITS = 8: q_a = the query's int8 codes.  BITS = 4: q_a[j] / q_b[j] = the
// sign-extended nibbles of dims 2j / 2j + 1; a stored word w gives lo = (w << 4) & 0xF0F0F0F0 (low nibbles x 16 as
// int8) and hi = w & 0xF0F0F0F0, so dp4a(lo, q_a) + dp4a(hi, q_b) = 16 x the nibble dot (dot_4bit_prepared,
// simd.rs:1347-1367: exact, per-dim products <= 49).
template <int BITS>
__global__ void __launch_bounds__(256)
two_pass_scan_kernel(const uint8_t* __restrict__ codes, uint32_t row_bytes, const uint8_t* __restrict__ tombstones,
                     const int8_t* __restrict__ q_a, const int8_t* __restrict__ q_b, uint64_t n_rows, uint64_t row_base,
                     unsigned long long* __restrict__ keys) {
    extern __shared__ __align__(16) unsigned char tp_smem[];
    int8_t* qa = reinterpret_cast<int8_t*>(tp_smem);
    int8_t* qb = qa + ((row_bytes + 15u) & ~15u);
    for (uint32_t i = threadIdx.x; i < row_bytes; i += blockDim.x) {
        qa[i] = q_a[i];
        if (BITS == 4) qb[i] = q_b[i];
    }
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, sub = lane & 7u, grp = lane >> 3;
    const bool vec = (row_bytes & 15u) == 0u;
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t base = warp_global * 4u; base < n_rows; base += n_warps * 4u) {  // warp-uniform: four rows per warp step
        const uint64_t row0 = base + grp;
        const bool live_row = row0 < n_rows;
        const uint8_t* p = codes + (live_row ? row0 : 0) * row_bytes;
        int acc = 0;
        if (live_row) {
            if (vec) {
                for (uint32_t c = sub; c < (row_bytes >> 4); c += 8u) {
                    const uint4 w = *reinterpret_cast<const uint4*>(p + c * 16u);
                    const int4 a = *reinterpret_cast<const int4*>(qa + c * 16u);
                    if (BITS == 8) {
                        acc = __dp4a((int)w.x, a.x, acc);
                        acc = __dp4a((int)w.y, a.y, acc);
                        acc = __dp4a((int)w.z, a.z, acc);
                        acc = __dp4a((int)w.w, a.w, acc);
                    } else {
                        const int4 b = *reinterpret_cast<const int4*>(qb + c * 16u);
                        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
                        const int as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            acc = __dp4a((int)((ws[t] << 4) & 0xF0F0F0F0u), as[t], acc);
                            acc = __dp4a((int)(ws[t] & 0xF0F0F0F0u), bs[t], acc);
                        }
                    }
                }
            } else {
                for (uint32_t j = sub; j < row_bytes; j += 8u) {
                    const uint32_t s = p[j];
                    if (BITS == 8) {
                        acc += (int)(int8_t)s * (int)qa[j];
                    } else {
                        acc += ((int)(int8_t)(s << 4)) * (int)qa[j] + ((int)(int8_t)(s & 0xF0u)) * (int)qb[j];
                    }
                }
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        if (sub == 0 && live_row) {
            if (BITS == 4) acc >>= 4;  // every term carried a factor 16
            keys[row0] = tombstoned(tombstones, row0) ? 0ull : make_key((float)acc, (uint32_t)(row_base + row0));
        }
    }
}

}  // namespace fsgpu
