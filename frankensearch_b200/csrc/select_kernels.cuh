// select_kernels.cuh — exact top-k for LARGE k (128 < k <= 4096) of one query: grid-wide radix select
// instead of per-CTA candidate buffers.
//
// The reference handles any `limit` with one bounded heap per 1024-row chunk and a serial merge
// (crates/frankensearch-index/src/search.rs:1013-1036, :1257-1327, :1704-1720); its result is "sort
// all live rows by (score_key desc, row asc), take the first `limit`" (search.rs:1655-1686).  The
// per-CTA buffers of scan_kernels.cuh reproduce that well for small k, but at k = 3000 (the fetch of
// a top-1000 search, sync_searcher.rs:654) every CTA keeps 3000 entries and one CTA merges
// grid x 3000 keys: 3.5 ms per query at 10 M rows against 1.3 ms at k = 10
// (profiles/r01_sweep_k_final.txt).  Here the pass over the corpus only WRITES one score per row
// (4 bytes against the 384-768 it reads) and the selection is a radix select over that array, which
// lives in the 126 MB L2:
//
//   int8 index:  scan_i8_all_kernel  approx score of every row from the int8 codes (HBM: n*dim bytes)
//                sel_hist x3         k-th largest approx score T            (n*4 bytes from L2 per pass)
//                sel_compact         rows with approx >= T - 2e (|approx - exact| <= e: a superset
//                                    of the exact top-k, same proof as mma_scan_kernels.cuh)
//                gather_list_keys    exact reference score of each listed row -> order keys
//                sel_hist x6         k-th largest KEY (64 bits: score, then lower row)
//                sel_compact + sel_emit   exactly k keys, sorted best first
//   f16 index:   score_all_kernel (exact key per row) -> sel_hist x6 -> sel_compact -> sel_emit
//
// Keys are distinct (distinct rows), so the 64-bit select returns exactly min(k, live rows) keys for
// any tie structure; nothing can overflow (the position lists are sized for every row).
#pragma once

#include "fsgpu_common.cuh"

namespace fsgpu {

constexpr uint32_t kSelBins = 2048;
constexpr int kSelMaxPasses = 6;
constexpr uint32_t kSelMaxK = 4096;  // sel_emit_kernel sorts the winners in shared memory

struct SelState {
    uint32_t hist[kSelMaxPasses][kSelBins];
    unsigned long long prefix;  // the bits decided so far, in place
    uint32_t k_rem;             // rank still to find among the values that carry `prefix`
    uint32_t all;               // 1: fewer than k live values exist — every live value is selected
    uint32_t ticket[kSelMaxPasses];
    uint32_t n_out;             // entries the compaction listed
    uint32_t pad;
};

template <class Key>
struct SelTraits;
template <>
struct SelTraits<uint32_t> {
    static constexpr int kPasses = 3;
    __host__ __device__ static int shift(int p) { return p == 0 ? 21 : p == 1 ? 10 : 0; }
    __host__ __device__ static int width(int p) { return p == 2 ? 10 : 11; }
};
template <>
struct SelTraits<unsigned long long> {
    static constexpr int kPasses = 6;
    __host__ __device__ static int shift(int p) { return p == 5 ? 0 : 53 - 11 * p; }
    __host__ __device__ static int width(int p) { return p == 5 ? 9 : 11; }
};

// One pass of the select: histogram of the next digit over the values that carry the prefix found so
// far (value 0 = excluded row); the LAST CTA to finish picks the digit that holds the k-th largest
// and publishes prefix / k_rem for the next pass.  `n_ptr` (nullable): the live length of `vals` when
// only the device knows it.
template <class Key>
__global__ void __launch_bounds__(256)
sel_hist_kernel(const Key* __restrict__ vals, uint64_t n, const uint32_t* __restrict__ n_ptr, SelState* st, int pass,
                uint32_t k) {
    using T = SelTraits<Key>;
    __shared__ uint32_t h[kSelBins];
    __shared__ uint32_t scan[256];
    __shared__ int s_last;
    if (n_ptr) n = min(n, (uint64_t)*n_ptr);
    const int shift = T::shift(pass), width = T::width(pass);
    const uint32_t mask = (1u << width) - 1u;
    unsigned long long prefix = 0ull;
    uint32_t all = 0u;
    if (pass > 0) {
        prefix = st->prefix;
        all = st->all;
    }
    for (uint32_t i = threadIdx.x; i < kSelBins; i += blockDim.x) h[i] = 0u;
    __syncthreads();
    if (!all) {
        const int prev_shift = pass > 0 ? T::shift(pass - 1) : 0;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
            const Key v = vals[i];
            if (v != 0 && (pass == 0 || ((unsigned long long)v >> prev_shift) == (prefix >> prev_shift)))
                atomicAdd(&h[(uint32_t)((unsigned long long)v >> shift) & mask], 1u);
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < kSelBins; i += blockDim.x)
        if (h[i]) atomicAdd(&st->hist[pass][i], h[i]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&st->ticket[pass], 1u) == gridDim.x - 1u;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // pick: walk the digits from the top until the running count reaches k_rem; thread t owns digits
    // [8t, 8t+8)
    const uint32_t k_rem = pass == 0 ? k : st->k_rem;
    uint32_t local[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        local[i] = __ldcg(&st->hist[pass][threadIdx.x * 8 + i]);
        sum += local[i];
    }
    scan[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {  // suffix sums over 256 groups (serial: 256 adds, once per pass)
        uint32_t run = 0;
        for (int t = 255; t >= 0; --t) {
            const uint32_t v = scan[t];
            scan[t] = run;  // entries in groups above t
            run += v;
        }
        h[0] = run;  // total
    }
    __syncthreads();
    const uint32_t total = h[0];
    if (pass == 0 && total < k_rem) {  // fewer than k live values: everything live is selected
        if (threadIdx.x == 0) {
            st->all = 1u;
            st->prefix = 0ull;
            st->k_rem = 0u;
        }
        return;
    }
    if (all) return;
    const uint32_t above = scan[threadIdx.x];
    if (above < k_rem && above + sum >= k_rem) {  // exactly one thread
        uint32_t acc = above;
        for (int i = 7; i >= 0; --i) {
            if (acc + local[i] >= k_rem) {
                st->prefix = prefix | ((unsigned long long)(threadIdx.x * 8 + i) << shift);
                st->k_rem = k_rem - acc;
                if (pass == 0) st->all = 0u;
                break;
            }
            acc += local[i];
        }
    }
}

// Lists the positions whose value reaches the selected threshold (after the last pass `prefix` IS the
// k-th largest value).  u32 scores may lower the threshold by `margin2` first (the candidate
// superset of an approximate score).  The callers size `positions` so that nothing can be dropped
// (every row for the approximate stage, kSelMaxK for the exact one); `cap` only guards the buffer.
template <class Key>
__global__ void __launch_bounds__(256)
sel_compact_kernel(const Key* __restrict__ vals, uint64_t n, const uint32_t* __restrict__ n_ptr,
                   const SelState* __restrict__ st, const float* __restrict__ margin2,
                   const uint32_t* __restrict__ take_all, uint32_t* __restrict__ positions, uint32_t cap,
                   uint32_t* __restrict__ n_out) {
    if (n_ptr) n = min(n, (uint64_t)*n_ptr);
    Key thr = (Key)st->prefix;
    // `take_all`: the approximate scores carry no usable bound for this query (non-finite query, bound
    // overflow): every live row goes on to the exact stage
    const bool everything = st->all != 0u || (take_all && *take_all != 0u);
    if (everything) thr = 1;
    if constexpr (sizeof(Key) == 4) {
        if (margin2 && !everything) thr = ordered_score(__fsub_rd(unordered_score((uint32_t)thr), *margin2));
    }
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t step = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x; i0 < n; i0 += step) {  // warp-uniform trip count
        const uint64_t i = i0 + threadIdx.x;
        const Key v = i < n ? vals[i] : (Key)0;
        const bool take = v != 0 && v >= thr;
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (m == 0u) continue;
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(n_out, __popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
        if (take && slot < cap) positions[slot] = (uint32_t)i;
    }
}

// Approximate score of EVERY row from the int8 codes, as an ascending u32 (0 = excluded row): the
// single-query dp4a pass of scan_kernels.cuh without its per-CTA top-k.
__global__ void __launch_bounds__(256)
scan_i8_all_kernel(const int8_t* __restrict__ codes, const uint8_t* __restrict__ tombstones,
                   const int8_t* __restrict__ q_codes, const float* __restrict__ qscale_ptr, uint64_t n_rows,
                   uint32_t dim, uint32_t* __restrict__ approx) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, rr = lane >> 3;
    const uint32_t nj = dim >> 7;  // 128-byte segments per row (dim % 128 == 0, dim <= 512)
    int y[4][4];
#pragma unroll
    for (uint32_t j = 0; j < 4; ++j) {
        if (j < nj) {
            const int4 v = *reinterpret_cast<const int4*>(q_codes + j * 128u + sub * 16u);
            y[j][0] = v.x; y[j][1] = v.y; y[j][2] = v.z; y[j][3] = v.w;
        } else {
            y[j][0] = y[j][1] = y[j][2] = y[j][3] = 0;
        }
    }
    const float qscale = *qscale_ptr;
    constexpr uint32_t kRows = 8 * 4 * 2;  // rows per CTA iteration: 8 warps x 4 rows per load x 2 loads in flight
    const uint64_t n_tiles = (n_rows + kRows - 1) / kRows;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t row0 = tile * kRows + (uint64_t)warp * 8u + rr;
        uint4 x[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t row = row0 + 4u * h;
            const uint64_t rowc = row < n_rows ? row : n_rows - 1;
            const uint4* p = reinterpret_cast<const uint4*>(codes + rowc * dim + sub * 16u);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (j < nj) {
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(x[h][j].x), "=r"(x[h][j].y), "=r"(x[h][j].z), "=r"(x[h][j].w)
                                 : "l"(p + j * 8u));
                }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t row = row0 + 4u * h;
            int acc = 0;
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (j < nj) {
                    acc = __dp4a((int)x[h][j].x, y[j][0], acc);
                    acc = __dp4a((int)x[h][j].y, y[j][1], acc);
                    acc = __dp4a((int)x[h][j].z, y[j][2], acc);
                    acc = __dp4a((int)x[h][j].w, y[j][3], acc);
                }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (sub == 0 && row < n_rows)
                approx[row] = tombstoned(tombstones, row) ? 0u : ordered_score(__fmul_rn((float)acc, qscale));
        }
    }
}

// Exact order key of every listed LOCAL row (reference arithmetic), one warp per row, grid-stride over
// a list whose length only the device knows.
__global__ void __launch_bounds__(256)
gather_list_keys_kernel(const uint16_t* __restrict__ slab, uint64_t row_base, uint32_t dim,
                        const float* __restrict__ query, const uint32_t* __restrict__ positions,
                        const uint32_t* __restrict__ n_positions, int reduce_order, int tail_fma,
                        unsigned long long* __restrict__ out_keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* q = reinterpret_cast<float*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) q[i] = query[i];
    __syncthreads();
    const uint32_t n = *n_positions;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += n_warps) {
        const uint32_t row = positions[i];
        const float s = warp_exact_dot(slab + (size_t)row * dim, q, dim, reduce_order, tail_fma);
        if (lane == 0) out_keys[i] = make_key(s, (uint32_t)(row_base + row));
    }
}

// The selected keys (exactly min(k, live rows) of them), sorted best first, as the result of one
// query.  Raw scores of the -inf/NaN class are recomputed (VectorHit carries the RAW score,
// search.rs:1549-1553).  One CTA of 1024 threads; dynamic shared memory: kSelMaxK keys.
__global__ void __launch_bounds__(1024)
sel_emit_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ positions,
                const uint32_t* __restrict__ n_sel, uint32_t k, const void* __restrict__ slab, int slab_is_f32,
                const float* __restrict__ query, uint64_t n_rows, uint64_t row_base, uint32_t dim, int reduce_order,
                int tail_fma, uint64_t* __restrict__ out_keys, fsgpu_hit_t* __restrict__ out_hits,
                uint32_t* __restrict__ out_count, uint32_t* __restrict__ error_flag) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t n = *n_sel;
    if (n > kSelMaxK) {  // cannot happen: keys are distinct (see header); reported, never silently wrong
        if (threadIdx.x == 0) atomicExch(error_flag, 1u);
        n = kSelMaxK;
    }
    const uint32_t n2 = next_pow2(max(n, 1u));
    for (uint32_t i = threadIdx.x; i < n2; i += blockDim.x) buf[i] = i < n ? keys[positions[i]] : 0ull;
    __syncthreads();
    cta_sort_desc(buf, n2);
    const uint32_t count = min(n, k);
    if (threadIdx.x == 0 && out_count) *out_count = count;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (uint32_t i = warp; i < k; i += n_warps) {  // warp-uniform
        const uint64_t key = i < count ? buf[i] : 0ull;
        float score = key ? key_score(key) : 0.0f;
        if (key && (uint32_t)(key >> 32) == kNegInfOrdered) {
            const uint64_t grow = key_row(key);
            if (grow >= row_base && grow - row_base < n_rows)
                score = warp_exact_row(slab, slab_is_f32, grow - row_base, query, dim, reduce_order, tail_fma);
        }
        if (lane == 0) {
            if (out_keys) out_keys[i] = key;
            if (out_hits) {
                fsgpu_hit_t h;
                h.row = key ? key_row(key) : 0xFFFFFFFFu;
                h.score = score;
                out_hits[i] = h;
            }
        }
    }
}

}  // namespace fsgpu
