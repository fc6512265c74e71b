// mma_scan_kernels.cuh — the batched (>= 3 queries) form of the exact f16 scan + top-k.
//
// A batch of B queries against the slab is a dense contraction [B, D] x [D, N] (SURVEY.md §0
// F5): on CUDA cores it is compute-bound at ~4 queries per corpus pass.  This path runs the
// contraction on the 5th-gen tensor cores and keeps the result EXACT by construction:
//
//   1. prep:    q_hat = f16(q) (what the tensor core can take), and a rigorous per-query bound
//               e_q >= |mma_score(row) - reference_score(row)| for every row of the index
//               (Cauchy-Schwarz on the rounding residual + accumulation slack, see prep kernel).
//   2. scan:    persistent warp-specialised kernel.  One CTA owns 128 queries (UMMA M = 128,
//               resident in shared memory) and streams 128-row slab tiles (UMMA N = 128) through
//               a TMA -> mbarrier -> tcgen05.mma -> TMEM pipeline.  TMEM lane = query, column =
//               corpus row, so each epilogue thread owns one query: it holds that query's gate
//               in a register and appends every row whose approximate score clears it to the
//               query's candidate list (a 3-input-max tree rejects 8 rows per 4 instructions).
//               Gates come from a sample cascade: level 0 keeps every score of ~16 strided
//               tiles, level 1 scans ~sqrt(16 * n_tiles) strided tiles against the k'-th best
//               of level 0, level 2 scans everything against the k'-th best of level 1 minus
//               2 e_q.  The k'-th best (k' = max(k, 8)) of any subset is a lower bound of the
//               corpus' k-th best, so every row of the exact top-k clears the final gate
//               (proof in DESIGN.md "Batched scan") and the final list is a superset of it.
//   3. refine:  one CTA per query re-scores the band [tau_approx - 2 e_q, inf) of its list with
//               the reference's exact accumulation tree (warp_exact_dot, simd.rs:398-446) and
//               selects the top-k with the reference's total order (search.rs:1655-1686).
//
// The slab is read from HBM once per launch for up to 148*128 queries; the other query blocks
// hit the same tiles in L2.  Queries the bound cannot cover (non-finite / f16-overflowing
// components, or a candidate list that overflowed, e.g. a huge tie band) are flagged and re-run
// by the caller on the exact CUDA-core kernel (scan_kernels.cuh) — still on the GPU.
#pragma once

#include "fsgpu_common.cuh"
#include "tc_ptx.cuh"

namespace fsgpu {

constexpr int kMmaThreads = 320;    // warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: epilogue
constexpr int kMmaEpiWarps = 8;     // two warps per TMEM lane quarter, each takes half of the columns
constexpr int kMmaM = 128;          // queries per CTA (UMMA M, cta_group::1)
constexpr int kMmaN = 128;          // corpus rows per tile (UMMA N)
constexpr int kMmaAccStages = 4;    // 4 x 128 TMEM columns = the whole 512-column TMEM
constexpr int kMmaMaxStages = 8;
constexpr uint32_t kMmaMaxK = 1024;  // larger k goes to the exact path (score-all + radix sort)
constexpr uint32_t kMmaMaxDim = 512;

// One candidate: the approximate (tensor-core) score and the GLOBAL row.
struct __align__(8) MmaCand {
    float score;
    uint32_t row;
};

struct MmaScanArgs {
    uint64_t n_rows, row_base;
    const uint8_t* tombstones;
    uint32_t n_kblocks;        // dim / 64
    uint32_t n_qblocks;        // ceil(batch / 128)
    uint32_t ctas_per_qblock;  // gridDim.x = n_qblocks * ctas_per_qblock
    uint32_t batch;
    uint32_t n_stages;         // B ring depth
    // the tiles of this level: tile_of(i) = i*stride + jitter(i) % stride, i in [0, count): one
    // pseudo-random tile per stratum (a constant stride camps on a few HBM channels: a strided
    // sample pass ran at 140 GB/s, profiles/r01_mma_v3_L1_strided_ncu.json)
    uint64_t tile_stride, tile_count;
    // full pass of the quad kernel after a sample level whose lists are still in place: the tiles that level
    // scanned (tile_of(i) for i < skip_count under stride skip_stride) are not scanned again — their rows above
    // this pass's gate are already in the lists (`carry`: every thread first compacts its list against its gate
    // and appends behind it).  Sound because every level's gate is a lower bound of the next one's.
    uint32_t skip_stride, skip_count, carry;
    uint32_t dump_group_max;   // first sample level: append only the best score of every 8-row group
    const float* gate;         // [n_qblocks*128] static per-query gate (nullptr = -inf: keep everything)
    const float* qscale;       // int8 form only: [n_qblocks*128] score = accumulator * qscale[query]
    const uint32_t* redo;      // [n_qblocks*128] != 0: query is served by the exact path, skip it
    // pacing (pair form, several query pairs per tile stream): progress[stream][query pair] = tiles
    // whose loads have been issued; a pair never runs more than `lead` tiles ahead of the slowest
    // pair that reads the same tiles, so those reads stay inside the L2 window.  nullptr = off.
    uint32_t* progress;
    uint32_t lead;
    MmaCand* cand;             // [gridDim.x][2][128][cap]: one private list per epilogue thread
    uint32_t* cand_count;      // [gridDim.x][2][128] appended entries (may exceed cap = overflow)
    uint32_t cap;
    long long* ts;             // FSGPU_MMA_TS: clock64 stamps of CTA 0's issuer / epilogue warps, tiles 64..127 of its stream
    uint32_t dbg;              // timing experiments only (FSGPU_MMA_DBG; results are wrong): 1 = epilogue skips its
                               // work, 2 = hot bits are computed but nothing is appended (quad kernel)
};

// ─── prep: q -> q_hat (f16), error bound, safety flags ──────────────────────────────────────
// One CTA per (padded) query slot.  For slot b < batch:
//   q_hat_i = f16_rn(q_i);  r = ||q - q_hat||_2;  nq = ||q||_2   (accumulated in f64)
//   e = max_row_norm * ( r + (D*2^-22 + (D/32+8)*2^-23) * (nq + r) ) * 1.01
// where, for every row x of the index (||x|| <= max_row_norm, all finite):
//   |sum x_i (q_i - q_hat_i)|            <= ||x|| r                      (Cauchy-Schwarz)
//   |tensor-core f32 accumulation error| <= D * 2^-22 * sum|x_i q_hat_i| (one truncation per add,
//                                           2x slack) <= D*2^-22 ||x|| (nq + r)
//   |reference rounding error|           <= (D/32+8) * 2^-23 * sum|x_i q_i|  (mul + D/32 chain adds
//                                           + 5 tree adds, round-to-nearest, 2x slack)
// margin2 = 2e rounded up.  Queries with a non-finite component or |q_i| > 65504 cannot be
// represented: they get q_hat = 0 and redo = 1.  Padding slots get redo = 0, q_hat = 0.
__global__ void __launch_bounds__(128)
mma_prep_queries_kernel(const float* __restrict__ queries, uint32_t batch, uint32_t dim, float max_row_norm,
                        __half* __restrict__ q_hat, float* __restrict__ margin2, uint32_t* __restrict__ redo) {
    const uint32_t b = blockIdx.x;
    __shared__ double s_r2[4], s_n2[4];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    double r2 = 0.0, n2 = 0.0;
    bool bad = false;
    if (b < batch) {
        for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) {
            const float q = queries[(size_t)b * dim + i];
            if (!(fabsf(q) <= 65504.0f)) bad = true;  // NaN, inf or f16 overflow
            const __half h = __float2half_rn(q);
            const double d = (double)q - (double)__half2float(h);
            r2 += d * d;
            n2 += (double)q * (double)q;
        }
    }
    if (bad) atomicOr(&s_bad, 1);
    for (int o = 16; o > 0; o >>= 1) {
        r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_r2[threadIdx.x >> 5] = r2;
        s_n2[threadIdx.x >> 5] = n2;
    }
    __syncthreads();
    const bool is_bad = s_bad != 0;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) {
        float q = 0.0f;
        if (b < batch && !is_bad) q = queries[(size_t)b * dim + i];
        q_hat[(size_t)b * dim + i] = __float2half_rn(q);
    }
    if (threadIdx.x == 0) {
        const double r = sqrt(s_r2[0] + s_r2[1] + s_r2[2] + s_r2[3]);
        const double nq = sqrt(s_n2[0] + s_n2[1] + s_n2[2] + s_n2[3]);
        const double d = (double)dim;
        const double slack = d * (1.0 / 4194304.0) + (d / 32.0 + 8.0) * (1.0 / 8388608.0);
        const double e = (double)max_row_norm * (r + slack * (nq + r)) * 1.01;
        float m2 = __double2float_ru(2.0 * e);
        if (!(m2 >= 0.0f) || is_bad) m2 = 0.0f;
        margin2[b] = m2;
        redo[b] = (b < batch && is_bad) ? 1u : 0u;
    }
}

// ─── index statistics for the bound: max row norm, all-finite flag ──────────────────────────
// stats[0] = bits of max ||row||_2 (f32, rounded up generously by the caller), stats[1] = 1 if
// any element is inf/NaN, stats[2] = f16 magnitude bits of the largest |element|.
__global__ void __launch_bounds__(256)
slab_stats_kernel(const uint16_t* __restrict__ slab, uint64_t n_rows, uint32_t dim, uint32_t* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    float local_max = 0.0f;
    uint32_t abs_bits = 0u;  // largest |element| as f16 magnitude bits (finite magnitudes order as integers)
    bool nonfinite = false;
    for (uint64_t row = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += (uint64_t)gridDim.x * 8) {
        const uint16_t* p = slab + row * dim;
        float acc = 0.0f;
        for (uint32_t i = lane; i < dim; i += 32) {
            const uint16_t bits = p[i];
            if ((bits & 0x7C00u) == 0x7C00u) nonfinite = true;
            abs_bits = max(abs_bits, (uint32_t)(bits & 0x7FFFu));
            const float x = h2f(bits);
            acc = __fmaf_ru(x, x, acc);
        }
        for (int o = 16; o > 0; o >>= 1) acc = __fadd_ru(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        local_max = fmaxf(local_max, acc);
    }
    if (__any_sync(0xffffffffu, nonfinite) && lane == 0) atomicOr(stats + 1, 1u);
    abs_bits = __reduce_max_sync(0xffffffffu, abs_bits);
    if (lane == 0 && abs_bits) atomicMax(stats + 2, abs_bits);
    if (lane == 0 && local_max > 0.0f && local_max == local_max)
        atomicMax(stats, __float_as_uint(__fsqrt_ru(local_max)));  // non-negative floats order as uints
}

// ─── int8 form: corpus codes, query codes, error bound ──────────────────────────────────────
// The reference's corpus-wide int8 quantiser (quantize_f16_slab_to_i8,
// crates/frankensearch-index/src/simd.rs:1842-1859): scale = 127 / max|x|, code =
// clamp(round_half_away(x * scale), -127, 127).  Here the codes feed tcgen05.mma kind::i8 (half the
// HBM / L2 / shared-memory bytes and twice the MMA rate of the f16 form) and the result stays EXACT:
// with x_i = sx*X_i + ex_i (sx = max|x| / 127 as used below, X the codes actually stored) and
// y_i = sy*Y_i + ey_i,
//     x.y = sx*sy * sum X_i Y_i  +  sum (sx X_i) ey_i  +  sum ex_i y_i
//     |x.y - sx*sy*acc| <= (||x|| + ||ex||) ||ey|| + ||ex|| ||y||          (Cauchy-Schwarz, twice)
// so with Ex = max over rows of ||ex_row|| (measured from the stored codes by this kernel, rounded
// up) every row's approximate score is within e_q of its reference score; the candidate superset,
// the exact re-score and the redo rules are those of the f16 form.
// One warp per row; stats[0] = bits of max ||ex_row||_2 (f32, rounded up).
__global__ void __launch_bounds__(256)
quantize_slab_i8_kernel(const uint16_t* __restrict__ slab, uint64_t n_rows, uint32_t dim, float scale, float sx,
                        int8_t* __restrict__ out, uint32_t* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    float local_max = 0.0f;
    for (uint64_t row = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5); row < n_rows; row += (uint64_t)gridDim.x * 8) {
        const uint16_t* p = slab + row * dim;
        int8_t* o = out + row * dim;
        float acc = 0.0f;
        for (uint32_t i = lane * 4u; i < dim; i += 128u) {  // dim % 128 == 0: four codes per lane per step
            const uint2 raw = *reinterpret_cast<const uint2*>(p + i);
            const uint16_t h[4] = {(uint16_t)(raw.x & 0xFFFFu), (uint16_t)(raw.x >> 16), (uint16_t)(raw.y & 0xFFFFu),
                                   (uint16_t)(raw.y >> 16)};
            uint32_t packed = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float x = h2f(h[j]);
                const float c = fminf(fmaxf(roundf(__fmul_rn(x, scale)), -127.0f), 127.0f);
                const float err = __fmaf_rn(-sx, c, x);  // one rounding; its 2^-24 relative error is in the 1.01
                acc = __fmaf_ru(err, err, acc);
                packed |= ((uint32_t)(uint8_t)(int8_t)(int)c) << (8 * j);
            }
            *reinterpret_cast<uint32_t*>(o + i) = packed;
        }
        for (int s2 = 16; s2 > 0; s2 >>= 1) acc = __fadd_ru(acc, __shfl_xor_sync(0xffffffffu, acc, s2));
        local_max = fmaxf(local_max, acc);
    }
    if (lane == 0 && local_max > 0.0f) atomicMax(stats, __float_as_uint(__fsqrt_ru(local_max)));
}

// One CTA per (padded) query slot: codes Y (int8), qscale = sx*sy, margin2 = 2 e_q rounded up, with
//   e_q = [ (R + Ex) ey + Ex nq + (D/32+8) 2^-23 R nq + 2^-21 (R + Ex)(nq + ey) ] * 1.01
// (R = max row norm, nq = ||y||, ey = ||y - sy Y||; third term: the reference's own f32 rounding,
// as in the f16 form; fourth: the roundings of qscale and of acc * qscale, 8x headroom).
__global__ void __launch_bounds__(128)
mma_prep_queries_i8_kernel(const float* __restrict__ queries, uint32_t batch, uint32_t dim, float max_row_norm,
                           float max_ex, float sx, int8_t* __restrict__ q_hat, float* __restrict__ margin2,
                           float* __restrict__ qscale, uint32_t* __restrict__ redo) {
    const uint32_t b = blockIdx.x;
    __shared__ double s_e2[4], s_n2[4];
    __shared__ float s_max[4];
    __shared__ int s_bad;
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    float amax = 0.0f;
    bool bad = false;
    if (b < batch) {
        for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) {
            const float q = queries[(size_t)b * dim + i];
            if (!(fabsf(q) <= 3.0e38f)) bad = true;  // NaN or inf
            amax = fmaxf(amax, fabsf(q));
        }
    }
    if (bad) atomicOr(&s_bad, 1);
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = amax;
    __syncthreads();
    const bool is_bad = s_bad != 0;
    amax = fmaxf(fmaxf(s_max[0], s_max[1]), fmaxf(s_max[2], s_max[3]));
    const bool zero = !(amax > 0.0f) || is_bad || b >= batch;
    const float scale = zero ? 0.0f : 127.0f / amax;
    const float sy = zero ? 0.0f : amax / 127.0f;
    double e2 = 0.0, n2 = 0.0;
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) {
        float q = 0.0f, c = 0.0f;
        if (!zero) {
            q = queries[(size_t)b * dim + i];
            c = fminf(fmaxf(roundf(__fmul_rn(q, scale)), -127.0f), 127.0f);
        }
        q_hat[(size_t)b * dim + i] = (int8_t)(int)c;
        const double d = (double)q - (double)sy * (double)c;
        e2 += d * d;
        n2 += (double)q * (double)q;
    }
    for (int o = 16; o > 0; o >>= 1) {
        e2 += __shfl_xor_sync(0xffffffffu, e2, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_e2[threadIdx.x >> 5] = e2;
        s_n2[threadIdx.x >> 5] = n2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const double ey = sqrt(s_e2[0] + s_e2[1] + s_e2[2] + s_e2[3]);
        const double nq = sqrt(s_n2[0] + s_n2[1] + s_n2[2] + s_n2[3]);
        const double d = (double)dim, R = (double)max_row_norm, Ex = (double)max_ex;
        const double e = ((R + Ex) * ey + Ex * nq + (d / 32.0 + 8.0) * (1.0 / 8388608.0) * R * nq +
                          (1.0 / 2097152.0) * (R + Ex) * (nq + ey)) * 1.01;
        float m2 = __double2float_ru(2.0 * e);
        // a finite query so large that the bound itself overflows cannot be covered: it is flagged BEFORE
        // the margin is clamped (flagging after the clamp could never fire)
        const bool overflow = !(m2 >= 0.0f) || !(m2 <= 3.0e38f);
        if (overflow || is_bad) m2 = 0.0f;
        margin2[b] = m2;
        qscale[b] = __fmul_rn(sx, sy);
        redo[b] = (b < batch && (is_bad || overflow)) ? 1u : 0u;
    }
}

// ─── the scan ───────────────────────────────────────────────────────────────────────────────
// Shared memory (1024-byte aligned): A[n_kblocks][16 KiB] | B[n_stages][16 KiB] | barriers.
__host__ __device__ inline size_t mma_scan_smem_bytes(uint32_t n_kblocks, uint32_t n_stages) {
    return 1024 + (size_t)(n_kblocks + n_stages) * kMmaTileBytes + 256;
}

__device__ __forceinline__ uint64_t mma_tile_of(const MmaScanArgs& args, uint64_t i) {
    if (args.tile_stride <= 1) return i;  // the full pass: no 64-bit modulo in the tile loop
    const uint64_t jitter = (uint64_t)(((uint32_t)i * 0x9E3779B1u) >> 8) % args.tile_stride;
    return i * args.tile_stride + jitter;
}

// Appends the rows of one 8-column group that clear the gate to this thread's private list (plain
// stores, no atomics, no returned value to wait for).  The common case (no tombstones, group inside the
// corpus) is STRAIGHT-LINE code: eight predicated 8-byte stores written in PTX, because from the C++
// form `if (p && c < cap) list[c] = e;` the compiler builds eight divergent branch regions
// (BSSY / BRA / BSYNC per column, ~190 instructions per group: 19 % of the int8 full pass went there,
// FSGPU_MMA_DBG=2).  The capacity test is made once per group (8 free slots, else the per-row form).
// The count keeps growing past `cap` so the consumer sees overflow.
// Value domain of the accumulators: f32 (kind::f16) or s32 (kind::i8; score = acc * qscale, the
// gate is compared in the integer domain).
template <bool I8>
struct MmaDom {
    using Gate = float;
    static __device__ __forceinline__ float val(uint32_t w) { return __uint_as_float(w); }
    static __device__ __forceinline__ float score(uint32_t w, float) { return __uint_as_float(w); }
};
template <>
struct MmaDom<true> {
    using Gate = int32_t;
    static __device__ __forceinline__ int32_t val(uint32_t w) { return (int32_t)w; }
    static __device__ __forceinline__ float score(uint32_t w, float qscale) {
        return __fmul_rn((float)(int32_t)w, qscale);  // |acc| <= dim * 127^2 < 2^24: the conversion is exact
    }
};

// one column: `if (hit) { list[c] = {score, row}; ++c; }` with a predicated store
__device__ __forceinline__ void mma_store_if(MmaCand* list, uint32_t& c, bool hit, float score, uint32_t row) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 a;\n"
        "setp.ne.u32 p, %2, 0;\n"
        "mad.wide.u32 a, %0, 8, %1;\n"
        "@p st.global.v2.b32 [a], {%3, %4};\n"
        "@p add.u32 %0, %0, 1;\n"
        "}\n"
        : "+r"(c)
        : "l"(list), "r"((uint32_t)hit), "r"(__float_as_uint(score)), "r"(row)
        : "memory");
}

template <bool I8>
__device__ __forceinline__ void mma_append8(const MmaScanArgs& args, MmaCand* list, uint32_t& count,
                                            const uint32_t (&w)[8], typename MmaDom<I8>::Gate gate, float qscale,
                                            uint64_t row0) {
    using D = MmaDom<I8>;
    const bool interior = row0 + 8u <= args.n_rows && args.tombstones == nullptr;
    const uint32_t grow0 = (uint32_t)(args.row_base + row0);
    uint32_t c = count;
    if (interior && c + 8u <= args.cap) {
        // a hot group nearly always holds exactly one row above the gate: pick it with selects and store once
        uint32_t m = 0u, one = w[0];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool p = D::val(w[i]) >= gate;
            m |= p ? (1u << i) : 0u;
            one = p ? w[i] : one;
        }
        if ((m & (m - 1u)) == 0u) {
            mma_store_if(list, c, m != 0u, D::score(one, qscale), grow0 + (uint32_t)__ffs(m) - 1u);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                mma_store_if(list, c, ((m >> i) & 1u) != 0u, D::score(w[i], qscale), grow0 + (uint32_t)i);
        }
    } else {  // tombstones, the corpus' last rows, or a list about to overflow: per-row tests
#pragma unroll  // (static indices: a rolled loop would move w[] to local memory for every caller)
        for (int i = 0; i < 8; ++i) {
            const uint64_t row = row0 + (uint32_t)i;
            if (D::val(w[i]) >= gate && row < args.n_rows && !tombstoned(args.tombstones, row)) {
                if (c < args.cap) {
                    MmaCand e;
                    e.score = D::score(w[i], qscale);
                    e.row = grow0 + (uint32_t)i;
                    list[c] = e;
                }
                ++c;
            }
        }
    }
    count = c;
}

// Max of 8 accumulator columns (3-input max tree): one bit of the thread's "hot group" mask.
__device__ __forceinline__ uint32_t mma_hot_bit(const uint32_t (&v)[32], int base, float gate, uint32_t bit) {
    const float m1 = fmaxf(fmaxf(__uint_as_float(v[base + 0]), __uint_as_float(v[base + 1])),
                           __uint_as_float(v[base + 2]));
    const float m2 = fmaxf(fmaxf(__uint_as_float(v[base + 3]), __uint_as_float(v[base + 4])),
                           __uint_as_float(v[base + 5]));
    const float m3 = fmaxf(fmaxf(__uint_as_float(v[base + 6]), __uint_as_float(v[base + 7])), m1);
    return fmaxf(m2, m3) >= gate ? bit : 0u;
}
__device__ __forceinline__ uint32_t mma_hot_bit(const uint32_t (&v)[32], int base, int32_t gate, uint32_t bit) {
    const int32_t m1 = max(max((int32_t)v[base + 0], (int32_t)v[base + 1]), (int32_t)v[base + 2]);
    const int32_t m2 = max(max((int32_t)v[base + 3], (int32_t)v[base + 4]), (int32_t)v[base + 5]);
    const int32_t m3 = max(max((int32_t)v[base + 6], (int32_t)v[base + 7]), m1);
    return max(m2, m3) >= gate ? bit : 0u;
}
template <class G>
__device__ __forceinline__ uint32_t mma_hot_bits(const uint32_t (&v)[32], G gate, uint32_t shift) {
    return mma_hot_bit(v, 0, gate, 1u << shift) | mma_hot_bit(v, 8, gate, 2u << shift) |
           mma_hot_bit(v, 16, gate, 4u << shift) | mma_hot_bit(v, 24, gate, 8u << shift);
}

// One accumulator tile of COLS columns for this thread's query.  Fast path: COLS/8 groups of 8
// columns -> one "some column clears the gate" bit each.  Slow path (rare): the warp re-reads each
// hot group and the owning lanes append — one compact copy of the append code instead of COLS
// unrolled ones (instruction cache).
template <int COLS, bool I8>
__device__ __forceinline__ void mma_epilogue_tile(const MmaScanArgs& args, uint32_t taddr, uint64_t tile_row0,
                                                  typename MmaDom<I8>::Gate gate, float qscale, MmaCand* list,
                                                  uint32_t& count) {
    static_assert(COLS % 64 == 0 && COLS <= 256, "hot mask is 32 bits of 8-column groups");
    uint32_t hot = 0;
    if constexpr (COLS <= 128) {
        // all chunks of the tile in flight at once (tcgen05.wait::ld waits for every outstanding load,
        // so a two-deep software pipeline serialises them: ~4 TMEM round trips per tile, which the
        // int8 form's 1536-cycle tile period does not hide — tensor pipe 70 % active before this)
        uint32_t v[COLS / 32][32];
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) tmem_ld_x32(taddr + c * 32u, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) hot |= mma_hot_bits(v[c], gate, 4 * c);
    } else {
        uint32_t va[32], vb[32];
        tmem_ld_x32(taddr, va);
#pragma unroll
        for (int c = 0; c < COLS / 32; c += 2) {
            tmem_ld_wait();
            tmem_ld_x32(taddr + (c + 1) * 32u, vb);  // in flight while the previous chunk is reduced
            hot |= mma_hot_bits(va, gate, 4 * c);
            tmem_ld_wait();
            if (c + 2 < COLS / 32) tmem_ld_x32(taddr + (c + 2) * 32u, va);
            hot |= mma_hot_bits(vb, gate, 4 * (c + 1));
        }
    }
    uint32_t hot_warp = __reduce_or_sync(0xffffffffu, hot);
    if (hot_warp == 0u) return;
    if (args.dbg & 2u) {
        count += __popc(hot);
        return;
    }
    // two groups in flight: the TMEM read of the next hot group overlaps the appends of this one
    auto check8 = [&](const uint32_t (&w)[8], uint32_t grp) {
        if (hot & (1u << grp)) mma_append8<I8>(args, list, count, w, gate, qscale, tile_row0 + grp * 8u);
    };
    uint32_t wa[8], wb[8];
    uint32_t ga = __ffs(hot_warp) - 1u, gb;
    hot_warp &= hot_warp - 1u;
    tmem_ld_x8(taddr + ga * 8u, wa);
#pragma unroll 1
    while (true) {
        tmem_ld_wait();
        gb = 0xFFFFFFFFu;
        if (hot_warp) {
            gb = __ffs(hot_warp) - 1u;
            hot_warp &= hot_warp - 1u;
            tmem_ld_x8(taddr + gb * 8u, wb);
        }
        check8(wa, ga);
        if (gb == 0xFFFFFFFFu) break;
        tmem_ld_wait();
        ga = 0xFFFFFFFFu;
        if (hot_warp) {
            ga = __ffs(hot_warp) - 1u;
            hot_warp &= hot_warp - 1u;
            tmem_ld_x8(taddr + ga * 8u, wa);
        }
        check8(wb, gb);
        if (ga == 0xFFFFFFFFu) break;
    }
}

// First sample level ("dump"): its list only feeds the k'-th-best selection that becomes the next
// gate, so instead of every score of the sample it keeps the best live score of each 8-row group.
// The k' largest group maxima belong to k' distinct live rows, hence the k'-th largest of them is
// still a lower bound of the corpus' k'-th best — and a sample 8x as large costs the same number of
// appends, which makes that gate ~8x tighter (the second level then runs at full speed instead
// of on the epilogue's slow path).  The row stored with a group maximum is the group's first row.
template <int COLS, bool I8>
__device__ __forceinline__ void mma_epilogue_dump_max(const MmaScanArgs& args, uint32_t taddr, uint64_t tile_row0,
                                                      bool live, float qscale, MmaCand* list, uint32_t& count) {
    // every lane of the warp runs the TMEM loads (tcgen05.ld is warp-collective); `live` only
    // predicates the appends
    using D = MmaDom<I8>;
    uint32_t v[32];
#pragma unroll 1
    for (int c = 0; c < COLS / 32; ++c) {
        tmem_ld_x32(taddr + c * 32u, v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const uint64_t row0 = tile_row0 + (uint32_t)(c * 32 + g * 8);
            uint32_t dead = 0u;  // bit i: column i does not count (past the corpus, or tombstoned)
            if (row0 >= args.n_rows)
                dead = 0xFFu;
            else if (row0 + 8u > args.n_rows)
                dead = (0xFFu << (uint32_t)(args.n_rows - row0)) & 0xFFu;
            if (args.tombstones && row0 < args.n_rows) dead |= __ldg(args.tombstones + (row0 >> 3));  // row0 % 8 == 0
            float m = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (!((dead >> i) & 1u)) m = fmaxf(m, D::score(v[g * 8 + i], qscale));
            if (live && (dead & 0xFFu) != 0xFFu) {
                if (count < args.cap) {
                    MmaCand e;
                    e.score = m;
                    e.row = (uint32_t)(args.row_base + row0);
                    list[count] = e;
                }
                ++count;
            }
        }
    }
}

// This thread's gate in the accumulator domain.  f32: the gate itself.  s32: the largest integer
// bound that keeps every accumulator whose score (acc * qscale, rounded) reaches the gate.
template <bool I8>
__device__ __forceinline__ typename MmaDom<I8>::Gate mma_thread_gate(const MmaScanArgs& args, uint32_t query, bool live,
                                                                    float* qscale_out) {
    if constexpr (!I8) {
        *qscale_out = 1.0f;
        return live ? (args.gate ? args.gate[query] : -INFINITY) : INFINITY;
    } else {
        const float qs = live ? args.qscale[query] : 0.0f;
        *qscale_out = qs;
        if (!live) return 0x7FFFFFFF;
        if (!args.gate) return (int32_t)0x80000000;
        const float gf = args.gate[query];
        if (!(qs > 0.0f)) return gf <= 0.0f ? (int32_t)0x80000000 : 0x7FFFFFFF;  // every score is 0
        const float t = __fdiv_rd(gf, qs);
        if (!(t > -2.0e9f)) return (int32_t)0x80000000;
        if (t >= 2.0e9f) return 0x7FFFFFFF;
        return (int32_t)floorf(t) - 2;  // two units of slack cover the roundings of t and of acc * qscale
    }
}

template <bool I8>
__global__ void __launch_bounds__(kMmaThreads, 1)
mma_scan_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x,
                const MmaScanArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    constexpr uint32_t kElems = I8 ? 128u : 64u;  // elements per 128-byte K-block row
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t a_smem = base;
    const uint32_t b_smem = a_smem + args.n_kblocks * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + (size_t)(args.n_kblocks + args.n_stages) * kMmaTileBytes);
    // barrier slots: [0..8) full, [8..16) empty, [16..20) tmem_full, [20..24) tmem_empty, 24 a_full
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (20u + a); };
    const uint32_t afull_bar = bar0 + 8u * 24u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t qb = blockIdx.x % args.n_qblocks;
    const uint32_t j0 = blockIdx.x / args.n_qblocks;
    const uint32_t g = args.ctas_per_qblock;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_x);
        for (uint32_t s = 0; s < args.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < kMmaAccStages; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), kMmaEpiWarps);  // one arrival per epilogue warp
        }
        mbar_init(afull_bar, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (whole warp loops, one elected lane issues) =====
        if (elect_one()) {
            mbar_expect_tx(afull_bar, args.n_kblocks * kMmaTileBytes);
            for (uint32_t kb = 0; kb < args.n_kblocks; ++kb)
                tma_load_2d(a_smem + kb * kMmaTileBytes, &tm_q, afull_bar, (int32_t)(kb * kElems),
                            (int32_t)(qb * kMmaM));
        }
        __syncwarp();
        uint32_t stage = 0, phase = 0;
        for (uint64_t i = j0; i < args.tile_count; i += g) {
            const int32_t row_coord = (int32_t)(mma_tile_of(args, i) * kMmaN);
            for (uint32_t kb = 0; kb < args.n_kblocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    mbar_expect_tx(full_bar(stage), kMmaTileBytes);
                    tma_load_2d(b_smem + stage * kMmaTileBytes, &tm_x, full_bar(stage), (int32_t)(kb * kElems),
                                row_coord);
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp loops, one elected lane issues) =====
        constexpr uint32_t idesc = I8 ? umma_idesc_i8(kMmaM, kMmaN) : umma_idesc_f16(kMmaM, kMmaN);
        const uint64_t a_desc0 = umma_desc_sw128(a_smem);
        const uint64_t b_desc0 = umma_desc_sw128(b_smem);
        mbar_wait(afull_bar, 0);
        tc_fence_after();
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (uint64_t i = j0; i < args.tile_count; i += g) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kMmaN;
            for (uint32_t kb = 0; kb < args.n_kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    // descriptors advance in units of 16 bytes: 1024 per K-block tile, 2 per 16 elements
                    const uint64_t a_desc = a_desc0 + (uint64_t)(kb * (kMmaTileBytes >> 4));
                    const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (kMmaTileBytes >> 4));
#pragma unroll
                    for (uint32_t k4 = 0; k4 < 4; ++k4) {  // four 32-byte K-steps per 128-byte K-block
                        if constexpr (I8)
                            umma_i8(d_tmem, a_desc + 2u * k4, b_desc + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                        else
                            umma_f16(d_tmem, a_desc + 2u * k4, b_desc + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the B stage when these MMAs retire
                    if (kb + 1 == args.n_kblocks) umma_commit(tfull_bar(acc));  // accumulator complete
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (++acc == kMmaAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: TMEM lane = query, column = corpus row =====
        const uint32_t quarter = warp & 3u;  // the TMEM lane quarter this warp may read
        const uint32_t m = quarter * 32u + lane;
        const uint32_t query = qb * kMmaM + m;
        const bool live = query < args.batch && args.redo[query] == 0u;
        float qscale;
        const typename MmaDom<I8>::Gate gate = mma_thread_gate<I8>(args, query, live, &qscale);
        const uint32_t half = (warp - 2u) >> 2;  // which half of the tile's columns this warp checks
        const size_t list_id = ((size_t)blockIdx.x * 2u + half) * kMmaM + m;
        MmaCand* list = args.cand + list_id * args.cap;
        uint32_t count = 0;
        uint32_t acc = 0, acc_phase = 0;
        for (uint64_t i = j0; i < args.tile_count; i += g) {
            const uint64_t tile = mma_tile_of(args, i);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kMmaN + half * (kMmaN / 2);
            if (args.dump_group_max) {
                mma_epilogue_dump_max<kMmaN / 2, I8>(args, taddr, tile * kMmaN + half * (kMmaN / 2), live, qscale, list, count);
            } else {
                mma_epilogue_tile<kMmaN / 2, I8>(args, taddr, tile * kMmaN + half * (kMmaN / 2), gate, qscale, list, count);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));  // accumulator drained -> MMA may reuse it
            if (++acc == kMmaAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        args.cand_count[list_id] = count;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ─── the scan, CTA-pair form ────────────────────────────────────────────────────────────────
// Two CTAs of a cluster (the two SMs of a TPC) run ONE tcgen05.mma.cta_group::2 of M = 256
// queries (128 per CTA, each CTA's q_hat tile in its own shared memory and its accumulators in
// its own TMEM) by N = 256 corpus rows (each CTA TMA-loads 128 of them).  Every slab byte is
// fetched from L2 once per 256 queries and read from shared memory once per 256-query MMA: half
// the L2->SM and shared-memory traffic per flop of the single-CTA form, which ran at the L2->SM
// fabric limit (profiles/r01_mma_v5_b1024_ncu.json: 54 % tensor-pipe active, time unchanged when
// half the MMAs were dropped).  Rank 0 (leader) issues the MMAs; barriers: `full`/`a_full`/
// `tmem_empty` live on the leader, `empty`/`tmem_full` are signalled in both CTAs by multicast
// commits.
constexpr int kPairN = 256;
constexpr int kPairAccStages = 2;  // 2 x 256 TMEM columns

template <bool I8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMmaThreads, 1)
mma_scan_pair_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x,
                     const MmaScanArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    constexpr uint32_t kElems = I8 ? 128u : 64u;  // elements per 128-byte K-block row
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t a_smem = base;
    const uint32_t b_smem = a_smem + args.n_kblocks * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + (size_t)(args.n_kblocks + args.n_stages) * kMmaTileBytes);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (20u + a); };
    const uint32_t afull_bar = bar0 + 8u * 24u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1;
    const uint32_t n_qpairs = args.n_qblocks >> 1;  // n_qblocks is even
    const uint32_t qb = (pair % n_qpairs) * 2u + rank;
    const uint32_t j0 = pair / n_qpairs;
    const uint32_t g = args.ctas_per_qblock;  // CTA pairs per query pair

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_x);
        for (uint32_t s = 0; s < args.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < kPairAccStages; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 2 * kMmaEpiWarps);  // one arrival per epilogue warp of BOTH CTAs
        }
        mbar_init(afull_bar, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs; completion bytes land on the leader's barriers) =====
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(afull_bar, 2u * args.n_kblocks * kMmaTileBytes);
            for (uint32_t kb = 0; kb < args.n_kblocks; ++kb)
                tma_load_2d_pair(a_smem + kb * kMmaTileBytes, &tm_q, afull_bar, (int32_t)(kb * kElems),
                                 (int32_t)(qb * kMmaM));
        }
        __syncwarp();
        uint32_t stage = 0, phase = 0, li = 0;
        volatile uint32_t* prog = args.progress ? args.progress + (size_t)j0 * n_qpairs : nullptr;
        for (uint64_t i = j0; i < args.tile_count; i += g, ++li) {
            const int32_t row_coord = (int32_t)(mma_tile_of(args, i) * kPairN + rank * kMmaN);
            // every 8th tile (an L2 round trip per tile would eat the TMA prefetch depth): wait for
            // the slowest pair of this tile stream
            if (prog && li > args.lead && (li & 7u) == 0u && lane == 0) {
                const long long t0 = clock64();
                for (uint32_t spins = 0;; ++spins) {
                    uint32_t slowest = 0xFFFFFFFFu;
                    for (uint32_t q = 0; q < n_qpairs; ++q) slowest = min(slowest, prog[q]);
                    if (slowest + args.lead >= li) break;
                    __nanosleep(200);
                    if ((spins & 255u) == 255u && clock64() - t0 > 8000000000ll) __trap();
                }
            }
            __syncwarp();
            for (uint32_t kb = 0; kb < args.n_kblocks; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * kMmaTileBytes);
                    tma_load_2d_pair(b_smem + stage * kMmaTileBytes, &tm_x, full_bar(stage),
                                     (int32_t)(kb * kElems), row_coord);
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (prog && rank == 0 && lane == 0 && (li & 3u) == 3u) prog[pair % n_qpairs] = li + 1u;
        }
        if (prog && rank == 0 && lane == 0) prog[pair % n_qpairs] = 0xFFFFFFF0u;  // done: never the slowest
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer (leader CTA only) =====
            constexpr uint32_t idesc = I8 ? umma_idesc_i8(2 * kMmaM, kPairN) : umma_idesc_f16(2 * kMmaM, kPairN);
            const uint64_t a_desc0 = umma_desc_sw128(a_smem);
            const uint64_t b_desc0 = umma_desc_sw128(b_smem);
            mbar_wait(afull_bar, 0);
            tc_fence_after();
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint64_t i = j0; i < args.tile_count; i += g) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kPairN;
                for (uint32_t kb = 0; kb < args.n_kblocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_desc = a_desc0 + (uint64_t)(kb * (kMmaTileBytes >> 4));
                        const uint64_t b_desc = b_desc0 + (uint64_t)(stage * (kMmaTileBytes >> 4));
#pragma unroll
                        for (uint32_t k4 = 0; k4 < 4; ++k4) {
                            if constexpr (I8)
                                umma_i8_pair(d_tmem, a_desc + 2u * k4, b_desc + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                            else
                                umma_f16_pair(d_tmem, a_desc + 2u * k4, b_desc + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                        }
                        umma_commit_pair(empty_bar(stage));  // frees this stage in BOTH CTAs
                        if (kb + 1 == args.n_kblocks) umma_commit_pair(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == args.n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++acc == kPairAccStages) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs): TMEM lane = query of this CTA, column = row of the pair tile =====
        const uint32_t quarter = warp & 3u;
        const uint32_t m = quarter * 32u + lane;
        const uint32_t query = qb * kMmaM + m;
        const bool live = query < args.batch && args.redo[query] == 0u;
        float qscale;
        const typename MmaDom<I8>::Gate gate = mma_thread_gate<I8>(args, query, live, &qscale);
        const uint32_t half = (warp - 2u) >> 2;
        const size_t list_id = ((size_t)blockIdx.x * 2u + half) * kMmaM + m;
        MmaCand* list = args.cand + list_id * args.cap;
        uint32_t count = 0;
        uint32_t acc = 0, acc_phase = 0;
        for (uint64_t i = j0; i < args.tile_count; i += g) {
            const uint64_t tile = mma_tile_of(args, i);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kPairN + half * (kPairN / 2);
            if (args.dump_group_max) {
                mma_epilogue_dump_max<kPairN / 2, I8>(args, taddr, tile * kPairN + half * (kPairN / 2), live, qscale, list, count);
            } else {
                mma_epilogue_tile<kPairN / 2, I8>(args, taddr, tile * kPairN + half * (kPairN / 2), gate, qscale, list, count);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);  // leader may reuse the accumulator
            if (++acc == kPairAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        args.cand_count[list_id] = count;
    }

    tc_fence_before();
    cluster_sync_all();  // neither CTA may leave while the other can still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ─── the scan, CTA-pair form with TWO query blocks per CTA (int8 only) ───────────────────────
// The int8 form of mma_scan_pair_kernel retires a 256-row tile in half the time, so at the same
// bytes per tile it asks the L2->SM fabric for twice the bandwidth — and stalls on it (tensor pipe
// 70 % active at 148 SMs x 32 B/clk x 1.9 GHz ~ 9 TB/s, profiles/r01_mma_pair_i8_b1024_ncu.json).
// Here every B stage is used by two MMAs groups: the pair holds 512 queries (two 128-query blocks
// per CTA, int8 so they fit: 2 x 48 KB at D = 384) and one accumulator per block; block 0's MMAs
// over a tile are followed by block 1's over the same stages, and each block's epilogue overlaps the
// other block's MMAs.  Half the L2->SM bytes per MMA.
// kEpiWarps = 16 (default; 8 = the first form): four epilogue warps per TMEM lane quarter, each takes a
// quarter (64) of an accumulator's columns.  A sub-block's epilogue must fit inside the other sub-block's
// 1536-cycle MMA group, and with two warps per scheduler it did not: ~330 cycles of TMEM reads
// (~100 B/clk per scheduler, tools/ubench_tmem.cu) that nothing overlapped, ~440 of gate tests on the
// half-rate integer pipe and ~350 of appends (half of all warp tiles hold a candidate) — tensor pipe 72-80 %
// active, epilogue warps busy 60-80 % of the time at an IPC of 0.15 (profiles/r02_mma_quad_i8_b1024_ncu.json).
template <uint32_t kEpiWarps>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * kEpiWarps, 1)
mma_scan_quad_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_x,
                     const MmaScanArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    constexpr uint32_t kElems = 128u;  // int8 codes per 128-byte K-block row
    constexpr uint32_t kSub = 2;       // query blocks per CTA
    constexpr uint32_t kParts = kEpiWarps / 4;      // column parts of an accumulator (one epilogue warp each per lane quarter)
    constexpr uint32_t kCols = kPairN / kParts;     // columns per epilogue warp and accumulator
    static_assert(kEpiWarps == 8 || kEpiWarps == 16, "two or four epilogue warps per TMEM lane quarter");
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t a_smem = base;
    const uint32_t b_smem = a_smem + kSub * args.n_kblocks * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + (size_t)(kSub * args.n_kblocks + args.n_stages) * kMmaTileBytes);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (20u + a); };
    const uint32_t afull_bar = bar0 + 8u * 24u;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1;
    const uint32_t n_qpairs = args.n_qblocks >> 2;  // query QUADS (4 blocks of 128): n_qblocks is a multiple of 4
    const uint32_t quad = pair % n_qpairs;
    const uint32_t j0 = pair / n_qpairs;
    const uint32_t g = args.ctas_per_qblock;  // CTA pairs per query quad
    // this CTA's query block of sub-block s: ((quad * 2 + s) * 2 + rank)
    auto qb_of = [&](uint32_t sub) { return (quad * 2u + sub) * 2u + rank; };
    // tile i of the full pass was scanned by the carried sample level (32-bit restatement of mma_tile_of)
    const uint32_t skip_s = args.skip_stride, skip_c = args.skip_count;
    auto skipped = [&](uint64_t i) -> bool {
        if (!skip_c) return false;
        const uint32_t t = (uint32_t)i, idx = t / skip_s;
        if (idx >= skip_c) return false;
        const uint32_t jitter = skip_s > 1u ? ((idx * 0x9E3779B1u) >> 8) % skip_s : 0u;
        return t == idx * skip_s + jitter;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_x);
        for (uint32_t s = 0; s < args.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < kSub; ++a) {  // one accumulator per sub-block
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 2 * kEpiWarps);  // one arrival per epilogue warp of BOTH CTAs
        }
        mbar_init(afull_bar, 1);
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (args.ts && blockIdx.x < 2 && threadIdx.x == 0) args.ts[blockIdx.x] = clock64();  // clock offset between the two SMs

    // Roles by warp id: the scheduler of an SM sub-partition prefers its eligible warp with the HIGHEST id, so the
    // two latency-critical single-warp roles take the top ids (as warps 0 and 1 they lost every issue slot the
    // epilogue warps of their sub-partitions wanted).
    if (warp == kEpiWarps) {
        // ===== TMA producer (both CTAs; completion bytes land on the leader's barriers) =====
        if (elect_one()) {
            if (rank == 0) mbar_expect_tx(afull_bar, 2u * kSub * args.n_kblocks * kMmaTileBytes);
            for (uint32_t sub = 0; sub < kSub; ++sub)
                for (uint32_t kb = 0; kb < args.n_kblocks; ++kb)
                    tma_load_2d_pair(a_smem + (sub * args.n_kblocks + kb) * kMmaTileBytes, &tm_q, afull_bar,
                                     (int32_t)(kb * kElems), (int32_t)(qb_of(sub) * kMmaM));
        }
        __syncwarp();
        uint32_t stage = 0, phase = 0, li = 0;
        volatile uint32_t* prog = args.progress ? args.progress + (size_t)j0 * n_qpairs : nullptr;
        for (uint64_t i = j0; i < args.tile_count; i += g, ++li) {
            const int32_t row_coord = (int32_t)(mma_tile_of(args, i) * kPairN + rank * kMmaN);
            // every 8th tile (an L2 round trip per tile would eat the TMA prefetch depth): wait for
            // the slowest pair of this tile stream
            if (prog && li > args.lead && (li & 7u) == 0u && lane == 0) {
                const long long t0 = clock64();
                for (uint32_t spins = 0;; ++spins) {
                    uint32_t slowest = 0xFFFFFFFFu;
                    for (uint32_t q = 0; q < n_qpairs; ++q) slowest = min(slowest, prog[q]);
                    if (slowest + args.lead >= li) break;
                    __nanosleep(200);
                    if ((spins & 255u) == 255u && clock64() - t0 > 8000000000ll) __trap();
                }
            }
            __syncwarp();
            for (uint32_t kb = 0; kb < (skipped(i) ? 0u : args.n_kblocks); ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * kMmaTileBytes);
                    tma_load_2d_pair(b_smem + stage * kMmaTileBytes, &tm_x, full_bar(stage),
                                     (int32_t)(kb * kElems), row_coord);
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (prog && rank == 0 && lane == 0 && (li & 3u) == 3u) prog[quad] = li + 1u;
        }
        if (prog && rank == 0 && lane == 0) prog[quad] = 0xFFFFFFF0u;  // done: never the slowest
    } else if (warp == kEpiWarps + 1) {
        if (rank == 0) {
            // ===== MMA issuer (leader CTA only) =====
            constexpr uint32_t idesc = umma_idesc_i8(2 * kMmaM, kPairN);
            const uint64_t a_desc0 = umma_desc_sw128(a_smem);
            const uint64_t b_desc0 = umma_desc_sw128(b_smem);
            mbar_wait(afull_bar, 0);
            tc_fence_after();
            uint32_t stage = 0, phase = 0, acc_phase = 0;
            const uint32_t n_kb = args.n_kblocks, n_st = args.n_stages;
            uint32_t li = 0;
            for (uint64_t i = j0; i < args.tile_count; i += g, ++li) {
                // sub-block 0 then sub-block 1 over the SAME B stages: the epilogue of one sub-block's
                // accumulator overlaps the MMAs of the other (one accumulator each, 2 x 256 columns).
                // ONE elected block per MMA group (all of a group's K-blocks, <= 16 MMAs, and its commits): with an
                // elect + reconvergence + eight vector->uniform register moves per K-block the issuer needed
                // 700-1200 cycles per group (timestamps, FSGPU_MMA_TS), and a group whose last four MMAs (512
                // cycles) are issued that late cannot finish inside its 1536-cycle slot.
                if (skipped(i)) continue;
                const uint32_t stage0 = stage;
                for (uint32_t kb = 0; kb < n_kb; ++kb) {  // the tile's stages have landed (they are requested together)
                    mbar_wait(full_bar(stage), phase);
                    if (++stage == n_st) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
#pragma unroll
                for (uint32_t sub = 0; sub < kSub; ++sub) {
                    const uint32_t tsi = li - 64u;  // (li = tiles of this stream so far: the stamps cover tiles 64..127)
                    long long* ts = (args.ts && blockIdx.x == 0 && tsi < 64u && lane == 0) ? args.ts + 8 + (tsi * 2u + sub) * 72u : nullptr;
                    if (ts) ts[0] = clock64();
                    mbar_wait(tempty_bar(sub), acc_phase ^ 1u);
                    if (ts) ts[1] = clock64();
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d_tmem = tmem_base + sub * kPairN;
                        const uint64_t a_desc = a_desc0 + (uint64_t)(sub * n_kb * (kMmaTileBytes >> 4));
#pragma unroll
                        for (uint32_t kb = 0; kb < kMmaMaxDim / 128u; ++kb) {
                            if (kb < n_kb) {
                                const uint32_t st = stage0 + kb >= n_st ? stage0 + kb - n_st : stage0 + kb;
                                const uint64_t b_desc = b_desc0 + (uint64_t)(st * (kMmaTileBytes >> 4));
#pragma unroll
                                for (uint32_t k4 = 0; k4 < 4; ++k4)
                                    umma_i8_pair(d_tmem, a_desc + (uint64_t)(kb * (kMmaTileBytes >> 4) + 2u * k4), b_desc + 2u * k4,
                                                 idesc, (kb | k4) != 0u ? 1u : 0u);
                                if (sub + 1 == kSub) umma_commit_pair(empty_bar(st));  // both sub-blocks have read it
                            }
                        }
                        umma_commit_pair(tfull_bar(sub));
                    }
                    __syncwarp();
                    if (ts) ts[2] = clock64();
                }
                acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue (both CTAs): TMEM lane = query of this CTA, column = row of the pair tile =====
        const uint32_t quarter = warp & 3u;
        const uint32_t m = quarter * 32u + lane;
        const uint32_t part = warp >> 2;  // columns [part * kCols, +kCols) of the pair tile = its rows
        bool live[kSub];
        float qscale[kSub];
        int32_t gate[kSub];
        MmaCand* list[kSub];
        size_t list_id[kSub];
        uint32_t count[kSub];
#pragma unroll
        for (uint32_t sub = 0; sub < kSub; ++sub) {
            const uint32_t query = qb_of(sub) * kMmaM + m;
            live[sub] = query < args.batch && args.redo[query] == 0u;
            gate[sub] = mma_thread_gate<true>(args, query, live[sub], &qscale[sub]);
            list_id[sub] = (((size_t)blockIdx.x * kSub + sub) * kParts + part) * kMmaM + m;
            list[sub] = args.cand + list_id[sub] * args.cap;
            count[sub] = 0;
            if (args.carry) {
                const uint32_t old = args.cand_count[list_id[sub]];
                if (old > args.cap) {
                    count[sub] = old;  // the sample level overflowed this list: it stays flagged (the query is redone)
                } else if (live[sub]) {
                    // keep what clears this pass's gate (four accumulator units of slack: a superset is always safe)
                    const float gf = args.gate ? args.gate[query] - 4.0f * fabsf(qscale[sub]) : -INFINITY;
                    uint32_t n = 0;
                    for (uint32_t e = 0; e < old; ++e) {
                        const MmaCand c = list[sub][e];
                        if (c.score >= gf) list[sub][n++] = c;
                    }
                    count[sub] = n;
                }
            }
        }
        uint32_t acc_phase = 0, li = 0;
        for (uint64_t i = j0; i < args.tile_count; i += g, ++li) {
            if (skipped(i)) continue;
            const uint64_t tile = mma_tile_of(args, i);
#pragma unroll
            for (uint32_t sub = 0; sub < kSub; ++sub) {
                const uint32_t tsi = li - 64u;
                long long* ts = (args.ts && blockIdx.x < 2 && tsi < 64u && lane == 0)
                                    ? args.ts + 8 + (tsi * 2u + sub) * 72u + 4 + (blockIdx.x * 8 + warp) * 4 : nullptr;
                mbar_wait(tfull_bar(sub), acc_phase);
                if (ts) ts[0] = clock64();
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + sub * kPairN + part * kCols;
                const uint64_t row0 = tile * kPairN + part * kCols;
                auto release = [&]() {  // leader may reuse this accumulator
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_bar(sub), 0);
                    if (ts) ts[2] = clock64();
                };
                if (args.dbg & 1u) {
                    release();
                } else if (args.dump_group_max) {
                    mma_epilogue_dump_max<kCols, true>(args, taddr, row0, live[sub], qscale[sub], list[sub], count[sub]);
                    release();
                } else {
                    mma_epilogue_tile<kCols, true>(args, taddr, row0, gate[sub], qscale[sub], list[sub], count[sub]);
                    release();
                }
                if (ts) ts[3] = clock64();
            }
            acc_phase ^= 1u;
        }
#pragma unroll
        for (uint32_t sub = 0; sub < kSub; ++sub) args.cand_count[list_id[sub]] = count[sub];
    }

    tc_fence_before();
    cluster_sync_all();  // neither CTA may leave while the other can still signal it
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ─── candidate lists as the gate / refine kernels see them ──────────────────────────────────
struct MmaLists {
    const MmaCand* cand;         // [grid][lists per CTA][128][cap]
    const uint32_t* cand_count;  // [grid][lists per CTA][128]
    uint32_t n_qblocks, ctas_per_qblock, cap;
    uint32_t pair;               // 1: lists were written by mma_scan_pair_kernel, 2: by mma_scan_quad_kernel
    uint32_t parts;              // lists per (CTA, query): one per epilogue warp of a lane quarter (2, quad form: 2 or 4)
};
// Query slot b has parts*ctas_per_qblock lists.  List j lives in CTA c(j/parts), part j%parts, where
// c(i) = qb + n_qblocks*i (single-CTA form) or 2*(qb/2 + (n_qblocks/2)*i) + qb%2 (pair form).
__device__ __forceinline__ uint32_t mma_list_count(const MmaLists& l) { return l.parts * l.ctas_per_qblock; }
__device__ __forceinline__ size_t mma_list_slot(const MmaLists& l, uint32_t b, uint32_t j) {
    const uint32_t qb = b / kMmaM, i = j / l.parts, part = j % l.parts;
    if (l.pair == 2u) {  // quad form: blocks (quad*2 + sub)*2 + rank, lists [cta][sub][part][128]
        const size_t cta = 2 * ((size_t)(qb >> 2) + (size_t)(l.n_qblocks >> 2) * i) + (qb & 1u);
        return ((cta * 2u + ((qb >> 1) & 1u)) * l.parts + part) * kMmaM + (b % kMmaM);
    }
    const size_t cta = l.pair ? 2 * ((size_t)(qb >> 1) + (size_t)(l.n_qblocks >> 1) * i) + (qb & 1u)
                              : (size_t)qb + (size_t)l.n_qblocks * i;
    return (cta * l.parts + part) * kMmaM + (b % kMmaM);
}

// ─── staging + radix select ─────────────────────────────────────────────────────────────────
// A query's candidates are spread over 2*ctas_per_qblock short lists.  Walking them one after
// another costs two dependent L2 round trips per list (36-296 lists): the first versions of the
// gate/refine kernels spent 100-500 us there.  Instead every warp takes whole lists in parallel
// and the entries are flattened into shared memory once.
constexpr uint32_t kMmaMaxLists = 2 * 160;      // >= 2 * SM count (quad form with 4 parts: 4 * 74 CTA pairs)
constexpr uint32_t kMmaStageScores = 16384;     // gate kernel: ordered scores only (64 KiB)
constexpr uint32_t kMmaStagePairs = 16384;      // refine kernel: ordered score + row (128 KiB); <= 64 * 256 (flag word)

struct MmaStageSmem {
    uint32_t* hist;   // [256]
    uint32_t* ctl;    // [4]
    uint32_t* offs;   // [kMmaMaxLists + 1] exclusive offsets of the lists in the flattened array
    uint32_t* score;  // [cap] ascending total-order image of the approximate scores
    uint32_t* row;    // [cap] or nullptr
    uint32_t cap;
};
__host__ __device__ inline size_t mma_stage_smem_bytes(uint32_t cap, bool with_rows) {
    return (size_t)(256 + 4 + kMmaMaxLists + 4) * 4 + (size_t)cap * (with_rows ? 8 : 4);
}
__device__ __forceinline__ MmaStageSmem carve_stage_smem(unsigned char* base, uint32_t cap, bool with_rows) {
    MmaStageSmem sm;
    sm.hist = reinterpret_cast<uint32_t*>(base);
    sm.ctl = sm.hist + 256;
    sm.offs = sm.ctl + 4;
    sm.score = sm.offs + kMmaMaxLists + 4;
    sm.row = with_rows ? sm.score + cap : nullptr;
    sm.cap = cap;
    return sm;
}

// CTA-collective.  Returns the total number of entries (lists clipped to their capacity); the
// flattened copy is valid iff total <= sm.cap.
__device__ __forceinline__ uint32_t mma_stage_lists(const MmaStageSmem& sm, const MmaLists& l, uint32_t b) {
    const uint32_t n_lists = mma_list_count(l);
    for (uint32_t j = threadIdx.x; j < n_lists; j += blockDim.x)
        sm.offs[j + 1] = min(l.cand_count[mma_list_slot(l, b, j)], l.cap);
    if (threadIdx.x == 0) sm.offs[0] = 0u;
    __syncthreads();
    if (threadIdx.x < 32) {  // warp 0: inclusive scan of the counts, 32 at a time
        uint32_t carry = 0;
        for (uint32_t base = 0; base < n_lists; base += 32) {
            const uint32_t j = base + threadIdx.x;
            uint32_t v = j < n_lists ? sm.offs[j + 1] : 0u;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
                if ((int)threadIdx.x >= o) v += t;
            }
            if (j < n_lists) sm.offs[j + 1] = v + carry;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();
    const uint32_t total = sm.offs[n_lists];
    if (total > sm.cap) return total;
    // one thread per flattened entry: its list is found by a binary search of the offsets, so every
    // global load of the copy is in flight at once (a warp walking whole lists paid a dependent L2 round
    // trip per list: 9-37 of them in a row)
    for (uint32_t i = threadIdx.x; i < total; i += blockDim.x) {
        uint32_t lo = 0, hi = n_lists;  // last j with offs[j] <= i
        while (hi - lo > 1u) {
            const uint32_t mid = (lo + hi) >> 1;
            if (sm.offs[mid] <= i) lo = mid; else hi = mid;
        }
        const MmaCand c = l.cand[mma_list_slot(l, b, lo) * l.cap + (i - sm.offs[lo])];
        sm.score[i] = ordered_score(c.score);
        if (sm.row) sm.row[i] = c.row;
    }
    __syncthreads();
    return total;
}

// Histogram update, warp-aggregated for the common case that every active lane hits one bin
// (scores of one query share their exponent byte).  All 32 lanes must call.
__device__ __forceinline__ void mma_hist_add(uint32_t* hist, bool active, uint32_t bin) {
    const uint32_t amask = __ballot_sync(0xffffffffu, active);
    if (amask == 0u) return;
    const uint32_t leader = __ffs(amask) - 1u;
    const uint32_t lbin = __shfl_sync(0xffffffffu, bin, leader);
    const uint32_t same = __ballot_sync(0xffffffffu, active && bin == lbin);
    if (same == amask) {
        if ((threadIdx.x & 31u) == leader) atomicAdd(&hist[lbin], __popc(amask));
    } else if (active) {
        atomicAdd(&hist[bin], 1u);
    }
}

// CTA-collective radix select (4 passes of 8 bits) of the k-th largest ordered score.  `staged`
// selects the flattened shared-memory copy, otherwise the lists are re-read from global memory
// every pass (only when they do not fit).  Requires total >= k.
__device__ __forceinline__ uint32_t mma_kth_best(const MmaStageSmem& sm, const MmaLists& l, uint32_t b, uint32_t k,
                                                 uint32_t total, bool staged) {
    const uint32_t step = blockDim.x;
    uint32_t prefix = 0, mask = 0, k_rem = k;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (uint32_t i = threadIdx.x; i < 256; i += step) sm.hist[i] = 0u;
        __syncthreads();
        if (staged) {
            for (uint32_t base = 0; base < total; base += step) {  // warp-uniform trip count
                const uint32_t i = base + threadIdx.x;
                const uint32_t u = i < total ? sm.score[i] : 0u;
                mma_hist_add(sm.hist, i < total && (u & mask) == prefix, (u >> shift) & 255u);
            }
        } else {
            for (uint32_t j = 0; j < mma_list_count(l); ++j) {
                const uint32_t n = sm.offs[j + 1] - sm.offs[j];
                const MmaCand* list = l.cand + mma_list_slot(l, b, j) * l.cap;
                for (uint32_t base = 0; base < n; base += step) {
                    const uint32_t i = base + threadIdx.x;
                    const uint32_t u = i < n ? ordered_score(list[i].score) : 0u;
                    mma_hist_add(sm.hist, i < n && (u & mask) == prefix, (u >> shift) & 255u);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {  // warp 0: the bin holding the k_rem-th largest; lane L owns bins [8L, 8L+8)
            uint32_t local[8], sum = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                local[i] = sm.hist[threadIdx.x * 8 + i];
                sum += local[i];
            }
            uint32_t above = 0;  // entries in bins owned by higher lanes
            for (int src = 31; src >= 0; --src) {
                const uint32_t v = __shfl_sync(0xffffffffu, sum, src);
                if (src > (int)threadIdx.x) above += v;
            }
            if (above < k_rem && above + sum >= k_rem) {  // exactly one lane
                uint32_t acc = above;
                for (int i = 7; i >= 0; --i) {
                    if (acc + local[i] >= k_rem) {
                        sm.ctl[0] = prefix | ((uint32_t)(threadIdx.x * 8 + i) << shift);
                        sm.ctl[1] = k_rem - acc;
                        break;
                    }
                    acc += local[i];
                }
            }
        }
        __syncthreads();
        prefix = sm.ctl[0];
        k_rem = sm.ctl[1];
        mask |= 255u << shift;
        __syncthreads();
    }
    return prefix;
}

// ─── gate: k'-th best approximate score of a level's lists -> the next level's static gate ──
// One CTA per query.  Any subset's k'-th best is a lower bound of the full corpus' k'-th (<= k-th)
// best, so the gate stays valid when a list overflowed (its first `cap` entries are used).
struct MmaGateArgs {
    MmaLists lists;
    const float* margin2;
    const uint32_t* redo;
    float* gate;            // out
    uint32_t k_sel;
    uint32_t stage_cap;     // staged scores per CTA (sized by the host from the expected list lengths)
    uint32_t keep_prev;     // `gate` already holds the previous level's (valid) gate: never go below it
};

__global__ void __launch_bounds__(256) mma_gate_kernel(const MmaGateArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t b = blockIdx.x;
    if (args.redo[b] != 0u) return;
    const MmaStageSmem sm = carve_stage_smem(smem_raw, args.stage_cap, false);
    const uint32_t total = mma_stage_lists(sm, args.lists, b);
    // a level that caught fewer than k' rows (its sample was unlucky) keeps the previous level's gate —
    // also a lower bound of the corpus' k'-th best, just a looser one — instead of opening the gate
    float gate = args.keep_prev ? args.gate[b] : -INFINITY;
    if (total >= args.k_sel) {  // CTA-uniform
        const uint32_t u = mma_kth_best(sm, args.lists, b, args.k_sel, total, total <= sm.cap);
        gate = fmaxf(gate, __fsub_rd(unordered_score(u), args.margin2[b]));
    }
    if (threadIdx.x == 0) args.gate[b] = gate;
}

// ─── exact gate from a sample level (int8 form) ─────────────────────────────────────────────
// `keys` holds, per query, the exact top-k of a SAMPLE of the corpus (refine run on that level's
// lists).  Their k-th best reference score tau_s is <= the corpus' k-th best, and every row of the
// true top-k has approx >= reference - e >= tau_s - e: a ONE-sided margin (e = margin2 / 2) below
// an exact score, where a gate taken from approximate scores needs 2e.  With the int8 bound
// (e ~ 0.25 sigma of the score distribution) that is ~4x fewer candidates in the full pass.
__global__ void __launch_bounds__(256)
mma_exact_gate_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint32_t k,
                      uint32_t slots, const float* __restrict__ margin2, float* __restrict__ gate) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= slots) return;
    if (counts[b] < k) return;  // keep the previous (valid) gate
    const float tau_s = key_score(keys[(size_t)b * k + (k - 1u)]);
    const float g = __fsub_rd(tau_s, __fmul_ru(margin2[b], 0.5f));
    gate[b] = fmaxf(gate[b], g);
}

// ─── refine: exact re-scoring of the candidate superset, one CTA per query ──────────────────
struct MmaRefineArgs {
    MmaLists lists;
    const float* margin2;        // [slots]
    uint32_t* redo;              // [slots]; set to 2 when a final list overflowed
    uint32_t k;
    uint32_t buf_cap;            // shared candidate buffer capacity (power of two)
    uint32_t stage_cap;          // staged (score, row) pairs per CTA
    const uint16_t* slab;
    const float* queries;        // [batch, dim] f32 (the ORIGINAL queries)
    uint64_t n_rows, row_base;
    uint32_t dim;
    int reduce_order, tail_fma;
    uint64_t* out_keys;          // [batch, k] (nullable)
    fsgpu_hit_t* out_hits;       // [batch, k] (nullable)
    uint32_t* out_counts;        // [batch] (nullable)
    uint32_t* error_flag;
    uint32_t* redo_any;          // set to 1 when any query of the launch needs the exact path
    uint32_t intermediate;       // 1: run on a SAMPLE level's lists to get an exact k-th best for the next
                                 // gate (mma_exact_gate_kernel); never flags a query, reports count 0 instead
};

// Re-scores (warp-cooperatively) the rows whose lanes hold `pass` and offers the exact keys.
__device__ __forceinline__ void mma_rescore_round(const MmaRefineArgs& args, const CandBuf& buf, const float* q,
                                                  bool pass, uint32_t grow) {
    uint32_t mask = __ballot_sync(0xffffffffu, pass);
    while (mask) {
        const uint32_t src = __ffs(mask) - 1u;
        mask &= mask - 1u;
        const uint32_t r = __shfl_sync(0xffffffffu, grow, src);
        const uint64_t local = (uint64_t)r - args.row_base;
        const float s = warp_exact_dot(args.slab + local * args.dim, q, args.dim, args.reduce_order, args.tail_fma);
        if ((threadIdx.x & 31u) == 0u) {
            const uint64_t exact = make_key(s, r);
            if (exact > *buf.tau && !cand_push(buf, args.buf_cap, exact)) atomicExch(args.error_flag, 1u);
        }
    }
}

// CTA-collective: exact reference scores of the `n` rows listed in `rows` (shared memory, GLOBAL row
// numbers), offered to the bounded top-k buffer.  Eight rows per warp step (warp_exact_dot8) when
// dim % 32 == 0, compaction whenever the buffer could fill.
__device__ __forceinline__ void mma_rescore_list(const MmaRefineArgs& args, const CandBuf& buf, const float* q,
                                                 const uint32_t* rows, uint32_t n) {
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, n_warps = blockDim.x >> 5;
    const uint32_t chunk = (args.buf_cap - args.k) & ~7u;  // pushes the buffer absorbs between compactions
    const bool wide = (args.dim & 31u) == 0u;
    for (uint32_t c0 = 0; c0 < n; c0 += chunk) {
        const uint32_t c1 = min(n, c0 + chunk);
        if (wide) {
            for (uint32_t e0 = c0 + warp * 8u; e0 < c1; e0 += n_warps * 8u) {
                const uint32_t e = e0 + (lane >> 2);
                const bool valid = e < c1;
                const uint32_t grow = rows[valid ? e : c1 - 1u];
                const float sx = warp_exact_dot8(args.slab + ((uint64_t)grow - args.row_base) * args.dim, q, args.dim,
                                                 args.reduce_order);
                if (valid && (lane & 3u) == 0u) {
                    const uint64_t exact = make_key(sx, grow);
                    if (exact > *buf.tau && !cand_push(buf, args.buf_cap, exact)) atomicExch(args.error_flag, 1u);
                }
            }
        } else {
            for (uint32_t e = c0 + warp; e < c1; e += n_warps) {
                const uint32_t grow = rows[e];
                const float sx = warp_exact_dot(args.slab + ((uint64_t)grow - args.row_base) * args.dim, q, args.dim,
                                                args.reduce_order, args.tail_fma);
                if (lane == 0u) {
                    const uint64_t exact = make_key(sx, grow);
                    if (exact > *buf.tau && !cand_push(buf, args.buf_cap, exact)) atomicExch(args.error_flag, 1u);
                }
            }
        }
        __syncthreads();
        if (c1 < n) cand_compact(buf, args.buf_cap, args.k);
    }
}

constexpr uint32_t kMmaTopListCap = 2048;  // dense list of the approximate top-k (ties included) in the refine

__global__ void __launch_bounds__(256, 5) mma_refine_kernel(const MmaRefineArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* tau = cand + args.buf_cap;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(tau + 1);
    __shared__ int s_overflow;
    const uint32_t b = blockIdx.x;
    const uint32_t step = blockDim.x;
    const MmaLists& l = args.lists;
    if (args.redo[b] != 0u) {  // the caller re-runs this query on the exact path
        if (threadIdx.x == 0) {
            if (args.intermediate)
                args.out_counts[b] = 0u;
            else
                atomicOr(args.redo_any, 1u);
        }
        return;
    }
    if (threadIdx.x == 0) {
        *cnt = 0u;
        *tau = 0ull;
        s_overflow = 0;
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < mma_list_count(l); j += step)
        if (l.cand_count[mma_list_slot(l, b, j)] > l.cap) s_overflow = 1;
    __syncthreads();
    if (s_overflow) {  // the superset is incomplete: exact path
        if (threadIdx.x == 0) {
            if (args.intermediate) {
                args.out_counts[b] = 0u;  // no exact gate from this level: the previous gate stays
            } else {
                args.redo[b] = 2u;
                atomicOr(args.redo_any, 1u);
            }
        }
        return;
    }
    const CandBuf buf{cand, cnt, tau};
    const uint32_t trigger = args.buf_cap - step;
    const MmaStageSmem sm = carve_stage_smem(smem_raw + ((size_t)args.buf_cap * 8 + 16), args.stage_cap, true);
    const uint32_t total = mma_stage_lists(sm, l, b);
    const bool staged = total <= sm.cap;

    // pass 1: tau_a = k-th best APPROXIMATE score (radix select); band gate in the ordered domain
    uint32_t gate_u = 0u;  // below every real score: keep everything when fewer than k entries exist
    uint32_t top_u = 0xFFFFFFFFu;  // ordered tau_a (staged lists only: see the two rounds below)
    if (total >= args.k) {
        const uint32_t u = mma_kth_best(sm, l, b, args.k, total, staged);
        gate_u = ordered_score(__fsub_rd(unordered_score(u), args.margin2[b]));
        if (staged) top_u = u;
    }

    // pass 2: every entry inside the band is re-scored exactly and competes on its exact key.  The
    // query is staged in shared memory first: read from global memory, every 32-element step of
    // every exact dot paid an L2 round trip for it.
    float* q = reinterpret_cast<float*>(sm.score + (size_t)sm.cap * 2);
    for (uint32_t i = threadIdx.x; i < args.dim; i += step) q[i] = args.queries[(size_t)b * args.dim + i];
    __syncthreads();
    if (staged) {
        // (a) gather the rows of the band into a dense list (it reuses the score array: every thread
        //     first folds the pass flags of its <= 64 entries into a register), (b) spread the exact
        //     dots evenly over the warps — no barrier per 256 entries, no warp idling while another
        //     re-scores (the first version spent half its time at those barriers).
        // Round 1: the approximate top-k (score >= tau_a) are re-scored first.  The k-th best of
        // their exact scores, tau_x, is a lower bound of the exact k-th best, so a row can only
        // belong to the top-k if approx >= tau_x - e (e = margin2 / 2): a one-sided margin below
        // an exact score, never lower than the two-sided band [tau_a - 2e, inf) it replaces
        // (tau_x >= tau_a - e).  With the int8 bound that is ~4x fewer exact dots in round 2.
        uint32_t lo_u = gate_u;
        if (top_u != 0xFFFFFFFFu) {
            // the approximate top-k as a dense list (normally exactly k rows; a tie band at tau_a can make
            // it longer than the list: then the flattened array is walked with a barrier per 256 entries)
            uint32_t* top_list = reinterpret_cast<uint32_t*>(q + args.dim);
            if (threadIdx.x == 0) sm.ctl[2] = 0u;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < total; i += step)
                if (sm.score[i] >= top_u) {
                    const uint32_t slot = atomicAdd(&sm.ctl[2], 1u);
                    if (slot < kMmaTopListCap) top_list[slot] = sm.row[i];
                }
            __syncthreads();
            const uint32_t n_top = sm.ctl[2];
            if (n_top <= kMmaTopListCap) {
                mma_rescore_list(args, buf, q, top_list, n_top);
            } else {
                for (uint32_t i0 = 0; i0 < total; i0 += step) {  // CTA-uniform trip count
                    const uint32_t i = i0 + threadIdx.x;
                    const bool pass = i < total && sm.score[i] >= top_u;
                    mma_rescore_round(args, buf, q, pass, pass ? sm.row[i] : 0u);
                    __syncthreads();
                    if (*cnt > trigger) cand_compact(buf, args.buf_cap, args.k);
                    __syncthreads();
                }
            }
            cand_compact(buf, args.buf_cap, args.k);
            if (*cnt >= args.k) {
                const float tau_x = key_score(cand[args.k - 1u]);
                const uint32_t lo2 = ordered_score(__fsub_rd(tau_x, __fmul_ru(args.margin2[b], 0.5f)));
                lo_u = max(lo_u, lo2);
            }
            __syncthreads();
        }
        uint64_t mine = 0;  // <= kMmaStagePairs / 256 = 64 entries per thread
        for (uint32_t r = 0, i = threadIdx.x; i < total; ++r, i += step)
            if (sm.score[i] >= lo_u && sm.score[i] < top_u) mine |= 1ull << r;
        if (threadIdx.x == 0) sm.ctl[2] = 0u;
        __syncthreads();
        uint32_t* band = sm.score;
        for (uint32_t r = 0, i = threadIdx.x; i < total; ++r, i += step)
            if (mine & (1ull << r)) band[atomicAdd(&sm.ctl[2], 1u)] = sm.row[i];
        __syncthreads();
        mma_rescore_list(args, buf, q, band, sm.ctl[2]);
    } else {
        for (uint32_t j = 0; j < mma_list_count(l); ++j) {
            const uint32_t n = sm.offs[j + 1] - sm.offs[j];
            const MmaCand* list = l.cand + mma_list_slot(l, b, j) * l.cap;
            for (uint32_t base = 0; base < n; base += step) {
                const uint32_t i = base + threadIdx.x;
                MmaCand c;
                c.score = 0.0f;
                c.row = 0;
                if (i < n) c = list[i];
                mma_rescore_round(args, buf, q, i < n && ordered_score(c.score) >= gate_u, c.row);
                __syncthreads();
                if (*cnt > trigger) cand_compact(buf, args.buf_cap, args.k);
                __syncthreads();
            }
        }
    }
    cand_compact(buf, args.buf_cap, args.k);
    const uint32_t count = *cnt;
    if (threadIdx.x == 0 && args.out_counts) args.out_counts[b] = count;
    for (uint32_t i = threadIdx.x; i < args.k; i += step) {
        const uint64_t key = i < count ? cand[i] : 0ull;
        if (args.out_keys) args.out_keys[(size_t)b * args.k + i] = key;
        if (args.out_hits) {
            fsgpu_hit_t h;
            h.row = key ? key_row(key) : 0xFFFFFFFFu;
            h.score = key ? key_score(key) : 0.0f;
            args.out_hits[(size_t)b * args.k + i] = h;
        }
    }
}

}  // namespace fsgpu
