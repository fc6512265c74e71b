// fusion_kernels.cuh — Reciprocal Rank Fusion and two-tier blend on the device
// (SURVEY.md §8a rows a14, a15, a17).  One CTA per query; everything lives in shared memory.
//
// Reference: crates/frankensearch-fusion/src/rrf.rs:118-121, :179-198, :1038-1210 and
// crates/frankensearch-fusion/src/blend.rs:35-77, :107-191, :213-286.
#pragma once

#include "fsgpu_common.cuh"

namespace fsgpu {

constexpr int kFusionThreads = 256;       // lists up to kFusionWideFrom entries
constexpr int kFusionWideThreads = 1024;  // longer lists: the sorts are the whole cost (0.88 ms for 6000
                                          // entries with 256 threads, profiles/r02_launches_select.md)
constexpr uint32_t kFusionWideFrom = 1024;
constexpr uint32_t kFusionMaxEntries = 8192;  // n_sem_max + n_lex_max (160 KB of shared memory)
constexpr uint32_t kTieBits = 18;             // tie ranks are ranks inside one candidate union

// Ascending u64 image of f64::total_cmp.
__device__ __forceinline__ uint64_t ordered_f64(double d) {
    const uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u ^ 0x8000000000000000ull);
}
// Ascending u32 image of f32::total_cmp WITHOUT the NaN fold (rrf/blend compare raw scores).
__device__ __forceinline__ uint32_t ordered_f32_raw(float f) {
    const uint32_t u = __float_as_uint(f);
    return (u >> 31) ? ~u : (u ^ 0x80000000u);
}
__device__ __forceinline__ float unordered_f32_raw(uint32_t o) {
    return __uint_as_float((o >> 31) ? (o ^ 0x80000000u) : ~o);
}

// Bitonic sort ascending over (hi, lo) with a u32 payload; n a power of two.  Pair-indexed (no idle
// threads); stages with j <= 16 stay inside the 64 entries a warp owns and use a warp barrier
// (fsgpu_common.cuh, cta_sort_desc).  KEYS_ONLY sorts `hi` alone (the join passes carry nothing else).
template <bool KEYS_ONLY>
__device__ __forceinline__ void cta_sort_asc_impl(uint64_t* hi, uint64_t* lo, uint32_t* pay, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        if ((k >> 1) > 16u) __syncthreads();
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const uint32_t i = ((t & ~(j - 1u)) << 1) | (t & (j - 1u));
                const uint32_t ixj = i | j;
                const uint64_t ah = hi[i], bh = hi[ixj];
                const bool asc = (i & k) == 0;
                if constexpr (KEYS_ONLY) {
                    if (asc ? (ah > bh) : (ah < bh)) {
                        hi[i] = bh;
                        hi[ixj] = ah;
                    }
                } else {
                    const uint64_t al = lo[i], bl = lo[ixj];
                    const bool a_gt_b = ah > bh || (ah == bh && al > bl);
                    const bool a_lt_b = ah < bh || (ah == bh && al < bl);
                    if (asc ? a_gt_b : a_lt_b) {
                        hi[i] = bh; hi[ixj] = ah;
                        lo[i] = bl; lo[ixj] = al;
                        const uint32_t p = pay[i];
                        pay[i] = pay[ixj];
                        pay[ixj] = p;
                    }
                }
            }
            if (j > 16u) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void cta_sort_asc_pairs(uint64_t* hi, uint64_t* lo, uint32_t* pay, uint32_t n) {
    cta_sort_asc_impl<false>(hi, lo, pay, n);
}
__device__ __forceinline__ void cta_sort_asc_keys(uint64_t* hi, uint32_t n) {
    cta_sort_asc_impl<true>(hi, nullptr, nullptr, n);
}

// rrf.rs:118-121 — 1.0 / (k + f64::from(rank as u32) + 1.0), each op one IEEE rounding.
__device__ __forceinline__ double rank_contribution(double k, uint32_t rank) {
    return __ddiv_rn(1.0, __dadd_rn(__dadd_rn(k, (double)rank), 1.0));
}

struct RrfArgs {
    double k, w_lex, w_sem;   // already sanitised on the host (rrf.rs:92-98, :124-130)
    int tiebreak;             // 0 LexicalThenId, 1 Hash
    const uint64_t* lex_ids;  // [batch, n_lex_max]   ids < 2^40
    const float* lex_scores;
    const uint32_t* lex_tie;  // nullable
    const uint32_t* lex_counts;
    uint32_t n_lex_max;
    const fsgpu_hit_t* sem_hits;  // [batch, n_sem_max] (row, score) — either this ...
    const uint32_t* sem_rows;     // ... or split arrays
    const float* sem_scores;
    const uint32_t* sem_tie;  // nullable
    const uint32_t* sem_counts;
    uint32_t n_sem_max;
    uint32_t limit, offset;
    fsgpu_fused_hit_t* out;   // [batch, limit]
    uint32_t* out_counts;
};

constexpr uint32_t kNone16 = 0xFFFFu;

template <int THREADS>
__global__ void __launch_bounds__(THREADS) rrf_fuse_kernel(const RrfArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t b = blockIdx.x;
    const uint32_t n_lex = min(args.lex_counts ? args.lex_counts[b] : args.n_lex_max, args.n_lex_max);
    const uint32_t n_sem = min(args.sem_counts ? args.sem_counts[b] : args.n_sem_max, args.n_sem_max);
    const uint32_t m = n_lex + n_sem;
    const uint32_t m2 = next_pow2(max(m, 1u));
    uint64_t* hi = reinterpret_cast<uint64_t*>(smem_raw);  // [m2]
    uint64_t* lo = hi + m2;                                  // [m2]
    uint32_t* pay = reinterpret_cast<uint32_t*>(lo + m2);   // [m2]
    __shared__ uint32_t n_rec;

    const uint64_t* lex_ids = args.lex_ids + (size_t)b * args.n_lex_max;
    const float* lex_scores = args.lex_scores + (size_t)b * args.n_lex_max;
    const uint32_t* lex_tie = args.lex_tie ? args.lex_tie + (size_t)b * args.n_lex_max : nullptr;
    const uint32_t* sem_tie = args.sem_tie ? args.sem_tie + (size_t)b * args.n_sem_max : nullptr;
    auto sem_row = [&](uint32_t r) -> uint32_t {
        return args.sem_hits ? args.sem_hits[(size_t)b * args.n_sem_max + r].row
                             : args.sem_rows[(size_t)b * args.n_sem_max + r];
    };
    auto sem_score = [&](uint32_t r) -> float {
        return args.sem_hits ? args.sem_hits[(size_t)b * args.n_sem_max + r].score
                             : args.sem_scores[(size_t)b * args.n_sem_max + r];
    };

    // 1) join by id: sort (id, source, rank); lexical entries of an id come first, in rank order.
    for (uint32_t i = threadIdx.x; i < m2; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < n_lex)
            key = (lex_ids[i] << 24) | (0ull << 23) | i;
        else if (i < m)
            key = ((uint64_t)sem_row(i - n_lex) << 24) | (1ull << 23) | (i - n_lex);
        hi[i] = key;
        lo[i] = 0;
        pay[i] = 0;
    }
    if (threadIdx.x == 0) n_rec = 0;
    __syncthreads();
    cta_sort_asc_keys(hi, m2);  // the join only orders the keys

    // 2) one record per distinct id: first semantic occurrence (+ first lexical occurrence), or
    //    a lexical-only record.  Records are staged in registers, written after a barrier
    //    because they overwrite the join keys.
    constexpr uint32_t kPerThread = (THREADS == kFusionThreads ? kFusionWideFrom : kFusionMaxEntries) / THREADS;
    uint64_t rec_hi[kPerThread], rec_lo[kPerThread];
    uint32_t rec_pay[kPerThread];
    uint32_t n_mine = 0;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const uint64_t key = hi[i];
        const uint64_t id = key >> 24;
        const bool is_sem = (key >> 23) & 1;
        const uint32_t rank = (uint32_t)(key & 0x7FFFFFu);
        const bool head = i == 0 || (hi[i - 1] >> 24) != id;
        uint32_t sem_rank = kNone16, lex_rank = kNone16;
        bool emit = false;
        if (is_sem) {
            const bool first_sem = head || !((hi[i - 1] >> 23) & 1);
            if (first_sem) {  // semantic dedup keeps the first occurrence (rrf.rs:1083-1089)
                emit = true;
                sem_rank = rank;
                if (!head) {  // walk back to the group head = first lexical occurrence
                    uint32_t h = i - 1;
                    while (h > 0 && (hi[h - 1] >> 24) == id) --h;
                    lex_rank = (uint32_t)(hi[h] & 0x7FFFFFu);
                }
            }
        } else if (head) {  // lexical head: lexical-only unless a semantic entry follows
            uint32_t f = i + 1;
            while (f < m && (hi[f] >> 24) == id && !((hi[f] >> 23) & 1)) ++f;
            const bool has_sem = f < m && (hi[f] >> 24) == id;
            if (!has_sem) {
                emit = true;
                lex_rank = rank;
            }
        }
        if (emit) {
            double score;
            if (sem_rank != kNone16) {  // rrf.rs:1096-1099
                score = __dmul_rn(rank_contribution(args.k, sem_rank), args.w_sem);
                if (lex_rank != kNone16)
                    score = __dadd_rn(score, __dmul_rn(rank_contribution(args.k, lex_rank), args.w_lex));
            } else {                    // rrf.rs:1128
                score = __dmul_rn(rank_contribution(args.k, lex_rank), args.w_lex);
            }
            const bool in_both = sem_rank != kNone16 && lex_rank != kNone16;
            const float lex_s = lex_rank != kNone16 ? lex_scores[lex_rank] : -INFINITY;
            uint32_t tie;
            if (sem_rank != kNone16)
                tie = sem_tie ? sem_tie[sem_rank] : sem_rank;
            else
                tie = lex_tie ? lex_tie[lex_rank] : n_sem + lex_rank;
            tie &= (1u << kTieBits) - 1u;
            const uint64_t lex_word =
                args.tiebreak == 0 ? (uint64_t)(~ordered_f32_raw(lex_s)) : 0ull;  // rrf.rs:185-196
            rec_hi[n_mine] = ~ordered_f64(score);  // rrf desc
            rec_lo[n_mine] = ((uint64_t)(in_both ? 0 : 1) << 63) | (lex_word << 31) |
                             ((uint64_t)tie << 13);
            rec_pay[n_mine] = (sem_rank << 16) | lex_rank;
            ++n_mine;
        }
    }
    __syncthreads();
    const uint32_t slot0 = n_mine ? atomicAdd(&n_rec, n_mine) : 0;
    __syncthreads();
    const uint32_t total = n_rec;
    const uint32_t t2 = next_pow2(max(total, 1u));
    for (uint32_t i = threadIdx.x; i < t2; i += blockDim.x) {
        hi[i] = ~0ull;
        lo[i] = ~0ull;
    }
    __syncthreads();
    for (uint32_t i = 0; i < n_mine; ++i) {
        hi[slot0 + i] = rec_hi[i];
        lo[slot0 + i] = rec_lo[i];
        pay[slot0 + i] = rec_pay[i];
    }
    __syncthreads();
    // 3) rank: (rrf desc, in_both first, lexical score desc, tie asc)  (rrf.rs:179-198)
    cta_sort_asc_pairs(hi, lo, pay, t2);

    const uint32_t window_end = min(total, args.offset + args.limit);
    const uint32_t n_out = window_end > args.offset ? window_end - args.offset : 0;
    for (uint32_t i = threadIdx.x; i < n_out; i += blockDim.x) {
        const uint32_t s = args.offset + i;
        const uint32_t sem_rank = pay[s] >> 16, lex_rank = pay[s] & 0xFFFFu;
        fsgpu_fused_hit_t h;
        h.rrf_score = __longlong_as_double((long long)([&] {
            const uint64_t o = ~hi[s];
            return (o >> 63) ? (o ^ 0x8000000000000000ull) : ~o;
        }()));
        h.semantic_rank = sem_rank == kNone16 ? -1 : (int32_t)sem_rank;
        h.lexical_rank = lex_rank == kNone16 ? -1 : (int32_t)lex_rank;
        h.semantic_row = sem_rank == kNone16 ? 0xFFFFFFFFu : sem_row(sem_rank);
        h.semantic_score = sem_rank == kNone16 ? 0.0f : sem_score(sem_rank);
        h.lexical_score = lex_rank == kNone16 ? 0.0f : lex_scores[lex_rank];
        h.in_both_sources = (sem_rank != kNone16 && lex_rank != kNone16) ? 1u : 0u;
        args.out[(size_t)b * args.limit + i] = h;
    }
    if (threadIdx.x == 0 && args.out_counts) args.out_counts[b] = n_out;
}

// ─── blend ──────────────────────────────────────────────────────────────────────────────────
// One CTA per query.  Every list is [batch, n_*_max] row-major; `*_counts[b]` (nullable = full) is
// the filled prefix of query b's list.  Rows and scores come either from fsgpu_hit records
// (`fast_hits`, `quality_hits`: what the device searches emit) or from split arrays.
struct BlendArgs {
    float alpha;                 // sanitised (blend.rs:518-524)
    const fsgpu_hit_t* fast_hits;   // [batch, n_fast_max] — either this ...
    const uint32_t* fast_rows;      // ... or split arrays
    const float* fast_scores;
    const uint32_t* fast_tie;    // nullable
    const uint32_t* fast_counts; // nullable
    uint32_t n_fast_max;
    uint32_t union_form;              // 1: a separately retrieved quality list, joined by row
    const fsgpu_hit_t* quality_hits;  // union form: [batch, n_quality_max] — either this ...
    const uint32_t* quality_rows;     // ... or split arrays
    const float* quality_scores;      // aligned form: [batch, n_fast_max], slot i belongs to fast hit i
    const uint8_t* quality_present;   // aligned form (nullable = all present)
    const uint32_t* quality_tie;      // union form, nullable
    const uint32_t* quality_counts;   // union form, nullable
    uint32_t n_quality_max;
    fsgpu_hit_t* out;            // [batch, out_stride]
    uint32_t out_stride;
    uint32_t* out_counts;        // [batch]
};

__device__ __forceinline__ float block_min(float v, float* scratch, bool is_max) {
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, w) : fminf(v, w);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = scratch[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = is_max ? fmaxf(r, scratch[w]) : fminf(r, scratch[w]);
    __syncthreads();
    return r;
}

struct NormBounds {  // blend.rs:35-77
    float min, range;
    bool saw_finite;
    __device__ __forceinline__ float apply(float s) const {
        if (!saw_finite || !isfinite(s)) return 0.0f;
        float v = range > 1.1920929e-07f ? __fdiv_rn(__fsub_rn(s, min), range) : 1.0f;
        if (v < 0.0f) v = 0.0f;
        if (v > 1.0f) v = 1.0f;
        return v;
    }
};

template <class ScoreAt>
__device__ __forceinline__ NormBounds fit_bounds(ScoreAt score_at, const uint8_t* present, uint32_t n,
                                                 float* scratch) {
    float mn = INFINITY, mx = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (present && !present[i]) continue;
        const float v = score_at(i);
        if (isfinite(v)) {
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    }
    mn = block_min(mn, scratch, false);
    mx = block_min(mx, scratch, true);
    NormBounds b;
    b.saw_finite = mn <= mx;  // at least one finite value seen
    b.min = mn;
    b.range = __fsub_rn(mx, mn);
    return b;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS) blend_two_tier_kernel(const BlendArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float scratch[THREADS / 32];
    __shared__ uint32_t n_rec;
    const uint32_t b = blockIdx.x;
    const bool union_form = args.union_form != 0;
    const uint32_t n_fast = min(args.fast_counts ? args.fast_counts[b] : args.n_fast_max, args.n_fast_max);
    const uint32_t n_q = union_form ? min(args.quality_counts ? args.quality_counts[b] : args.n_quality_max,
                                          args.n_quality_max)
                                    : 0;
    const uint32_t m = n_fast + n_q;
    // the carve-up must not depend on the query: sized for the widest possible list
    const uint32_t m2_max = next_pow2(max(args.n_fast_max + (union_form ? args.n_quality_max : 0), 1u));
    const uint32_t m2 = next_pow2(max(m, 1u));
    uint64_t* hi = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* lo = hi + m2_max;
    uint32_t* pay = reinterpret_cast<uint32_t*>(lo + m2_max);

    const size_t f0 = (size_t)b * args.n_fast_max, q0 = (size_t)b * (union_form ? args.n_quality_max : args.n_fast_max);
    auto fast_row = [&](uint32_t i) -> uint32_t { return args.fast_hits ? args.fast_hits[f0 + i].row : args.fast_rows[f0 + i]; };
    auto fast_score = [&](uint32_t i) -> float { return args.fast_hits ? args.fast_hits[f0 + i].score : args.fast_scores[f0 + i]; };
    auto q_row = [&](uint32_t i) -> uint32_t { return args.quality_hits ? args.quality_hits[q0 + i].row : args.quality_rows[q0 + i]; };
    auto q_score = [&](uint32_t i) -> float {
        return (union_form && args.quality_hits) ? args.quality_hits[q0 + i].score : args.quality_scores[q0 + i];
    };
    const uint8_t* q_present = (!union_form && args.quality_present) ? args.quality_present + q0 : nullptr;
    const uint32_t* fast_tie = args.fast_tie ? args.fast_tie + f0 : nullptr;
    const uint32_t* quality_tie = (union_form && args.quality_tie) ? args.quality_tie + q0 : nullptr;

    const NormBounds fb = fit_bounds(fast_score, nullptr, n_fast, scratch);
    const NormBounds qb = fit_bounds(q_score, q_present, union_form ? n_q : n_fast, scratch);
    // join by row: (row, source, position); fast entries first within a row
    for (uint32_t i = threadIdx.x; i < m2; i += blockDim.x) {
        uint64_t key = ~0ull;
        if (i < n_fast)
            key = ((uint64_t)fast_row(i) << 24) | i;
        else if (i < m)
            key = ((uint64_t)q_row(i - n_fast) << 24) | (1ull << 23) | (i - n_fast);
        hi[i] = key;
        lo[i] = 0;
        pay[i] = 0;
    }
    if (threadIdx.x == 0) n_rec = 0;
    __syncthreads();
    cta_sort_asc_keys(hi, m2);  // the join only orders the keys

    constexpr uint32_t kPerThread = (THREADS == kFusionThreads ? kFusionWideFrom : kFusionMaxEntries) / THREADS;
    uint64_t rec_hi[kPerThread];
    uint32_t rec_pay[kPerThread];
    uint32_t n_mine = 0;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        const uint64_t key = hi[i];
        const uint64_t row = key >> 24;
        const bool head = i == 0 || (hi[i - 1] >> 24) != row;
        if (!head) continue;  // one record per document (first occurrence wins, blend.rs:130-152)
        const bool head_is_quality = (key >> 23) & 1;
        const uint32_t pos = (uint32_t)(key & 0x7FFFFFu);
        bool has_fast = !head_is_quality, has_q = false;
        float f = 0.0f, q = 0.0f;
        uint32_t tie;
        if (has_fast) {
            f = fb.apply(fast_score(pos));
            tie = fast_tie ? fast_tie[pos] : (uint32_t)row;
            if (union_form) {
                uint32_t j = i + 1;
                while (j < m && (hi[j] >> 24) == row && !((hi[j] >> 23) & 1)) ++j;
                if (j < m && (hi[j] >> 24) == row) {
                    has_q = true;
                    q = qb.apply(q_score((uint32_t)(hi[j] & 0x7FFFFFu)));
                }
            } else {
                // aligned: first fast occurrence (in fast order) that carries a quality score
                for (uint32_t j = i; j < m && (hi[j] >> 24) == row; ++j) {
                    const uint32_t p = (uint32_t)(hi[j] & 0x7FFFFFu);
                    if (!q_present || q_present[p]) {
                        has_q = true;
                        q = qb.apply(q_score(p));
                        break;
                    }
                }
            }
        } else {
            has_q = true;
            q = qb.apply(q_score(pos));
            tie = quality_tie ? quality_tie[pos] : (uint32_t)row;
        }
        float s;
        if (has_fast && has_q)  // alpha.mul_add(q, (1 - alpha) * f)   blend.rs:256-261
            s = __fmaf_rn(args.alpha, q, __fmul_rn(__fsub_rn(1.0f, args.alpha), f));
        else
            s = has_fast ? f : q;
        if (!isfinite(s)) s = 0.0f;
        rec_hi[n_mine] = ((uint64_t)(~ordered_f32_raw(s)) << 32) | tie;
        rec_pay[n_mine] = (uint32_t)row;
        ++n_mine;
    }
    __syncthreads();
    const uint32_t slot0 = n_mine ? atomicAdd(&n_rec, n_mine) : 0;
    __syncthreads();
    const uint32_t total = n_rec;
    const uint32_t t2 = next_pow2(max(total, 1u));
    for (uint32_t i = threadIdx.x; i < t2; i += blockDim.x) {
        hi[i] = ~0ull;
        lo[i] = 0;
    }
    __syncthreads();
    for (uint32_t i = 0; i < n_mine; ++i) {
        hi[slot0 + i] = rec_hi[i];
        pay[slot0 + i] = rec_pay[i];
    }
    __syncthreads();
    cta_sort_asc_pairs(hi, lo, pay, t2);  // score desc (total_cmp), then tie asc (blend.rs:272-276)
    fsgpu_hit_t* out = args.out + (size_t)b * args.out_stride;
    for (uint32_t i = threadIdx.x; i < args.out_stride; i += blockDim.x) {
        fsgpu_hit_t h;
        h.row = i < total ? pay[i] : 0xFFFFFFFFu;
        h.score = i < total ? unordered_f32_raw(~(uint32_t)(hi[i] >> 32)) : 0.0f;
        out[i] = h;
    }
    if (threadIdx.x == 0) args.out_counts[b] = total;
}

// ─── potion / Model2Vec ─────────────────────────────────────────────────────────────────────
// model2vec_embedder.rs:312-335, :435-451 and embed/src/simd.rs:74-116.  One CTA per query:
// thread d owns output dimension d and adds the gathered rows in token order (coalesced across
// d); the squared norm is accumulated sequentially over d exactly like the reference loop.
__global__ void __launch_bounds__(256)
potion_embed_kernel(const float* __restrict__ table, uint64_t vocab, uint32_t dim,
                    const uint32_t* __restrict__ ids, const uint64_t* __restrict__ offsets,
                    float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* v = reinterpret_cast<float*>(smem_raw);  // [dim]
    __shared__ float inv_norm_s;
    __shared__ uint32_t count_s;
    const uint32_t b = blockIdx.x;
    const uint64_t t0 = offsets[b], t1 = offsets[b + 1];
    if (threadIdx.x == 0) {
        uint32_t c = 0;
        for (uint64_t t = t0; t < t1; ++t) c += (uint64_t)ids[t] < vocab ? 1u : 0u;
        count_s = c;
    }
    __syncthreads();
    const uint32_t count = count_s;
    const float inv = count ? __fdiv_rn(1.0f, (float)count) : 0.0f;
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x) {
        float s = 0.0f;
        for (uint64_t t = t0; t < t1; ++t) {
            const uint64_t id = ids[t];
            if (id < vocab) s = add_rn(s, __ldg(table + id * dim + d));
        }
        v[d] = mul_rn(s, inv);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float norm_sq = 0.0f;
        for (uint32_t d = 0; d < dim; ++d) norm_sq = add_rn(norm_sq, mul_rn(v[d], v[d]));
        const bool ok = count > 0 && isfinite(norm_sq) && norm_sq > 1.1920929e-07f;
        inv_norm_s = ok ? __fdiv_rn(1.0f, __fsqrt_rn(norm_sq)) : 0.0f;
    }
    __syncthreads();
    const float inv_norm = inv_norm_s;
    for (uint32_t d = threadIdx.x; d < dim; d += blockDim.x)
        out[(size_t)b * dim + d] = inv_norm == 0.0f ? 0.0f : mul_rn(v[d], inv_norm);
}

}  // namespace fsgpu
