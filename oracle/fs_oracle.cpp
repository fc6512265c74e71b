// fs_oracle.cpp — CPU restatement of frankensearch's semantic-tier hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library, and only as the checker / the CPU arm.  The product (frankensearch_b200/) never
// links, imports or executes it.
//
// Parity pinning: the reference is a Rust workspace and cannot be compiled in this image (no
// cargo/rustc), so this restatement is pinned against the reference's own known-answer tests
// and literal fixtures (see tests/test_oracle_golden.py, each case cites the reference test it
// replays).  One detail is NOT pinnable from the sources on disk: the lane order of
// `wide::f32x8::reduce_add` (wide 1.6.1 is an un-vendored crates.io dependency,
// /root/reference/Cargo.lock:6062-6065).  It is a run-time switch here (`reduce_order`); ids are
// insensitive to it except across exact f32 near-ties, scores move by <= 3 f32 roundings.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference/).
// Build: see oracle/Makefile (-O3 -mavx2 -mf16c -mfma -ffp-contract=off; no fast-math).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#if defined(__AVX2__) && defined(__F16C__)
#include <immintrin.h>
#define FSO_HAVE_AVX2 1
#else
#define FSO_HAVE_AVX2 0
#endif

#define FSO_API extern "C" __attribute__((visibility("default")))

namespace {

// ───────────────────────────── f16 <-> f32 ─────────────────────────────────────────────────
// crates/frankensearch-index/src/simd.rs:67-81 (widen8_f16_lanes): bit-exact with
// half::f16::to_f32 for every one of the 65 536 patterns (test simd_f16_widen_is_bit_exact,
// simd.rs:2711).  Written here as the textbook IEEE widening, not the magic-multiply form.
inline float f16_bits_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t exp = (h >> 10) & 0x1fu;
    uint32_t mant = h & 0x3ffu;
    uint32_t out;
    if (exp == 0) {
        if (mant == 0) {
            out = sign;  // +-0
        } else {
            // subnormal: value = mant * 2^-24; normalise
            int e = -1;
            do {
                mant <<= 1;
                ++e;
            } while ((mant & 0x400u) == 0);
            out = sign | (uint32_t)(127 - 15 - e) << 23 | (mant & 0x3ffu) << 13;
        }
    } else if (exp == 0x1f) {
        out = sign | 0x7f800000u | mant << 13;  // inf / nan (payload preserved)
    } else {
        out = sign | (exp + 112u) << 23 | mant << 13;
    }
    float f;
    std::memcpy(&f, &out, 4);
    return f;
}

// crates/frankensearch-index/src/simd.rs:2245-2304 (encode_f32_to_f16_extend): IEEE
// round-to-nearest-even, identical to `half::f16::from_f32` / `vcvtps2ph imm=RN`.
inline uint16_t f32_to_f16_bits_rne(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    const uint32_t absx = x & 0x7fffffffu;
    if (absx >= 0x7f800000u) {  // inf / nan
        if (absx == 0x7f800000u) return (uint16_t)(sign | 0x7c00u);
        // NaN: keep top mantissa bits, force quiet bit (matches vcvtps2ph and half)
        return (uint16_t)(sign | 0x7c00u | 0x0200u | ((absx >> 13) & 0x3ffu));
    }
    if (absx >= 0x477ff000u) {  // >= 65520 rounds to inf
        return (uint16_t)(sign | 0x7c00u);
    }
    if (absx < 0x33000001u) {  // <= 2^-25 rounds to zero (tie at exactly 2^-25 -> even = 0)
        return (uint16_t)sign;
    }
    const int32_t e = (int32_t)(absx >> 23) - 127;
    uint32_t mant = (absx & 0x7fffffu) | 0x800000u;
    if (e < -14) {
        // subnormal half: shift so that the result unit is 2^-24
        const int shift = (-14 - e) + 13;  // 14..24
        const uint32_t half_val = mant >> shift;
        const uint32_t rem = mant & ((1u << shift) - 1u);
        const uint32_t halfway = 1u << (shift - 1);
        uint32_t r = half_val;
        if (rem > halfway || (rem == halfway && (half_val & 1u))) ++r;
        return (uint16_t)(sign | r);
    }
    uint32_t half_exp = (uint32_t)(e + 15);
    uint32_t half_mant = (mant >> 13) & 0x3ffu;
    const uint32_t rem = mant & 0x1fffu;
    uint32_t r = (half_exp << 10) | half_mant;
    if (rem > 0x1000u || (rem == 0x1000u && (r & 1u))) ++r;  // carry may bump exponent: correct
    return (uint16_t)(sign | r);
}

// ───────────────────────────── dot product ────────────────────────────────────────────────
// reduce_order: lane order of wide::f32x8::reduce_add (UNVERIFIED, see header).
//   0 HALVES_PAIRWISE   ((v0+v1)+(v2+v3)) + ((v4+v5)+(v6+v7))   two f32x4 halves, SSE pairwise
//   1 AVX_TREE          ((v0+v4)+(v2+v6)) + ((v1+v5)+(v3+v7))   extract-high / movehl / shuffle
//   2 HALVES_SEQUENTIAL (((v0+v1)+v2)+v3) + (((v4+v5)+v6)+v7)   two halves, array sum
//   3 HALVES_STRIDE2    ((v0+v2)+(v1+v3)) + ((v4+v6)+(v5+v7))   two halves, movehl then shuffle
//   4 SEQUENTIAL        ((((((v0+v1)+v2)+v3)+v4)+v5)+v6)+v7
inline float reduce8(const float v[8], int order) {
    switch (order) {
        case 1: return ((v[0] + v[4]) + (v[2] + v[6])) + ((v[1] + v[5]) + (v[3] + v[7]));
        case 2: return (((v[0] + v[1]) + v[2]) + v[3]) + (((v[4] + v[5]) + v[6]) + v[7]);
        case 3: return ((v[0] + v[2]) + (v[1] + v[3])) + ((v[4] + v[6]) + (v[5] + v[7]));
        case 4: return ((((((v[0] + v[1]) + v[2]) + v[3]) + v[4]) + v[5]) + v[6]) + v[7];
        default: return ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
    }
}

// Scalar restatement of dot_product_f16_bytes_f32_generic / _avx2
// (crates/frankensearch-index/src/simd.rs:398-446, :532-571) and of the slice twin
// dot_product_f16_f32 (simd.rs:255-302, :308-339).  Four 8-lane accumulators, chunk c goes to
// accumulator c%4 while whole groups of four remain, left-over chunks all go to accumulator 0,
// product and sum are SEPARATE roundings (this TU is built with -ffp-contract=off),
// (s0+s1)+(s2+s3), 8-lane horizontal add, scalar tail.  tail_fma=1: bytes kernel tail
// (`val.mul_add(q, result)`, simd.rs:440-444); tail_fma=0: slice kernel tail
// (`result += val*q`, simd.rs:298-300).
float dot_f16_f32_scalar(const uint16_t* row, const float* q, uint32_t dim, int reduce_order,
                         int tail_fma) {
    const uint32_t chunks = dim / 8;
    float s[4][8];
    for (auto& a : s)
        for (float& l : a) l = 0.0f;
    uint32_t c = 0;
    while (c + 4 <= chunks) {
        for (uint32_t a = 0; a < 4; ++a)
            for (uint32_t l = 0; l < 8; ++l) {
                const uint32_t e = (c + a) * 8 + l;
                const float p = f16_bits_to_f32(row[e]) * q[e];
                s[a][l] = s[a][l] + p;
            }
        c += 4;
    }
    while (c < chunks) {
        for (uint32_t l = 0; l < 8; ++l) {
            const uint32_t e = c * 8 + l;
            const float p = f16_bits_to_f32(row[e]) * q[e];
            s[0][l] = s[0][l] + p;
        }
        ++c;
    }
    float v[8];
    for (uint32_t l = 0; l < 8; ++l) v[l] = (s[0][l] + s[1][l]) + (s[2][l] + s[3][l]);
    float result = reduce8(v, reduce_order);
    for (uint32_t e = chunks * 8; e < dim; ++e) {
        const float val = f16_bits_to_f32(row[e]);
        if (tail_fma)
            result = std::fmaf(val, q[e], result);
        else
            result = result + val * q[e];
    }
    return result;
}

#if FSO_HAVE_AVX2
// Same arithmetic with 256-bit registers for the CPU baseline timing (vcvtph2ps + vmulps +
// vaddps, never vfmadd) — the shape of simd.rs:398-446.  Bit-identical to the scalar form
// above (checked in tests/test_oracle_golden.py::test_avx2_dot_matches_scalar).
float dot_f16_f32_avx2(const uint16_t* row, const float* q, uint32_t dim, int reduce_order,
                       int tail_fma) {
    const uint32_t chunks = dim / 8;
    __m256 s0 = _mm256_setzero_ps(), s1 = s0, s2 = s0, s3 = s0;
    auto prod = [&](uint32_t c) {
        const __m128i h = _mm_loadu_si128(reinterpret_cast<const __m128i*>(row + c * 8));
        return _mm256_mul_ps(_mm256_cvtph_ps(h), _mm256_loadu_ps(q + c * 8));
    };
    uint32_t c = 0;
    for (; c + 4 <= chunks; c += 4) {
        s0 = _mm256_add_ps(s0, prod(c));
        s1 = _mm256_add_ps(s1, prod(c + 1));
        s2 = _mm256_add_ps(s2, prod(c + 2));
        s3 = _mm256_add_ps(s3, prod(c + 3));
    }
    for (; c < chunks; ++c) s0 = _mm256_add_ps(s0, prod(c));
    alignas(32) float v[8];
    _mm256_store_ps(v, _mm256_add_ps(_mm256_add_ps(s0, s1), _mm256_add_ps(s2, s3)));
    float result = reduce8(v, reduce_order);
    for (uint32_t e = chunks * 8; e < dim; ++e) {
        const float val = f16_bits_to_f32(row[e]);
        if (tail_fma)
            result = std::fmaf(val, q[e], result);
        else
            result = result + val * q[e];
    }
    return result;
}
#endif

inline float dot_fast(const uint16_t* row, const float* q, uint32_t dim, int order, int tail_fma) {
#if FSO_HAVE_AVX2
    return dot_f16_f32_avx2(row, q, dim, order, tail_fma);
#else
    return dot_f16_f32_scalar(row, q, dim, order, tail_fma);
#endif
}

// dot_product_f32_f32 (crates/frankensearch-index/src/simd.rs:161-222 AVX2, :1559-1587 generic) —
// the score of a resident WAL row (f32 embedding).  Same four accumulators over whole groups of
// 32, but the left-over 8-chunks are added AFTER `(acc0+acc1)+(acc2+acc3)`, and the scalar tail is
// always `result += a*b` (mul then add).
float dot_f32_f32_scalar(const float* a, const float* b, uint32_t dim, int reduce_order) {
    const uint32_t chunks = dim / 8, groups = dim / 32;
    float s[4][8];
    for (auto& acc : s)
        for (float& l : acc) l = 0.0f;
    for (uint32_t g = 0; g < groups; ++g)
        for (uint32_t acc = 0; acc < 4; ++acc)
            for (uint32_t l = 0; l < 8; ++l) {
                const uint32_t e = g * 32 + acc * 8 + l;
                const float p = a[e] * b[e];
                s[acc][l] = s[acc][l] + p;
            }
    float v[8];
    for (uint32_t l = 0; l < 8; ++l) v[l] = (s[0][l] + s[1][l]) + (s[2][l] + s[3][l]);
    for (uint32_t c = groups * 4; c < chunks; ++c)
        for (uint32_t l = 0; l < 8; ++l) {
            const float p = a[c * 8 + l] * b[c * 8 + l];
            v[l] = v[l] + p;
        }
    float result = reduce8(v, reduce_order);
    for (uint32_t e = chunks * 8; e < dim; ++e) {
        const float p = a[e] * b[e];
        result = result + p;
    }
    return result;
}

// ───────────────────────────── top-k ordering ─────────────────────────────────────────────
// crates/frankensearch-index/src/search.rs:1655-1661 (score_key), :91-126 (HeapEntry::cmp),
// :1673-1686 (compare_best_first, candidate_is_better).
inline float score_key(float s) { return std::isnan(s) ? -INFINITY : s; }

// f32::total_cmp as an integer compare (core::f32::total_cmp's own bit trick).
inline int32_t total_order_key(float f) {
    int32_t b;
    std::memcpy(&b, &f, 4);
    b ^= (int32_t)((uint32_t)(b >> 31) >> 1);
    return b;
}
inline int total_cmp(float a, float b) {
    const int32_t x = total_order_key(a), y = total_order_key(b);
    return x < y ? -1 : (x > y ? 1 : 0);
}

struct Entry {
    uint64_t index;
    float score;
};

// true when `l` ranks strictly before `r` (search.rs:1673-1678 compare_best_first == Less)
inline bool best_first_less(const Entry& l, const Entry& r) {
    const int c = total_cmp(score_key(r.score), score_key(l.score));
    if (c != 0) return c < 0;
    return l.index < r.index;
}
// heap comparator: "largest" == worst so front() is the cutoff (search.rs:112-123)
struct WorstOnTop {
    bool operator()(const Entry& a, const Entry& b) const { return best_first_less(a, b); }
};

// search.rs:1688-1702 insert_candidate
inline void insert_candidate(std::vector<Entry>& heap, const Entry& cand, size_t limit) {
    if (limit == 0) return;
    if (heap.size() < limit) {
        heap.push_back(cand);
        std::push_heap(heap.begin(), heap.end(), WorstOnTop{});
        return;
    }
    const Entry& worst = heap.front();
    if (best_first_less(cand, worst)) {
        std::pop_heap(heap.begin(), heap.end(), WorstOnTop{});
        heap.back() = cand;
        std::push_heap(heap.begin(), heap.end(), WorstOnTop{});
    }
}

inline bool tombstoned(const uint8_t* bitmap, uint64_t row) {
    return bitmap != nullptr && ((bitmap[row >> 3] >> (row & 7)) & 1u) != 0;
}

// search.rs:1257-1327 scan_range_chunk (F16 arm) / in_memory.rs:3490-3528 scan_range.
// The reference reads the tombstone bit from the 16-byte record table (flags & 1); the oracle
// takes the same bit as a packed bitmap (1 bit per row).
void scan_range_chunk(const uint16_t* slab, uint32_t dim, const uint8_t* tomb, uint64_t start,
                      uint64_t end, const float* q, size_t limit, int order, int tail_fma,
                      std::vector<Entry>& heap) {
    heap.clear();
    heap.reserve(std::min<size_t>(limit, end - start) + 1);
    float cutoff = -INFINITY;
    for (uint64_t i = start; i < end; ++i) {
        if (tombstoned(tomb, i)) continue;
        const float score = dot_fast(slab + i * dim, q, dim, order, tail_fma);
        if (heap.size() < limit || score_key(score) >= cutoff) {
            insert_candidate(heap, Entry{i, score}, limit);
            if (heap.size() >= limit) cutoff = score_key(heap.front().score);
        }
    }
}

constexpr uint64_t kChunkRows = 1024;  // search.rs PARALLEL_CHUNK_SIZE (search.rs:1013-1036)

}  // namespace

// ═════════════════════════════ exported C entry points ════════════════════════════════════

FSO_API float fso_f16_to_f32(uint16_t h) { return f16_bits_to_f32(h); }
FSO_API uint16_t fso_f32_to_f16(float f) { return f32_to_f16_bits_rne(f); }

// hw=1 uses vcvtps2ph (the reference's AVX2 arm, simd.rs:2271-2297); hw=0 the portable arm.
FSO_API void fso_encode_f32_to_f16(const float* src, uint64_t n, uint16_t* dst, int hw) {
    uint64_t i = 0;
#if FSO_HAVE_AVX2
    if (hw) {
        for (; i + 8 <= n; i += 8) {
            const __m128i h =
                _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), h);
        }
    }
#else
    (void)hw;
#endif
    for (; i < n; ++i) dst[i] = f32_to_f16_bits_rne(src[i]);
}

FSO_API void fso_decode_f16_to_f32(const uint16_t* src, uint64_t n, float* dst, int hw) {
    uint64_t i = 0;
#if FSO_HAVE_AVX2
    if (hw) {
        for (; i + 8 <= n; i += 8)
            _mm256_storeu_ps(dst + i, _mm256_cvtph_ps(_mm_loadu_si128(
                                          reinterpret_cast<const __m128i*>(src + i))));
    }
#else
    (void)hw;
#endif
    for (; i < n; ++i) dst[i] = f16_bits_to_f32(src[i]);
}

FSO_API int fso_have_avx2(void) { return FSO_HAVE_AVX2; }

// impl: 0 = scalar restatement, 1 = AVX2 restatement (falls back to scalar when not built in)
FSO_API float fso_dot_f16_f32(const uint16_t* row, const float* q, uint32_t dim, int reduce_order,
                              int tail_fma, int impl) {
#if FSO_HAVE_AVX2
    if (impl == 1) return dot_f16_f32_avx2(row, q, dim, reduce_order, tail_fma);
#else
    (void)impl;
#endif
    return dot_f16_f32_scalar(row, q, dim, reduce_order, tail_fma);
}

// VectorIndex::search_top_k_internal / InMemoryVectorIndex::search_top_k_with_params
// (crates/frankensearch-index/src/search.rs:426-494, :1013-1036, :1704-1720, :1493-1500;
//  crates/frankensearch-index/src/in_memory.rs:2651-2710, :3280-3299, :3530-3561).
//   * limit == 0 or n == 0          -> 0 hits                    (search.rs:438-440)
//   * limit >= n                    -> score every live row, sort (search.rs:449-473)
//   * otherwise 1024-row chunks, one bounded heap per chunk, serial merge in chunk order,
//     final best-first sort.  `threads` only changes who scores which chunk.
// Returns the number of hits written (<= limit).  out_rows / out_scores need `min(limit,n)` slots.
FSO_API uint64_t fso_search_top_k(const uint16_t* slab, uint64_t n, uint32_t dim,
                                  const uint8_t* tombstones, const float* query, uint64_t limit,
                                  int threads, int reduce_order, int tail_fma, uint64_t* out_rows,
                                  float* out_scores) {
    if (limit == 0 || n == 0) return 0;
    std::vector<Entry> winners;
    if (limit >= n) {
        winners.resize(n);
        std::atomic<uint64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const uint64_t c = next.fetch_add(1);
                const uint64_t s = c * kChunkRows;
                if (s >= n) break;
                const uint64_t e = std::min(n, s + kChunkRows);
                for (uint64_t i = s; i < e; ++i)
                    winners[i] = Entry{i, dot_fast(slab + i * dim, query, dim, reduce_order, tail_fma)};
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < std::max(1, threads); ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        if (tombstones) {
            std::vector<Entry> live;
            live.reserve(n);
            for (const Entry& e : winners)
                if (!tombstoned(tombstones, e.index)) live.push_back(e);
            winners.swap(live);
        }
    } else {
        const uint64_t n_chunks = (n + kChunkRows - 1) / kChunkRows;
        std::vector<std::vector<Entry>> partial(n_chunks);
        std::atomic<uint64_t> next{0};
        auto work = [&]() {
            for (;;) {
                const uint64_t c = next.fetch_add(1);
                if (c >= n_chunks) break;
                scan_range_chunk(slab, dim, tombstones, c * kChunkRows,
                                 std::min(n, (c + 1) * kChunkRows), query, (size_t)limit,
                                 reduce_order, tail_fma, partial[c]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < std::max(1, threads); ++t) pool.emplace_back(work);
        work();
        for (auto& t : pool) t.join();
        // merge_partial_heaps (search.rs:1704-1720): serial, chunk order
        std::vector<Entry> merged;
        merged.reserve((size_t)limit + 1);
        for (auto& heap : partial)
            for (const Entry& e : heap) insert_candidate(merged, e, (size_t)limit);
        winners.swap(merged);
    }
    std::sort(winners.begin(), winners.end(), best_first_less);
    const uint64_t out_n = std::min<uint64_t>(winners.size(), limit);
    for (uint64_t i = 0; i < out_n; ++i) {
        out_rows[i] = winners[i].index;
        out_scores[i] = winners[i].score;
    }
    return out_n;
}

// ─── quantised two-pass searches ────────────────────────────────────────────────────────────
// quantize_f16_slab_to_i8_generic (crates/frankensearch-index/src/simd.rs:1842-1859): one corpus-wide scale
// 127 / max|x| (f32::max ignores NaN), round half away, clamp, `as i8`.
FSO_API void fso_quantize_slab_i8(const uint16_t* slab, uint64_t n_elems, int8_t* out) {
    float max_abs = 0.0f;
    for (uint64_t i = 0; i < n_elems; ++i) {
        const float a = std::fabs(f16_bits_to_f32(slab[i]));
        if (a > max_abs) max_abs = a;  // NaN compares false: ignored, as f32::max does
    }
    if (max_abs <= 0.0f) {
        memset(out, 0, n_elems);
        return;
    }
    const float scale = 127.0f / max_abs;
    for (uint64_t i = 0; i < n_elems; ++i) {
        const float v = std::fmin(std::fmax(std::round(f16_bits_to_f32(slab[i]) * scale), -127.0f), 127.0f);
        out[i] = (int8_t)(std::isnan(v) ? 0 : (int)v);
    }
}
// pack_f16_slab_to_4bit_generic (simd.rs:2201-2233): scale 7 / max|x| (0 when max|x| <= 1e-9), nibble_of_4bit
// (simd.rs:1892-1896), byte j = dims 2j (low nibble) | 2j + 1 (high nibble), ceil(dim / 2) bytes per vector.
static inline uint8_t nibble_of(float value, float scale) {
    const float v = std::fmin(std::fmax(std::round(value * scale), -7.0f), 7.0f);
    return (uint8_t)((int8_t)(std::isnan(v) ? 0 : (int)v)) & 0x0F;
}
FSO_API void fso_pack_slab_4bit(const uint16_t* slab, uint64_t n_rows, uint32_t dim, uint8_t* out) {
    if (dim == 0) return;
    float max_abs = 0.0f;
    for (uint64_t i = 0; i < n_rows * dim; ++i) {
        const float a = std::fabs(f16_bits_to_f32(slab[i]));
        if (a > max_abs) max_abs = a;
    }
    const float scale = max_abs > 1e-9f ? 7.0f / max_abs : 0.0f;
    const uint32_t bpv = (dim + 1) / 2;
    memset(out, 0, n_rows * bpv);
    for (uint64_t v = 0; v < n_rows; ++v)
        for (uint32_t d = 0; d < dim; ++d) {
            const uint8_t nib = nibble_of(f16_bits_to_f32(slab[v * dim + d]), scale);
            out[v * bpv + d / 2] |= (d % 2 == 0) ? nib : (uint8_t)(nib << 4);
        }
}
// VectorIndex::search_top_k_int8_two_pass / search_top_k_4bit_two_pass (search.rs:571-650, :876-946) on a slab without
// WAL rows: candidate_count = max(min(k * max(mult, 1), n), min(k, n)); pass 1 = integer dot of the codes with the
// query's own codes (quantize_i8_query search.rs:1610-1622, pack_4bit_query :1640-1655; dot_i8_i8 / dot_4bit_prepared are
// exact integers), ranked by (score as f32 under score_key, lower index) — the heap keys of search.rs:141-156 and
// HeapEntry order agree (int8_heap_keys_match_legacy_score_and_index_order, search.rs:3112); pass 2 = exact f16 dot of
// the candidates, top k by the same total order.  Returns the number of hits.
FSO_API uint64_t fso_search_two_pass(const uint16_t* slab, uint64_t n, uint32_t dim, const uint8_t* tombstones,
                                     const float* query, uint64_t k, uint64_t mult, int bits, int reduce_order,
                                     int tail_fma, uint64_t* out_rows, float* out_scores) {
    if (k == 0 || n == 0) return 0;
    const uint64_t want = k * std::max<uint64_t>(mult, 1);
    const uint64_t cand = std::max(std::min(want, n), std::min(k, n));
    float q_max = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) {
        const float a = std::fabs(query[i]);
        if (a > q_max) q_max = a;
    }
    std::vector<int> qc(dim, 0);
    std::vector<Entry> all;
    all.reserve(n);
    if (bits == 8) {
        std::vector<int8_t> codes(n * dim);
        fso_quantize_slab_i8(slab, n * dim, codes.data());
        if (q_max > 0.0f) {
            const float scale = 127.0f / q_max;
            for (uint32_t i = 0; i < dim; ++i) {
                const float v = std::fmin(std::fmax(std::round(query[i] * scale), -127.0f), 127.0f);
                qc[i] = std::isnan(v) ? 0 : (int)v;
            }
        }
        for (uint64_t r = 0; r < n; ++r) {
            if (tombstoned(tombstones, r)) continue;
            int32_t acc = 0;
            for (uint32_t i = 0; i < dim; ++i) acc += (int32_t)codes[r * dim + i] * qc[i];
            all.push_back(Entry{r, (float)acc});
        }
    } else {
        const uint32_t bpv = (dim + 1) / 2;
        std::vector<uint8_t> codes(n * bpv);
        fso_pack_slab_4bit(slab, n, dim, codes.data());
        const float scale = q_max > 1e-9f ? 7.0f / q_max : 0.0f;
        for (uint32_t i = 0; i < dim; ++i) {
            const uint8_t nib = nibble_of(query[i], scale);
            qc[i] = (int)(int8_t)(nib << 4) >> 4;  // sign-extended nibble
        }
        for (uint64_t r = 0; r < n; ++r) {
            if (tombstoned(tombstones, r)) continue;
            int32_t acc = 0;
            for (uint32_t i = 0; i < dim; ++i) {
                const uint8_t b = codes[r * bpv + i / 2];
                const int x = (i % 2 == 0) ? ((int)(int8_t)(b << 4) >> 4) : ((int)(int8_t)(b & 0xF0) >> 4);
                acc += x * qc[i];
            }
            all.push_back(Entry{r, (float)acc});
        }
    }
    std::sort(all.begin(), all.end(), best_first_less);
    if (all.size() > cand) all.resize(cand);
    for (Entry& e : all) e.score = dot_fast(slab + e.index * dim, query, dim, reduce_order, tail_fma);
    std::sort(all.begin(), all.end(), best_first_less);
    const uint64_t out_n = std::min<uint64_t>(all.size(), k);
    for (uint64_t i = 0; i < out_n; ++i) {
        out_rows[i] = all[i].index;
        out_scores[i] = all[i].score;
    }
    return out_n;
}

FSO_API float fso_dot_f32_f32(const float* a, const float* b, uint32_t dim, int reduce_order) {
    return dot_f32_f32_scalar(a, b, dim, reduce_order);
}

// VectorIndex::search_top_k_internal with resident WAL rows (search.rs:426-494): the main-slab heap
// (as fso_search_top_k; `exclude` = tombstones | !filter) then scan_wal (search.rs:1449-1475) into
// the SAME bounded heap — WAL entry w is skipped when the filter rejects it or its score is not
// finite, and enters as index WAL_INDEX_BIT | w (wal.rs:557-569), i.e. after every main row on
// equal scores.  Winners are sorted best-first (resolve_hits, search.rs:1493-1500); a WAL winner is
// written as row n + w (resolve_wal_hit, search.rs:1583-1597).  The doc-id part of
// resolve_sorted_entries (shadowing, dedup) is host logic: oracle/np_oracle.py resolve_sorted_entries.
FSO_API uint64_t fso_search_top_k_wal(const uint16_t* slab, uint64_t n, uint32_t dim, const uint8_t* exclude,
                                      const float* wal, uint64_t n_wal, const uint8_t* wal_allow,
                                      const float* query, uint64_t limit, int threads, int reduce_order,
                                      int tail_fma, uint64_t* out_rows, float* out_scores) {
    if (limit == 0 || (n == 0 && n_wal == 0)) return 0;
    constexpr uint64_t kWalBit = 1ull << 63;
    std::vector<Entry> heap;
    if (n > 0) {
        const uint64_t cap = std::min<uint64_t>(limit, n);
        std::vector<uint64_t> rows(cap);
        std::vector<float> scores(cap);
        // limit >= n + n_wal without a filter is the reference's collect-all path; with the WAL the
        // heap path and the collect-all path agree (full_recall_collect_all_matches_heap_prefix_with_wal,
        // search.rs:2688), so the main part is taken from fso_search_top_k either way
        const uint64_t got = fso_search_top_k(slab, n, dim, exclude, query, limit, threads, reduce_order,
                                              tail_fma, rows.data(), scores.data());
        for (uint64_t i = 0; i < got; ++i) insert_candidate(heap, Entry{rows[i], scores[i]}, (size_t)limit);
    }
    for (uint64_t w = 0; w < n_wal; ++w) {
        if (wal_allow && !((wal_allow[w >> 3] >> (w & 7)) & 1u)) continue;
        const float score = dot_f32_f32_scalar(wal + w * dim, query, dim, reduce_order);
        if (!std::isfinite(score)) continue;
        insert_candidate(heap, Entry{kWalBit | w, score}, (size_t)limit);
    }
    std::sort(heap.begin(), heap.end(), best_first_less);
    for (size_t i = 0; i < heap.size(); ++i) {
        out_rows[i] = (heap[i].index & kWalBit) ? n + (heap[i].index & ~kWalBit) : heap[i].index;
        out_scores[i] = heap[i].score;
    }
    return heap.size();
}

// TwoTierIndex::quality_scores_for_hits -> dot_query_at
// (crates/frankensearch-index/src/two_tier.rs:1566-1631, :1946-1973; lib.rs:3229-3239).
// rows[i] == UINT64_MAX means "no aligned quality row": present[i] = 0.
FSO_API void fso_scores_for_rows(const uint16_t* slab, uint64_t n, uint32_t dim, const float* query,
                                 const uint64_t* rows, uint64_t n_rows, int reduce_order,
                                 int tail_fma, float* out_scores, uint8_t* out_present) {
    for (uint64_t i = 0; i < n_rows; ++i) {
        if (rows[i] >= n) {
            out_scores[i] = 0.0f;
            out_present[i] = 0;
            continue;
        }
        out_scores[i] = dot_fast(slab + rows[i] * dim, query, dim, reduce_order, tail_fma);
        out_present[i] = 1;
    }
}

// ───────────────────────────── FNV-1a ─────────────────────────────────────────────────────
// crates/frankensearch-index/src/lib.rs:6120-6127 (row hash) and
// crates/frankensearch-fusion/src/rrf.rs:68-75 (Hash tie-break) — same function.
FSO_API uint64_t fso_fnv1a64(const uint8_t* bytes, uint64_t len) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (uint64_t i = 0; i < len; ++i) {
        h ^= bytes[i];
        h *= 0x00000100000001b3ull;
    }
    return h;
}

// ───────────────────────────── RRF ────────────────────────────────────────────────────────
namespace {
inline std::string_view sv(const uint8_t* bytes, const uint64_t* off, uint64_t i) {
    return std::string_view(reinterpret_cast<const char*>(bytes) + off[i], off[i + 1] - off[i]);
}
// rrf.rs:118-121
inline double rank_contribution(double k, uint64_t rank) {
    const uint32_t r = rank > 0xffffffffull ? 0xffffffffu : (uint32_t)rank;
    return 1.0 / (k + (double)r + 1.0);
}
inline int total_cmp64(double a, double b) {
    int64_t x, y;
    std::memcpy(&x, &a, 8);
    std::memcpy(&y, &b, 8);
    x ^= (int64_t)((uint64_t)(x >> 63) >> 1);
    y ^= (int64_t)((uint64_t)(y >> 63) >> 1);
    return x < y ? -1 : (x > y ? 1 : 0);
}
struct Fused {
    std::string_view doc_id;
    double rrf;
    int64_t lex_rank, sem_rank;
    int64_t sem_list_pos;  // position in the semantic input (carries index/score), -1 if none
    float lex_score;
    bool has_lex, in_both;
};
}  // namespace

// rrf_fuse / rrf_fuse_for_vector_lane -> rrf_fuse_merge_inner
// (crates/frankensearch-fusion/src/rrf.rs:282-320, :1038-1210; comparator :179-198;
//  sanitisers :92-98, :124-130).  No graph lane (graph_weight = 0 in the path north_star names).
// Doc ids are byte strings: `*_bytes` + `*_off[n+1]`.  tiebreak: 0 LexicalThenId, 1 Hash.
// Outputs (capacity `limit`): position in the semantic list (-1 none), position in the lexical
// list (-1 none; first occurrence), rrf score, in_both flag.  Returns number of fused hits.
FSO_API uint64_t fso_rrf_fuse(const uint8_t* lex_bytes, const uint64_t* lex_off,
                              const float* lex_scores, uint64_t n_lex, const uint8_t* sem_bytes,
                              const uint64_t* sem_off, uint64_t n_sem, double k, double w_lex,
                              double w_sem, int tiebreak, uint64_t limit, uint64_t offset,
                              int dedup_semantic, int64_t* out_sem_pos, int64_t* out_lex_pos,
                              double* out_rrf, uint8_t* out_in_both) {
    if (!(std::isfinite(k) && k >= 0.0)) k = 60.0;                  // rrf.rs:124-130
    if (!(std::isfinite(w_lex) && w_lex > 0.0)) w_lex = 1.0;        // rrf.rs:92-98
    if (!(std::isfinite(w_sem) && w_sem > 0.0)) w_sem = 1.0;

    std::unordered_map<std::string_view, std::pair<uint64_t, float>> lex_map;
    lex_map.reserve(n_lex * 2 + 1);
    for (uint64_t r = 0; r < n_lex; ++r)
        lex_map.emplace(sv(lex_bytes, lex_off, r), std::make_pair(r, lex_scores[r]));  // first wins

    std::vector<Fused> results;
    results.reserve(n_lex + n_sem);
    std::unordered_set<std::string_view> seen;
    for (uint64_t r = 0; r < n_sem; ++r) {
        const std::string_view id = sv(sem_bytes, sem_off, r);
        if (dedup_semantic && !seen.insert(id).second) continue;
        double score = rank_contribution(k, r) * w_sem;             // rrf.rs:1096
        Fused f{id, 0.0, -1, (int64_t)r, (int64_t)r, 0.0f, false, false};
        auto it = lex_map.find(id);
        if (it != lex_map.end()) {
            score += rank_contribution(k, it->second.first) * w_lex;  // rrf.rs:1097-1099
            f.lex_rank = (int64_t)it->second.first;
            f.lex_score = it->second.second;
            f.has_lex = f.in_both = true;
            lex_map.erase(it);
        }
        f.rrf = score;
        results.push_back(f);
    }
    for (auto& kv : lex_map) {                                       // rrf.rs:1121-1146
        Fused f{kv.first, rank_contribution(k, kv.second.first) * w_lex,
                (int64_t)kv.second.first, -1, -1, kv.second.second, true, false};
        results.push_back(f);
    }
    const uint64_t window = limit + offset;
    if (window == 0) return 0;
    auto before = [tiebreak](const Fused& a, const Fused& b) {     // rrf.rs:179-198
        int c = total_cmp64(b.rrf, a.rrf);
        if (c != 0) return c < 0;
        if (a.in_both != b.in_both) return a.in_both;              // true first
        if (tiebreak == 0) {
            const float la = a.has_lex ? a.lex_score : -INFINITY;
            const float lb = b.has_lex ? b.lex_score : -INFINITY;
            c = total_cmp(lb, la);
            if (c != 0) return c < 0;
        } else {
            const uint64_t ha = fso_fnv1a64((const uint8_t*)a.doc_id.data(), a.doc_id.size());
            const uint64_t hb = fso_fnv1a64((const uint8_t*)b.doc_id.data(), b.doc_id.size());
            if (ha != hb) return ha < hb;
        }
        return a.doc_id < b.doc_id;  // byte-wise, like Rust str::cmp
    };
    std::sort(results.begin(), results.end(), before);              // select_nth + sort == sort
    uint64_t out = 0;
    for (uint64_t i = offset; i < results.size() && i < window; ++i, ++out) {
        out_sem_pos[out] = results[i].sem_list_pos;
        out_lex_pos[out] = results[i].lex_rank;
        out_rrf[out] = results[i].rrf;
        out_in_both[out] = results[i].in_both ? 1 : 0;
    }
    return out;
}

// ───────────────────────────── blend ──────────────────────────────────────────────────────
namespace {
struct NormBounds {  // crates/frankensearch-fusion/src/blend.rs:35-77
    float min = INFINITY, range = 0.0f;
    bool saw_finite = false;
    void fit(const float* s, const uint8_t* present, uint64_t n) {
        float mx = -INFINITY;
        for (uint64_t i = 0; i < n; ++i) {
            if (present && !present[i]) continue;
            if (std::isfinite(s[i])) {
                min = s[i] < min ? s[i] : min;
                mx = s[i] > mx ? s[i] : mx;
                saw_finite = true;
            }
        }
        range = mx - min;
    }
    float apply(float s) const {
        if (!saw_finite || !std::isfinite(s)) return 0.0f;
        float v = range > 1.1920929e-07f ? (s - min) / range : 1.0f;
        if (v < 0.0f) v = 0.0f;
        if (v > 1.0f) v = 1.0f;
        return v;
    }
};
inline float sanitize_score(float s) { return std::isfinite(s) ? s : 0.0f; }  // blend.rs:526-532
inline float sanitize_alpha(float a) {                                        // blend.rs:518-524
    if (!std::isfinite(a)) return 0.7f;
    return a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);
}
struct Blended {
    std::string_view doc_id;
    uint32_t index;
    float score;
};
uint64_t emit_blended(std::vector<Blended>& v, uint32_t* out_index, float* out_score,
                      int64_t* out_src, const std::unordered_map<std::string_view, int64_t>& src) {
    std::sort(v.begin(), v.end(), [](const Blended& l, const Blended& r) {  // blend.rs:272-276
        const int c = total_cmp(sanitize_score(r.score), sanitize_score(l.score));
        if (c != 0) return c < 0;
        return l.doc_id < r.doc_id;
    });
    for (uint64_t i = 0; i < v.size(); ++i) {
        out_index[i] = v[i].index;
        out_score[i] = v[i].score;
        out_src[i] = src.at(v[i].doc_id);
    }
    return v.size();
}
}  // namespace

// blend_two_tier (crates/frankensearch-fusion/src/blend.rs:107-191).  out_src[i] >= 0: position
// in the fast list of the doc; < 0: -(position in quality list) - 1 (quality-only doc).
FSO_API uint64_t fso_blend_two_tier(const uint8_t* fast_bytes, const uint64_t* fast_off,
                                    const uint32_t* fast_index, const float* fast_scores,
                                    uint64_t n_fast, const uint8_t* q_bytes, const uint64_t* q_off,
                                    const uint32_t* q_index, const float* q_scores, uint64_t n_q,
                                    float blend_factor, uint32_t* out_index, float* out_score,
                                    int64_t* out_src) {
    const float alpha = sanitize_alpha(blend_factor);
    NormBounds fb, qb;
    fb.fit(fast_scores, nullptr, n_fast);
    qb.fit(q_scores, nullptr, n_q);
    struct Pair {
        uint32_t index;
        float fast = 0, quality = 0;
        bool has_fast = false, has_quality = false;
    };
    std::unordered_map<std::string_view, Pair> merged;
    std::unordered_map<std::string_view, int64_t> src;
    std::vector<std::string_view> order;
    for (uint64_t i = 0; i < n_fast; ++i) {
        const auto id = sv(fast_bytes, fast_off, i);
        auto [it, fresh] = merged.try_emplace(id);
        if (fresh) {
            it->second.index = fast_index[i];
            order.push_back(id);
            src[id] = (int64_t)i;
        }
        if (!it->second.has_fast) {
            it->second.has_fast = true;
            it->second.fast = fb.apply(fast_scores[i]);
            it->second.index = fast_index[i];
        }
    }
    for (uint64_t i = 0; i < n_q; ++i) {
        const auto id = sv(q_bytes, q_off, i);
        auto [it, fresh] = merged.try_emplace(id);
        if (fresh) {
            it->second.index = q_index[i];
            order.push_back(id);
            src[id] = -(int64_t)i - 1;
        }
        if (!it->second.has_quality) {
            it->second.has_quality = true;
            it->second.quality = qb.apply(q_scores[i]);
        }
    }
    std::vector<Blended> v;
    v.reserve(order.size());
    for (const auto& id : order) {
        const Pair& p = merged[id];
        float s;
        if (p.has_fast && p.has_quality)
            s = std::fmaf(alpha, p.quality, (1.0f - alpha) * p.fast);  // alpha.mul_add(q,(1-a)*f)
        else if (p.has_fast)
            s = p.fast;
        else if (p.has_quality)
            s = p.quality;
        else
            s = 0.0f;
        v.push_back(Blended{id, p.index, sanitize_score(s)});
    }
    return emit_blended(v, out_index, out_score, out_src, src);
}

// blend_two_tier_aligned / _aligned_unique (blend.rs:213-286, :296-338): quality_scores[i] is
// the optional quality score of fast hit i (present[i] == 0 -> None).
FSO_API uint64_t fso_blend_two_tier_aligned(const uint8_t* fast_bytes, const uint64_t* fast_off,
                                            const uint32_t* fast_index, const float* fast_scores,
                                            uint64_t n_fast, const float* quality_scores,
                                            const uint8_t* quality_present, float blend_factor,
                                            uint32_t* out_index, float* out_score,
                                            int64_t* out_src) {
    const float alpha = sanitize_alpha(blend_factor);
    NormBounds fb, qb;
    fb.fit(fast_scores, nullptr, n_fast);
    qb.fit(quality_scores, quality_present, n_fast);
    struct Pair {
        uint32_t index;
        float fast = 0, quality = 0;
        bool has_fast = false, has_quality = false;
    };
    std::unordered_map<std::string_view, Pair> merged;
    std::unordered_map<std::string_view, int64_t> src;
    std::vector<std::string_view> order;
    for (uint64_t i = 0; i < n_fast; ++i) {
        const auto id = sv(fast_bytes, fast_off, i);
        auto [it, fresh] = merged.try_emplace(id);
        if (fresh) {
            order.push_back(id);
            src[id] = (int64_t)i;
        }
        if (!it->second.has_fast) {
            it->second.has_fast = true;
            it->second.fast = fb.apply(fast_scores[i]);
            it->second.index = fast_index[i];
        }
        if (quality_present[i] && !it->second.has_quality) {
            it->second.has_quality = true;
            it->second.quality = qb.apply(quality_scores[i]);
        }
    }
    std::vector<Blended> v;
    for (const auto& id : order) {
        const Pair& p = merged[id];
        const float s = p.has_quality ? std::fmaf(alpha, p.quality, (1.0f - alpha) * p.fast) : p.fast;
        v.push_back(Blended{id, p.index, sanitize_score(s)});
    }
    return emit_blended(v, out_index, out_score, out_src, src);
}

// ───────────────────────────── potion / Model2Vec ────────────────────────────────────────
// Model2VecEmbedder::embed_token_ids + finish_mean_pool_and_normalize + accumulate rows
// (crates/frankensearch-embed/src/model2vec_embedder.rs:312-335, :435-451;
//  crates/frankensearch-embed/src/simd.rs:74-116): OOV ids dropped, rows added element-wise in
// token order, x(1/count), sequential norm_sq, x 1/sqrt or zero-fill.
FSO_API uint32_t fso_potion_embed(const float* table, uint64_t vocab, uint32_t dim,
                                  const uint32_t* ids, uint64_t n_ids, float* out) {
    for (uint32_t d = 0; d < dim; ++d) out[d] = 0.0f;
    uint32_t count = 0;
    for (uint64_t t = 0; t < n_ids; ++t) {
        if ((uint64_t)ids[t] >= vocab) continue;
        const float* row = table + (uint64_t)ids[t] * dim;
        for (uint32_t d = 0; d < dim; ++d) out[d] = out[d] + row[d];
        ++count;
    }
    if (count == 0) return 0;
    const float inv = 1.0f / (float)count;
    float norm_sq = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) {
        out[d] = out[d] * inv;
        norm_sq = norm_sq + out[d] * out[d];
    }
    if (std::isfinite(norm_sq) && norm_sq > 1.1920929e-07f) {
        const float inv_norm = 1.0f / std::sqrt(norm_sq);
        for (uint32_t d = 0; d < dim; ++d) out[d] = out[d] * inv_norm;
    } else {
        for (uint32_t d = 0; d < dim; ++d) out[d] = 0.0f;
    }
    return count;
}

// core l2_normalize contract used by the FastEmbed adapter's second normalisation
// (crates/frankensearch-embed/src/fastembed_embedder.rs:416-426): sequential sum of squares,
// scale by 1/sqrt when norm_sq is finite and > f32::EPSILON, else zero-fill.
FSO_API void fso_l2_normalize(float* v, uint32_t dim) {
    float norm_sq = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) norm_sq = norm_sq + v[d] * v[d];
    if (std::isfinite(norm_sq) && norm_sq > 1.1920929e-07f) {
        const float inv = 1.0f / std::sqrt(norm_sq);
        for (uint32_t d = 0; d < dim; ++d) v[d] = v[d] * inv;
    } else {
        for (uint32_t d = 0; d < dim; ++d) v[d] = 0.0f;
    }
}

// ───────────────────────────── synthetic corpora ─────────────────────────────────────────
// The reference's bench generators, verbatim in behaviour
// (crates/frankensearch-index/benches/fsvi_int8_two_pass.rs:199-231): xorshift64 raw vectors,
// sequential-f32 L2 normalise with the `norm > 1e-12` guard and per-element division,
// clustered rows = normalize(centroid[i % C] + noise * raw_vector(seed)).
namespace {
inline void raw_vector(uint64_t seed, uint32_t dim, float* out) {
    uint64_t s = seed | 1ull;
    for (uint32_t d = 0; d < dim; ++d) {
        s ^= s << 13;
        s ^= s >> 7;
        s ^= s << 17;
        out[d] = (float)(s >> 40) / 8388608.0f - 1.0f;
    }
}
inline void normalize(float* v, uint32_t dim) {
    float acc = 0.0f;
    for (uint32_t d = 0; d < dim; ++d) acc = acc + v[d] * v[d];
    const float norm = std::sqrt(acc);
    if (norm > 1e-12f)
        for (uint32_t d = 0; d < dim; ++d) v[d] = v[d] / norm;
}
}  // namespace

FSO_API void fso_raw_vector(uint64_t seed, uint32_t dim, float* out) { raw_vector(seed, dim, out); }
FSO_API void fso_normalize(float* v, uint32_t dim) { normalize(v, dim); }

// kind 0: uniform  row i = normalize(raw_vector(seed_base + i))
// kind 1: clustered row i = normalize(centroid[i % n_centroids] + noise * raw_vector(seed_base + i))
//         centroid c = normalize(raw_vector(0xc0000000 + c))
// Writes f32 rows [row_start, row_start + n_rows) into out_f32 (nullable) and their RNE f16
// encoding into out_f16 (nullable).
FSO_API void fso_synth_rows(int kind, uint64_t seed_base, uint64_t row_start, uint64_t n_rows,
                            uint32_t dim, uint32_t n_centroids, float noise, int threads,
                            float* out_f32, uint16_t* out_f16) {
    std::vector<float> centroids;
    if (kind == 1) {
        centroids.resize((size_t)n_centroids * dim);
        for (uint32_t c = 0; c < n_centroids; ++c) {
            raw_vector(0xc0000000ull + c, dim, centroids.data() + (size_t)c * dim);
            normalize(centroids.data() + (size_t)c * dim, dim);
        }
    }
    std::atomic<uint64_t> next{0};
    auto work = [&]() {
        std::vector<float> v(dim);
        for (;;) {
            const uint64_t blk = next.fetch_add(1);
            const uint64_t s = blk * 4096;
            if (s >= n_rows) break;
            const uint64_t e = std::min(n_rows, s + 4096);
            for (uint64_t r = s; r < e; ++r) {
                const uint64_t i = row_start + r;
                raw_vector(seed_base + i, dim, v.data());
                if (kind == 1) {
                    const float* c = centroids.data() + (size_t)(i % n_centroids) * dim;
                    for (uint32_t d = 0; d < dim; ++d) v[d] = c[d] + noise * v[d];
                }
                normalize(v.data(), dim);
                if (out_f32) std::memcpy(out_f32 + r * dim, v.data(), (size_t)dim * 4);
                if (out_f16) fso_encode_f32_to_f16(v.data(), dim, out_f16 + r * dim, 1);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < std::max(1, threads); ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
}
