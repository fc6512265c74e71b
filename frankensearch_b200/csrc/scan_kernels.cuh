// scan_kernels.cuh — the fused exact f16 scan + top-k kernels.
//
// Reference path replaced: VectorIndex::search_top_k_internal -> scan_parallel ->
// scan_range_chunk -> dot_product_f16_bytes_f32 -> insert_candidate -> merge_partial_heaps ->
// resolve_hits (crates/frankensearch-index/src/search.rs:426-494, :1013-1036, :1257-1327,
// :1688-1720, :1493-1500; simd.rs:398-446).
//
// Data layout in HBM: the slab is n_rows x dim IEEE f16, row-major, 16-byte aligned rows
// (dim % 8 == 0 on the fast path).  One pass streams it exactly once for QB queries.
#pragma once

#include "fsgpu_common.cuh"

namespace fsgpu {

constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / 32;

// Device-side "redo" launches: the batched tensor-core search flags the queries its bound cannot
// cover (non-finite / overflowing components, a candidate list that overflowed) and the SAME stream
// then runs the exact kernels below over just those queries — no host round trip.  A launch with
// `flags` set serves the flagged queries of ranks [first, first + max) (rank = position among the
// flagged, ascending query index), one after another, partial slot i = rank - first.
constexpr uint32_t kRedoMaxPerLaunch = 1024;
struct RedoArgs {
    const uint32_t* flags;  // [n] != 0: the query needs the exact path; nullptr = a normal launch
    const uint32_t* any;    // nullable: nothing is flagged when *any == 0 (the common case: exit at once)
    uint32_t n;             // queries of the (sub-)batch
    uint32_t first, max;    // flagged ranks served by this launch (max <= kRedoMaxPerLaunch)
    uint32_t limit;         // more flagged queries than this: serve none and count a bail (the caller
                            // re-runs the batch another way)
    uint32_t* bail;         // bail counter
    uint32_t* slots;        // [max] out: query index whose partials sit in slot i (written by CTA 0)
    uint32_t* round_n;      // out: slots filled by this launch
    uint32_t* served;       // nullable: += flagged queries (CTA 0 of the first launch)
};

struct ScanArgs {
    RedoArgs redo;
    const uint16_t* slab;      // [n_rows, dim] f16 bits
    const uint8_t* tombstones; // packed bitmap or nullptr
    const float* queries;      // [QB, dim] f32 (device)
    uint64_t n_rows;           // local rows
    uint64_t row_base;         // global row of local row 0
    uint32_t dim;
    uint32_t k;                // per-CTA keep
    uint32_t cap;              // candidate buffer capacity (power of two)
    uint32_t sync_every;       // tiles between compaction checks
    int reduce_order;
    int tail_fma;
    int allow_packed;          // 0 forces the scalar mul.rn/add.rn path (A/B switch)
    uint64_t* partial;         // [gridDim.x, QB, k] keys, 0-padded
    uint32_t* error_flag;      // set to 1 on a capacity contract violation
};

__device__ __forceinline__ uint4 ld_stream_16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void unpack8(const uint4& x, float f[8]) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&x.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&x.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&x.z));
    const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&x.w));
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
    f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

// CTA-collective: collects the flagged queries of ranks [first, first + max) into `list` (shared,
// kRedoMaxPerLaunch entries) and returns how many this launch serves (0: nothing to do).  Every CTA of
// the launch computes the same list; CTA 0 publishes it for the merge launch that follows.
__device__ __forceinline__ uint32_t redo_collect(const RedoArgs& r, uint32_t* list) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_running;
    if (r.any && *r.any == 0u) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *r.round_n = 0u;
        return 0u;
    }
    if (threadIdx.x == 0) s_running = 0u;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (uint32_t base = 0; base < r.n; base += blockDim.x) {  // CTA-uniform trip count
        const uint32_t b = base + threadIdx.x;
        const bool f = b < r.n && r.flags[b] != 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        uint32_t before = s_running;
        for (uint32_t w = 0; w < warp; ++w) before += s_warp[w];
        const uint32_t rank = before + __popc(m & ((1u << lane) - 1u));
        if (f && rank >= r.first && rank - r.first < r.max) list[rank - r.first] = b;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
            for (uint32_t w = 0; w < n_warps; ++w) t += s_warp[w];
            s_running += t;
        }
        __syncthreads();
    }
    const uint32_t total = s_running;
    if (total > r.limit) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (r.first == 0u) atomicAdd(r.bail, 1u);
            *r.round_n = 0u;
        }
        return 0u;
    }
    const uint32_t mine = total > r.first ? min(r.max, total - r.first) : 0u;
    if (blockIdx.x == 0) {
        for (uint32_t i = threadIdx.x; i < mine; i += blockDim.x) r.slots[i] = list[i];
        if (threadIdx.x == 0) {
            *r.round_n = mine;
            if (r.served && r.first == 0u) atomicAdd(r.served, total);
        }
    }
    return mine;
}

// Shared-memory carve-up shared by the scan kernels.
struct ScanSmem {
    uint64_t* cand;  // [QB][cap]
    uint64_t* tau;   // [QB]
    float* q;        // [QB][dim]
    uint32_t* cnt;   // [QB]
};
__host__ __device__ inline size_t scan_smem_bytes(int qb, uint32_t cap, uint32_t dim) {
    const size_t tau_slots = ((size_t)qb + 1) & ~(size_t)1;  // keeps q[] 16-byte aligned (float4 loads)
    return (size_t)qb * cap * 8 + tau_slots * 8 + (size_t)qb * dim * 4 + (size_t)qb * 4 + 16;
}
__device__ __forceinline__ ScanSmem carve_scan_smem(unsigned char* base, int qb, uint32_t cap,
                                                    uint32_t dim) {
    ScanSmem s;
    s.cand = reinterpret_cast<uint64_t*>(base);
    s.tau = s.cand + (size_t)qb * cap;
    s.q = reinterpret_cast<float*>(s.tau + ((qb + 1) & ~1));
    s.cnt = reinterpret_cast<uint32_t*>(s.q + (size_t)qb * dim);
    return s;
}

// CTA-collective: at a sync point, compact every buffer that could overflow before the next
// one and refresh the per-thread float thresholds.
template <int QB>
__device__ __forceinline__ void scan_sync_point(const ScanSmem& sm, uint32_t cap, uint32_t k,
                                                uint32_t trigger, bool force, float thr[QB]) {
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        const uint32_t c = sm.cnt[qi];
        if (force || c > trigger) {  // CTA-uniform: read after the barrier
            CandBuf b{sm.cand + (size_t)qi * cap, sm.cnt + qi, sm.tau + qi};
            cand_compact(b, cap, k);
        }
    }
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        const uint64_t t = sm.tau[qi];
        thr[qi] = t ? key_score(t) : -INFINITY;
    }
}

template <int QB>
__device__ __forceinline__ void scan_offer(const ScanSmem& sm, const ScanArgs& args, int qi,
                                           float score, uint64_t local_row) {
    const uint64_t key = make_key(score, (uint32_t)(args.row_base + local_row));
    if (key > sm.tau[qi] && !tombstoned(args.tombstones, local_row)) {
        CandBuf b{sm.cand + (size_t)qi * args.cap, sm.cnt + qi, sm.tau + qi};
        if (!cand_push(b, args.cap, key)) atomicExch(args.error_flag, 1u);
    }
}

template <int QB>
__device__ __forceinline__ void scan_write_partials(const ScanSmem& sm, const ScanArgs& args, uint32_t slot = 0) {
    for (int qi = 0; qi < QB; ++qi) {
        const uint32_t c = min(sm.cnt[qi], args.k);
        uint64_t* out = args.partial + (((size_t)slot * gridDim.x + blockIdx.x) * QB + qi) * args.k;
        const uint64_t* src = sm.cand + (size_t)qi * args.cap;
        for (uint32_t i = threadIdx.x; i < args.k; i += blockDim.x) out[i] = i < c ? src[i] : 0ull;
    }
}

// ─── fast path: dim = 32*NJ, QB queries per pass, R row-groups per thread ───────────────────
// Thread mapping (per warp): lane = 4r + a handles accumulator `a` (chunks 4j+a) of row r of
// the warp's 8-row group: eight chains (a, l=0..7) per query live in registers, 16-byte loads
// cover a row with 4 ADJACENT lanes x NJ loads (each warp-level load = 8 rows x 64 contiguous
// bytes; a quarter-warp touches 2 half-lines = 2 L1 wavefronts.  The first version used
// lane = 8a + r: 32 wavefronts per load and l1tex at 97 % — profiles/r01_scan_l1tex.md).
// The reference tree `(s0+s1)+(s2+s3)` is two xor-shuffles (1, 2); every lane then holds
// V[0..7] and applies the configured 8-lane order.  Products and sums are separate
// IEEE roundings (mul.rn / add.rn), so scores are bit-identical to simd.rs:398-446.
// Packed-pair arithmetic (sm_100 f32x2 pipe): one issue slot multiplies / adds two chains.
// ptxas contracts `mul.rn.f32x2` + `add.rn.f32x2` into FFMA2 (a fused rounding the reference
// does not have), but it cannot fuse across differing FTZ modes, so the add carries `.ftz`.
// That is bit-identical to the IEEE add as long as no operand or result is subnormal, which
// the kernel guarantees by only taking this path when every non-zero |q_i| >= 2^-76: f16
// values are integer multiples of 2^-24, so every product and every partial sum is an integer
// multiple of 2^-123 — never a non-zero value below 2^-126.  Other queries use the scalar path.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2_rn(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t add2_rn_ftz(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// Scores of R row-groups x QB queries for one warp iteration; v[rr][qi][0..7] on every lane.
template <int NJ, int QB, int R, bool PACKED>
__device__ __forceinline__ void scan_rows(const uint4 (&x)[R][NJ], const float* __restrict__ q_s,
                                          int a, float (&v)[R][QB][8]) {
    constexpr uint32_t dim = NJ * 32;
    if constexpr (PACKED) {
        uint64_t acc[R][QB][4];
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
            for (int qi = 0; qi < QB; ++qi)
#pragma unroll
                for (int m = 0; m < 4; ++m) acc[rr][qi][m] = 0ull;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            uint64_t xp[R][4];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                float xf[8];
                unpack8(x[rr][j], xf);
#pragma unroll
                for (int m = 0; m < 4; ++m) xp[rr][m] = pack2(xf[2 * m], xf[2 * m + 1]);
            }
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const ulonglong2* qp = reinterpret_cast<const ulonglong2*>(q_s + qi * dim + (4 * j + a) * 8);
                const ulonglong2 q0 = qp[0], q1 = qp[1];
                const uint64_t qv[4] = {q0.x, q0.y, q1.x, q1.y};
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        acc[rr][qi][m] = add2_rn_ftz(acc[rr][qi][m], mul2_rn(xp[rr][m], qv[m]));
            }
        }
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
            for (int qi = 0; qi < QB; ++qi)
#pragma unroll
                for (int m = 0; m < 4; ++m) unpack2(acc[rr][qi][m], v[rr][qi][2 * m], v[rr][qi][2 * m + 1]);
    } else {
#pragma unroll
        for (int rr = 0; rr < R; ++rr)
#pragma unroll
            for (int qi = 0; qi < QB; ++qi)
#pragma unroll
                for (int l = 0; l < 8; ++l) v[rr][qi][l] = 0.0f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            float xf[R][8];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) unpack8(x[rr][j], xf[rr]);
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const float4* qp = reinterpret_cast<const float4*>(q_s + qi * dim + (4 * j + a) * 8);
                const float4 q0 = qp[0], q1 = qp[1];
                const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                for (int rr = 0; rr < R; ++rr)
#pragma unroll
                    for (int l = 0; l < 8; ++l)
                        v[rr][qi][l] = add_rn(v[rr][qi][l], mul_rn(xf[rr][l], qv[l]));
            }
        }
    }
    // reference tree (s0+s1)+(s2+s3) across the four lanes of a row
#pragma unroll
    for (int rr = 0; rr < R; ++rr)
#pragma unroll
        for (int qi = 0; qi < QB; ++qi)
#pragma unroll
            for (int l = 0; l < 8; ++l) {
                float s = v[rr][qi][l];
                s = add_rn(s, __shfl_xor_sync(0xffffffffu, s, 1));  // s0+s1 | s2+s3
                s = add_rn(s, __shfl_xor_sync(0xffffffffu, s, 2));  // (s0+s1)+(s2+s3)
                v[rr][qi][l] = s;
            }
}

template <int NJ, int QB, int R, bool PACKED>
__device__ __forceinline__ void scan_loop(const ScanArgs& args, const ScanSmem& sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int a = lane & 3, r = lane >> 2;
    constexpr int kRowsPerWarp = 8 * R;
    constexpr int kTileRows = kScanWarps * kRowsPerWarp;
    const uint64_t n = args.n_rows;
    const uint64_t n_tiles = (n + kTileRows - 1) / kTileRows;
    const uint32_t trigger = args.cap - args.sync_every * kTileRows;
    const uint4* slab4 = reinterpret_cast<const uint4*>(args.slab);

    float thr[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) thr[qi] = -INFINITY;

    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint64_t row0 = tile * kTileRows + (uint64_t)warp * kRowsPerWarp + r;
        uint4 x[R][NJ];
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const uint64_t row = row0 + rr * 8;
            const uint64_t rowc = row < n ? row : n - 1;
            const uint4* p = slab4 + rowc * (NJ * 4) + a;
#pragma unroll
            for (int j = 0; j < NJ; ++j) x[rr][j] = ld_stream_16(p + 4 * j);
        }
        float v[R][QB][8];
        scan_rows<NJ, QB, R, PACKED>(x, sm.q, a, v);
#pragma unroll
        for (int rr = 0; rr < R; ++rr) {
            const uint64_t row = row0 + rr * 8;
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const float score = reduce8(v[rr][qi], args.reduce_order);
                if (!(score < thr[qi])) {  // rare once tau is established; NaN passes
                    if ((qi & 3) == a && row < n) scan_offer<QB>(sm, args, qi, score, row);
                }
            }
        }
        if ((it + 1) % args.sync_every == 0)
            scan_sync_point<QB>(sm, args.cap, args.k, trigger, false, thr);
    }
    scan_sync_point<QB>(sm, args.cap, args.k, trigger, true, thr);
}

template <int NJ, int QB, int R>
__global__ void __launch_bounds__(kScanThreads)
scan_topk_fast_kernel(const ScanArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t dim = NJ * 32;
    const ScanSmem sm = carve_scan_smem(smem_raw, QB, args.cap, dim);
    __shared__ uint32_t s_redo[QB == 1 ? kRedoMaxPerLaunch : 1];
    uint32_t n_serve = 1;
    if (QB == 1 && args.redo.flags) n_serve = redo_collect(args.redo, s_redo);  // a redo launch (QB == 1 only)
    for (uint32_t slot = 0; slot < n_serve; ++slot) {
        const float* queries = (QB == 1 && args.redo.flags) ? args.queries + (size_t)s_redo[slot] * dim : args.queries;
        bool q_ok = true;  // packed-path precondition: every non-zero |q_i| >= 2^-76
        for (uint32_t i = threadIdx.x; i < QB * dim; i += blockDim.x) {
            const float qv = queries[i];
            sm.q[i] = qv;
            const uint32_t mag = __float_as_uint(qv) & 0x7FFFFFFFu;
            q_ok = q_ok && (mag == 0u || mag >= 0x19800000u);
        }
        if (threadIdx.x < QB) {
            sm.cnt[threadIdx.x] = 0u;
            sm.tau[threadIdx.x] = 0ull;
        }
        const bool packed = __syncthreads_and(q_ok ? 1 : 0) != 0 && args.allow_packed != 0;
        if (packed)
            scan_loop<NJ, QB, R, true>(args, sm);
        else
            scan_loop<NJ, QB, R, false>(args, sm);
        scan_write_partials<QB>(sm, args, slot);
        __syncthreads();  // the next query re-uses the shared buffers
    }
}

// ─── generic path: any dim (tails included), one warp per row, one query per pass ───────────
__global__ void __launch_bounds__(kScanThreads)
scan_topk_generic_kernel(const ScanArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const ScanSmem sm = carve_scan_smem(smem_raw, 1, args.cap, args.dim);
    __shared__ uint32_t s_redo[kRedoMaxPerLaunch];
    uint32_t n_serve = 1;
    if (args.redo.flags) n_serve = redo_collect(args.redo, s_redo);
    for (uint32_t slot = 0; slot < n_serve; ++slot) {
        const float* query = args.redo.flags ? args.queries + (size_t)s_redo[slot] * args.dim : args.queries;
        for (uint32_t i = threadIdx.x; i < args.dim; i += blockDim.x) sm.q[i] = query[i];
        if (threadIdx.x == 0) {
            sm.cnt[0] = 0u;
            sm.tau[0] = 0ull;
        }
        __syncthreads();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int kTileRows = kScanWarps;  // one row per warp per iteration
        const uint64_t n = args.n_rows;
        const uint64_t n_tiles = (n + kTileRows - 1) / kTileRows;
        const uint32_t trigger = args.cap - args.sync_every * kTileRows;
        float thr[1] = {-INFINITY};
        uint32_t it = 0;
        for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const uint64_t row = tile * kTileRows + warp;
            const uint64_t rowc = row < n ? row : n - 1;
            const float score = warp_exact_dot(args.slab + rowc * args.dim, sm.q, args.dim,
                                               args.reduce_order, args.tail_fma);
            if (!(score < thr[0]) && lane == 0 && row < n) scan_offer<1>(sm, args, 0, score, row);
            if ((it + 1) % args.sync_every == 0)
                scan_sync_point<1>(sm, args.cap, args.k, trigger, false, thr);
        }
        scan_sync_point<1>(sm, args.cap, args.k, trigger, true, thr);
        scan_write_partials<1>(sm, args, slot);
        __syncthreads();
    }
}

// ─── merge: one CTA per query over `n_lists` lists of `k_in` keys ───────────────────────────
// merge_partial_heaps (search.rs:1704-1720) + the best-first sort of resolve_hits
// (search.rs:1493-1500).  Also the cross-shard merge after the all-gather.
struct MergeArgs {
    const uint64_t* keys;     // key of (query b, list g, slot i) at keys[g*list_stride + b*query_stride + i]
    const float* scores;      // optional raw scores, same addressing (travel with the keys)
    const fsgpu_hit_t* hits_in;  // ... or the hits that travelled with the keys (score field is used)
    uint64_t list_stride;
    uint64_t query_stride;
    uint32_t n_lists;
    uint32_t k_in;
    const uint32_t* k_in_used;  // optional (n_lists == 1): only the first min(*k_in_used, k_in) slots are filled
    uint32_t k_out;
    uint32_t cap;             // power of two, >= 2*k_out
    uint64_t* out_keys;       // [batch, k_out] (nullable)
    fsgpu_hit_t* out_hits;    // [batch, k_out] (nullable)
    uint32_t* out_counts;     // [batch] (nullable)
    // optional exact re-computation of -inf/NaN class scores from the local slab
    const void* slab;         // f16 slab, or f32 when slab_is_f32 (an f32-quantised FSVI file)
    int slab_is_f32;
    const float* queries;     // [batch, dim]
    uint64_t n_rows, row_base;
    uint32_t dim;
    int reduce_order, tail_fma;
    uint32_t* error_flag;
    // merge of a redo launch: CTA i reads the lists of slot i and writes the result of query
    // redo_slots[i]; CTAs at or past *redo_round_n have nothing to do
    const uint32_t* redo_slots;
    const uint32_t* redo_round_n;
};

__global__ void __launch_bounds__(kScanThreads) merge_topk_kernel(const MergeArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* tau = cand + args.cap;
    uint32_t* cnt = reinterpret_cast<uint32_t*>(tau + 1);
    const uint32_t b_in = blockIdx.x;
    uint32_t b = blockIdx.x;  // the query whose outputs this CTA writes
    if (args.redo_slots) {
        if (b_in >= *args.redo_round_n) return;
        b = args.redo_slots[b_in];
    }
    if (threadIdx.x == 0) {
        *cnt = 0u;
        *tau = 0ull;
    }
    __syncthreads();
    const CandBuf buf{cand, cnt, tau};
    const uint64_t total = args.k_in_used ? (uint64_t)min(*args.k_in_used, args.k_in) : (uint64_t)args.n_lists * args.k_in;
    const uint32_t step = blockDim.x;
    // between compaction checks at most `chunk` pushes can happen
    const uint32_t chunk = max(step, ((args.cap - args.k_out) / 2 / step) * step);
    const uint32_t trigger = args.cap - chunk;
    for (uint64_t base = 0; base < total; base += chunk) {
        const uint64_t t = *tau;
        for (uint32_t o = threadIdx.x; o < chunk; o += step) {
            const uint64_t idx = base + o;
            if (idx < total) {
                const uint64_t g = idx / args.k_in, i = idx % args.k_in;
                const uint64_t key = args.keys[g * args.list_stride + b_in * args.query_stride + i];
                if (key > t) {
                    if (!cand_push(buf, args.cap, key)) atomicExch(args.error_flag, 1u);
                }
            }
        }
        __syncthreads();
        if (*cnt > trigger) cand_compact(buf, args.cap, args.k_out);
        __syncthreads();
    }
    cand_compact(buf, args.cap, args.k_out);
    const uint32_t count = *cnt;
    if (threadIdx.x == 0 && args.out_counts) args.out_counts[b] = count;
    if (args.out_keys)
        for (uint32_t i = threadIdx.x; i < args.k_out; i += step)
            args.out_keys[(size_t)b * args.k_out + i] = i < count ? cand[i] : 0ull;
    if (args.out_hits) {
        for (uint32_t i = threadIdx.x; i < args.k_out; i += step) {
            fsgpu_hit_t h;
            h.row = i < count ? key_row(cand[i]) : 0xFFFFFFFFu;
            h.score = i < count ? key_score(cand[i]) : 0.0f;
            args.out_hits[(size_t)b * args.k_out + i] = h;
        }
        __syncthreads();
        // score_key folded NaN into -inf for ordering; VectorHit carries the RAW score
        // (search.rs:1549-1553), so -inf class entries get their raw value back.
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (uint32_t i = warp; i < count; i += kScanWarps) {
            const uint64_t key = cand[i];
            if ((uint32_t)(key >> 32) != kNegInfOrdered) continue;  // warp-uniform
            float raw = -INFINITY;
            bool have = false;
            if (args.scores || args.hits_in) {
                for (uint64_t idx = lane; idx < total && !have; idx += 32) {
                    const uint64_t g = idx / args.k_in, j = idx % args.k_in;
                    const uint64_t off = g * args.list_stride + b_in * args.query_stride + j;
                    if (args.keys[off] == key) {
                        raw = args.scores ? args.scores[off] : args.hits_in[off].score;
                        have = true;
                    }
                }
                const unsigned m = __ballot_sync(0xffffffffu, have);
                if (m) raw = __shfl_sync(0xffffffffu, raw, __ffs(m) - 1);
                have = m != 0;
            }
            if (!have && args.slab) {
                const uint64_t grow = key_row(key);
                if (grow >= args.row_base && grow - args.row_base < args.n_rows) {
                    raw = warp_exact_row(args.slab, args.slab_is_f32, grow - args.row_base,
                                         args.queries + (size_t)b * args.dim, args.dim, args.reduce_order, args.tail_fma);
                }
            }
            if (lane == 0) args.out_hits[(size_t)b * args.k_out + i].score = raw;
        }
    }
}

// ─── payload of merged keys ─────────────────────────────────────────────────────────────────
// After a cross-shard merge: the value that travelled with each surviving key (e.g. the quality-tier
// score its owning rank computed before the all-gather, two_tier.rs:1566-1631).  The input lists are
// the per-rank results — best first, i.e. DESCENDING keys, 0-padded — so the key is found by a binary
// search per list; keys are unique across lists (distinct global rows).
__global__ void __launch_bounds__(256)
merge_payload_kernel(const uint64_t* __restrict__ keys, uint64_t list_stride, uint64_t query_stride,
                     uint32_t n_lists, uint32_t k_in, const float* __restrict__ payload, uint64_t pl_list_stride,
                     uint64_t pl_query_stride, const uint64_t* __restrict__ merged_keys, uint32_t k_out,
                     float* __restrict__ out_payload, uint8_t* __restrict__ out_present) {
    const uint32_t b = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k_out) return;
    const uint64_t key = merged_keys[(size_t)b * k_out + i];
    float val = 0.0f;
    bool found = false;
    if (key != 0ull) {
        for (uint32_t g = 0; g < n_lists && !found; ++g) {
            const uint64_t* list = keys + g * list_stride + b * query_stride;
            uint32_t lo = 0, hi = k_in;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (list[mid] > key) lo = mid + 1; else hi = mid;
            }
            if (lo < k_in && list[lo] == key) {
                val = payload[g * pl_list_stride + b * pl_query_stride + lo];
                found = true;
            }
        }
    }
    out_payload[(size_t)b * k_out + i] = val;
    if (out_present) out_present[(size_t)b * k_out + i] = found ? 1 : 0;
}

// ─── resident WAL rows: scan_wal (search.rs:1449-1475) ──────────────────────────────────────
// One warp per (query, WAL row): exact f32 dot, non-finite scores skipped (search.rs:1466-1470),
// optional allow bit (the filter evaluated by the host, search.rs:1457-1465).  WAL row w carries the
// hit row `wal_base + w` (record_count + wal_idx, search.rs:1583-1597), which also reproduces the
// heap order of the tagged index (every main row before every WAL row on equal scores, lower WAL
// position first: wal.rs:557-569 with search.rs:1673-1678).  Writes the per-query list
// [main keys (k) | WAL keys (n_wal)] that merge_topk_kernel reduces to the final top-k.
__global__ void __launch_bounds__(kScanThreads)
wal_keys_kernel(const float* __restrict__ wal, uint32_t n_wal, uint64_t wal_base, uint32_t dim,
                const float* __restrict__ queries, const uint8_t* __restrict__ allow, uint64_t allow_bit0,
                const uint64_t* __restrict__ main_keys, uint32_t k, int reduce_order,
                uint64_t* __restrict__ out) {
    const uint32_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const size_t stride = (size_t)k + n_wal;
    if (blockIdx.x == 0)
        for (uint32_t i = threadIdx.x; i < k; i += blockDim.x)
            out[b * stride + i] = main_keys ? main_keys[(size_t)b * k + i] : 0ull;
    const uint32_t w = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
    if (w >= n_wal) return;
    uint64_t key = 0ull;
    const uint64_t bit = allow_bit0 + w;
    const bool allowed = allow == nullptr || ((__ldg(allow + (bit >> 3)) >> (bit & 7)) & 1u);
    if (allowed) {
        const float s = warp_exact_dot_f32(wal + (size_t)w * dim, queries + (size_t)b * dim, dim, reduce_order);
        if (isfinite(s)) key = make_key(s, (uint32_t)(wal_base + w));
    }
    if (lane == 0) out[b * stride + k + w] = key;
}

// ─── doc-id-hash filters on the device ──────────────────────────────────────────────────────
// BitsetFilter (crates/frankensearch-core/src/filter.rs:330-383): `matches_doc_id_hash(h) ==
// Some(set.contains(h))`, applied to the 8-byte hash of each record before heap admission
// (search.rs:1329-1447).  One thread per 8 rows: binary search of each row's hash in the sorted
// allow-list, one byte of the EXCLUSION bitmap (tombstones | !allowed) out.  With `positions` the
// allowed live rows are also appended to a list — the selective arm, try_gather_filtered
// (search.rs:1114-1161), which scores only those rows; `*n_positions` may exceed `cap` (overflow:
// the host falls back to the scan, the result is the same).
__device__ __forceinline__ bool hash_in_sorted(const uint64_t* __restrict__ allowed, uint32_t n, uint64_t h) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        const uint64_t v = __ldg(allowed + mid);
        if (v < h) lo = mid + 1; else hi = mid;
    }
    return lo < n && __ldg(allowed + lo) == h;
}

__global__ void __launch_bounds__(256)
hash_filter_kernel(const uint64_t* __restrict__ row_hashes, uint64_t n_rows, const uint8_t* __restrict__ tomb,
                   const uint64_t* __restrict__ allowed, uint32_t n_allowed, uint8_t* __restrict__ excl,
                   uint32_t* __restrict__ positions, uint32_t cap, uint32_t* __restrict__ n_positions) {
    const uint64_t n_bytes = (n_rows + 7) / 8;
    for (uint64_t byte = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; byte < n_bytes;
         byte += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t dead = tomb ? tomb[byte] : 0u;
        uint32_t out = 0xFFu;
        for (uint32_t j = 0; j < 8; ++j) {
            const uint64_t r = byte * 8 + j;
            if (r >= n_rows) break;
            if (((dead >> j) & 1u) == 0 && hash_in_sorted(allowed, n_allowed, row_hashes[r])) {
                out &= ~(1u << j);
                if (positions) {
                    const uint32_t pos = atomicAdd(n_positions, 1u);
                    if (pos < cap) positions[pos] = (uint32_t)r;
                }
            }
        }
        excl[byte] = (uint8_t)out;
    }
}

// gather_range (search.rs:1196-1255): exact score of each listed row, one warp per (query, row).
// Slots past the list are key 0 (empty); merge_topk_kernel keeps the best k.
__global__ void __launch_bounds__(kScanThreads)
gather_keys_kernel(const uint16_t* __restrict__ slab, uint64_t row_base, uint32_t dim,
                   const float* __restrict__ queries, const uint32_t* __restrict__ positions,
                   const uint32_t* __restrict__ n_positions, uint32_t cap, int reduce_order, int tail_fma,
                   uint64_t* __restrict__ out) {
    const uint32_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
    if (i >= cap) return;
    uint64_t key = 0ull;
    if (i < min(*n_positions, cap)) {
        const uint32_t row = positions[i];
        const float s = warp_exact_dot(slab + (size_t)row * dim, queries + (size_t)b * dim, dim, reduce_order, tail_fma);
        key = make_key(s, (uint32_t)(row_base + row));
    }
    if (lane == 0) out[(size_t)b * cap + i] = key;
}

// ─── single-query int8 pass 1 ─────────────────────────────────────────────
// The reference's int8 two-pass idea (search.rs:514-998: an int8 scan picks candidates, an exact
// f16 re-score ranks them) made EXACT.  Pass 1 streams the int8 codes (half the bytes of the f16
// slab: this is the HBM-bound regime), writes every row's approximate score and keeps the CTA's
// best k of them.  The caller then (i8_gate_kernel) re-scores the approximate top-k exactly: their
// k-th best exact score tau is a lower bound of the corpus' true k-th best, and every row of the
// true top-k has approx >= tau - 2e (|approx - reference| <= e, mma_prep_queries_i8_kernel), so
// i8_select_kernel lists exactly those rows and gather_keys_kernel + merge_topk_kernel rank them
// with the reference arithmetic.  For one query the list is a few dozen rows.
// Lane mapping: 8 adjacent lanes cover one row with 16-byte loads (4 rows per warp-level load,
// 128 contiguous bytes each), dp4a against the query codes held in registers, 3 xor-shuffles.
struct I8ScanArgs {
    const int8_t* codes;        // [n_rows, dim]
    const uint8_t* tombstones;  // packed bitmap or nullptr
    const int8_t* q_codes;      // [QB, dim]
    const float* qscale;        // [QB] score = acc * qscale
    uint64_t n_rows, row_base;
    uint32_t dim;               // multiple of 128, <= 512
    uint32_t k, cap, sync_every;
    float* approx;              // [QB, n_rows] out; -inf for excluded rows
    uint64_t* partial;          // [gridDim.x, QB, k] approximate keys
    uint32_t* error_flag;
};

constexpr int kI8RowsPerIter = kScanWarps * 4 * 2;  // 4 rows per warp-load, 2 loads in flight per lane

// QB queries share one pass over the codes (1, 2 or 4: the pass stays HBM-bound, dp4a work grows).
template <int QB>
__global__ void __launch_bounds__(kScanThreads) scan_i8_kernel(const I8ScanArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* cand = reinterpret_cast<uint64_t*>(smem_raw);          // [QB][cap]
    uint64_t* tau = cand + (size_t)QB * args.cap;                      // [QB]
    uint32_t* cnt = reinterpret_cast<uint32_t*>(tau + QB);            // [QB]
    if (threadIdx.x < QB) {
        cnt[threadIdx.x] = 0u;
        tau[threadIdx.x] = 0ull;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane & 7, rr = lane >> 3;
    const uint32_t nj = args.dim >> 7;  // 128-byte segments per row
    int y[QB][4][4];
    float qscale[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        qscale[qi] = args.qscale[qi];
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) {
            if (j < nj) {
                const int4 v = *reinterpret_cast<const int4*>(args.q_codes + (size_t)qi * args.dim + j * 128u + sub * 16u);
                y[qi][j][0] = v.x; y[qi][j][1] = v.y; y[qi][j][2] = v.z; y[qi][j][3] = v.w;
            } else {
                y[qi][j][0] = y[qi][j][1] = y[qi][j][2] = y[qi][j][3] = 0;
            }
        }
    }
    __syncthreads();
    const uint64_t n = args.n_rows;
    const uint64_t n_tiles = (n + kI8RowsPerIter - 1) / kI8RowsPerIter;
    const uint32_t trigger = args.cap - args.sync_every * kI8RowsPerIter;
    float thr[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) thr[qi] = -INFINITY;
    uint32_t it = 0;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const uint64_t row0 = tile * kI8RowsPerIter + (uint64_t)warp * 8u + rr;  // this lane's rows: row0, row0 + 4
        uint4 x[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t row = row0 + 4u * h;
            const uint64_t rowc = row < n ? row : n - 1;
            const uint4* p = reinterpret_cast<const uint4*>(args.codes + rowc * args.dim + sub * 16u);
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (j < nj) x[h][j] = ld_stream_16(p + j * 8u);  // + j * 128 bytes
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint64_t row = row0 + 4u * h;
            const bool mine = sub == 0 && row < n;
            const bool dead = mine && tombstoned(args.tombstones, row);
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                int acc = 0;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j)
                    if (j < nj) {
                        acc = __dp4a((int)x[h][j].x, y[qi][j][0], acc);
                        acc = __dp4a((int)x[h][j].y, y[qi][j][1], acc);
                        acc = __dp4a((int)x[h][j].z, y[qi][j][2], acc);
                        acc = __dp4a((int)x[h][j].w, y[qi][j][3], acc);
                    }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                acc += __shfl_xor_sync(0xffffffffu, acc, 4);
                if (mine) {
                    const float s = __fmul_rn((float)acc, qscale[qi]);
                    args.approx[(size_t)qi * n + row] = dead ? -INFINITY : s;
                    if (!dead && !(s < thr[qi])) {
                        const uint64_t key = make_key(s, (uint32_t)(args.row_base + row));
                        const CandBuf buf{cand + (size_t)qi * args.cap, cnt + qi, tau + qi};
                        if (key > tau[qi] && !cand_push(buf, args.cap, key)) atomicExch(args.error_flag, 1u);
                    }
                }
            }
        }
        if ((it + 1) % args.sync_every == 0) {
            __syncthreads();
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const CandBuf buf{cand + (size_t)qi * args.cap, cnt + qi, tau + qi};
                if (cnt[qi] > trigger) cand_compact(buf, args.cap, args.k);  // CTA-uniform
            }
            __syncthreads();
#pragma unroll
            for (int qi = 0; qi < QB; ++qi) {
                const uint64_t t = tau[qi];
                thr[qi] = t ? key_score(t) : -INFINITY;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        const CandBuf buf{cand + (size_t)qi * args.cap, cnt + qi, tau + qi};
        cand_compact(buf, args.cap, args.k);
    }
    __syncthreads();
    for (int qi = 0; qi < QB; ++qi) {
        const uint32_t c = min(cnt[qi], args.k);
        uint64_t* out = args.partial + ((size_t)blockIdx.x * QB + qi) * args.k;
        const uint64_t* src = cand + (size_t)qi * args.cap;
        for (uint32_t i = threadIdx.x; i < args.k; i += blockDim.x) out[i] = i < c ? src[i] : 0ull;
    }
}

// One CTA: exact scores of the approximate top-k (`keys`, 0 = empty) -> gate = (k-th best exact) -
// margin2 (= 2e, rounded down), or -inf when fewer than k rows exist.  Resets the position count.
__global__ void __launch_bounds__(kScanThreads)
i8_gate_kernel(const uint64_t* __restrict__ keys, uint32_t k, const uint16_t* __restrict__ slab, uint64_t row_base,
               uint32_t dim, const float* __restrict__ query, const float* __restrict__ margin2, int reduce_order,
               int tail_fma, float* __restrict__ gate_out, uint32_t* __restrict__ n_positions) {
    __shared__ float s_min[kScanWarps];
    __shared__ uint32_t s_cnt[kScanWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float mn = INFINITY;
    uint32_t cnt = 0;
    for (uint32_t i = warp; i < k; i += kScanWarps) {
        const uint64_t key = keys[i];
        if (key == 0ull) continue;  // warp-uniform
        const uint64_t local = (uint64_t)key_row(key) - row_base;
        const float s = warp_exact_dot(slab + local * dim, query, dim, reduce_order, tail_fma);
        mn = fminf(mn, s);
        ++cnt;
    }
    if (lane == 0) {
        s_min[warp] = mn;
        s_cnt[warp] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t total = 0;
        for (int w = 0; w < kScanWarps; ++w) {
            mn = fminf(mn, s_min[w]);
            total += s_cnt[w];
        }
        *gate_out = total >= k ? __fsub_rd(mn, *margin2) : -INFINITY;
        *n_positions = 0u;
    }
}

// Rows whose approximate score clears the gate (excluded rows carry -inf and never do).
__global__ void __launch_bounds__(256)
i8_select_kernel(const float* __restrict__ approx, uint64_t n_rows, const float* __restrict__ gate,
                 uint32_t* __restrict__ positions, uint32_t cap, uint32_t* __restrict__ n_positions,
                 uint32_t* __restrict__ overflow) {
    const float g = *gate;
    for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += (uint64_t)gridDim.x * blockDim.x) {
        const float a = approx[r];
        if (a >= g && a != -INFINITY) {
            const uint32_t pos = atomicAdd(n_positions, 1u);
            if (pos < cap)
                positions[pos] = (uint32_t)r;
            else
                atomicExch(overflow, 1u);  // the caller re-runs the query on the f16 scan
        }
    }
}

// ─── gather-dot: quality_scores_for_hits (two_tier.rs:1566-1631, :1946-1973) ────────────────
__global__ void __launch_bounds__(kScanThreads)
scores_for_rows_kernel(const void* __restrict__ slab, int slab_is_f32, uint64_t n_rows, uint64_t row_base,
                       uint32_t dim, const float* __restrict__ queries,
                       const uint32_t* __restrict__ rows, uint32_t row_stride, uint32_t n_per_query,
                       int reduce_order, int tail_fma, float* __restrict__ out_scores,
                       uint8_t* __restrict__ out_present) {
    // `row_stride` = 1 for a plain row array, 2 when the rows are read out of fsgpu_hit records
    const uint32_t b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const uint32_t i = blockIdx.x * kScanWarps + (threadIdx.x >> 5);
    if (i >= n_per_query) return;
    const size_t o = (size_t)b * n_per_query + i;
    const uint64_t grow = rows[o * row_stride];
    const bool ok = grow != 0xFFFFFFFFull && grow >= row_base && grow - row_base < n_rows;
    float s = 0.0f;
    if (ok)
        s = warp_exact_row(slab, slab_is_f32, grow - row_base, queries + (size_t)b * dim, dim, reduce_order, tail_fma);
    if (lane == 0) {
        out_scores[o] = s;
        if (out_present) out_present[o] = ok ? 1 : 0;
    }
}

// ─── zero-signal census (VectorIndex::zero_signal_state, lib.rs:2441-2459) ──────────────────
// counts[0] = tombstoned rows, counts[1] = live rows whose stored vector is usable: every element
// finite and the SEQUENTIAL f32 sum of squares positive and finite (vector_signal_usable,
// lib.rs:6133-6142).  One thread per row, elements in order: a lazy pass that only runs to classify an
// EMPTY result (search.rs:206-260), so simplicity beats coalescing.
__global__ void __launch_bounds__(256)
census_kernel(const void* __restrict__ slab, int slab_is_f32, uint64_t n_rows, uint32_t dim,
              const uint8_t* __restrict__ tombstones, unsigned long long* __restrict__ counts) {
    unsigned long long dead = 0, usable = 0;
    for (uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; row < n_rows; row += (uint64_t)gridDim.x * blockDim.x) {
        if (tombstoned(tombstones, row)) {
            ++dead;
            continue;
        }
        float norm_sq = 0.0f;
        bool finite = true;
        for (uint32_t i = 0; i < dim && finite; ++i) {
            const float v = slab_is_f32 ? static_cast<const float*>(slab)[row * dim + i]
                                        : h2f(static_cast<const uint16_t*>(slab)[row * dim + i]);
            if (!isfinite(v)) finite = false;
            norm_sq = add_rn(norm_sq, mul_rn(v, v));
        }
        if (finite && norm_sq > 0.0f && isfinite(norm_sq)) ++usable;
    }
    for (int o = 16; o > 0; o >>= 1) {
        dead += __shfl_xor_sync(0xffffffffu, dead, o);
        usable += __shfl_xor_sync(0xffffffffu, usable, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (dead) atomicAdd(counts, dead);
        if (usable) atomicAdd(counts + 1, usable);
    }
}

// ─── score-all: the `limit >= n` / very large k arm (search.rs:449-473) ─────────────────────
// Writes one order key per live row (0 for tombstoned rows); the caller sorts descending.
__global__ void __launch_bounds__(kScanThreads)
score_all_kernel(const void* __restrict__ slab, int slab_is_f32, const uint8_t* __restrict__ tombstones,
                 uint64_t n_rows, uint64_t row_base, uint32_t dim,
                 const float* __restrict__ query, int reduce_order, int tail_fma,
                 uint64_t* __restrict__ out_keys) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* q = reinterpret_cast<float*>(smem_raw);
    for (uint32_t i = threadIdx.x; i < dim; i += blockDim.x) q[i] = query[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    for (uint64_t row = (uint64_t)blockIdx.x * kScanWarps + (threadIdx.x >> 5); row < n_rows;
         row += (uint64_t)gridDim.x * kScanWarps) {
        const float s = warp_exact_row(slab, slab_is_f32, row, q, dim, reduce_order, tail_fma);
        if (lane == 0)
            out_keys[row] = tombstoned(tombstones, row) ? 0ull
                                                        : make_key(s, (uint32_t)(row_base + row));
    }
}

// Emits the first `k_eff` keys of a descending-sorted key array as the result of one query:
// count = number of live (non-zero) keys, raw scores restored for the -inf/NaN class.
__global__ void __launch_bounds__(kScanThreads)
emit_sorted_prefix_kernel(const uint64_t* __restrict__ sorted, uint32_t k_eff, uint32_t k_out,
                          const void* __restrict__ slab, int slab_is_f32, const float* __restrict__ query,
                          uint64_t n_rows, uint64_t row_base, uint32_t dim, int reduce_order,
                          int tail_fma, uint64_t* __restrict__ out_keys,
                          fsgpu_hit_t* __restrict__ out_hits, uint32_t* __restrict__ out_count) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && out_count && k_eff == 0) *out_count = 0;
    for (uint32_t i = warp_global; i < k_out; i += n_warps) {  // warp-uniform loop
        const uint64_t key = i < k_eff ? sorted[i] : 0ull;
        if (out_keys && lane == 0) out_keys[i] = key;
        if (key != 0 && lane == 0 && out_count) {
            const bool last = (i + 1 == k_eff) || sorted[i + 1] == 0ull;
            if (last) *out_count = i + 1;
        }
        if (i == 0 && key == 0 && lane == 0 && out_count) *out_count = 0;
        if (out_hits) {
            float score = key ? key_score(key) : 0.0f;
            if (key && (uint32_t)(key >> 32) == kNegInfOrdered) {
                const uint64_t grow = key_row(key);
                if (grow >= row_base && grow - row_base < n_rows)
                    score = warp_exact_row(slab, slab_is_f32, grow - row_base, query, dim, reduce_order, tail_fma);
            }
            if (lane == 0) {
                fsgpu_hit_t h;
                h.row = key ? key_row(key) : 0xFFFFFFFFu;
                h.score = score;
                out_hits[i] = h;
            }
        }
    }
}

}  // namespace fsgpu
