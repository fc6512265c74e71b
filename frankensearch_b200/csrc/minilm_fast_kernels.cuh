// minilm_fast_kernels.cuh — the f16 form of the MiniLM-L6-v2 encoder (the default for query lengths <= 32).
//
// The split-f16 form (minilm_kernels.cuh) reproduces an f32 forward to ~2e-6, but it moves every activation
// three times (f32 + hi + lo halves) and issues every product three times: at 1024 x 32 tokens the encoder
// spent 3.5 ms, ~1.5 ms of it as plain HBM traffic (8.7 GB per batch), against 2.6 ms for the exact scan of
// 10 M rows it feeds.  The contract is "within 1e-3 relative on cosine scores", which f16 OPERANDS with f32
// accumulation meet with two orders of magnitude to spare (tests/test_gpu_minilm.py: cosine to the f32
// reference >= 1 - 1e-6, score error <= 1e-4), so this form carries activations as ONE f16 copy:
//
//   h (f16) --QKV GEMM--> qkv (f16) --attention (mma.sync, one warp per (sequence, head))--> ctx (f16)
//     --out-proj GEMM (+bias)--> pre (f32) --(+ residual h) LayerNorm--> h (f16)
//     --FFN-in GEMM (+bias, erf-GELU)--> ffn (f16) --FFN-out GEMM (+bias)--> pre (f32) --(+h) LayerNorm--> h (f16)
//
// The two pre-LayerNorm sums stay f32 (LayerNorm subtracts a mean: its input is the one place where an f16
// rounding would be amplified).  One GEMM kernel serves all four linears: persistent, warp-specialised
// (TMA -> mbarrier ring -> tcgen05.mma kind::f16, M = 128 x N = 128, f32 accumulators double-buffered in
// TMEM), epilogue warps read their TMEM rows, apply bias / GELU, write the 32 x 64 sub-tile into a swizzled
// shared-memory box and hand it to a TMA STORE — full 128-byte row segments leave the SM without a transpose
// (the row-per-thread global stores of the first epilogue were what bound those GEMMs).
#pragma once

#include <mma.h>

#include "minilm_kernels.cuh"

namespace fsgpu {

constexpr int kFastThreads = 320;  // warps 0-7 epilogue (lane quarter = warp % 4, column half = warp / 4), 8 TMA, 9 MMA
constexpr int kFastStages = 5;     // 5 x (A 16 KiB + W 16 KiB)
constexpr int kFastAcc = 2;        // 2 x 128 TMEM columns

// erf-GELU for the f16 form: 0.5 x (1 + tanh(y(x))) with y = x (a + b x^2 + c x^4) fitted so that tanh(y) follows
// erf(x / sqrt 2) to 3.7e-5 on the whole axis (|x| clamped to 8 inside y: beyond it tanh is 1 to f32 precision
// and the fitted polynomial would turn over), i.e. |GELU error| <= 5.5e-5 + 2.4e-4 |x| from tanh.approx.f32's
// 2^-11 — below the f16 rounding of the result it feeds.  9 instructions and one MUFU instead of ~18 and two:
// the FFN-in epilogue evaluates 50 M of these per layer and was bound by them (131 of the layer's 323 us).
__device__ __forceinline__ float gelu_tanh_fit(float x) {
    const float xc = fminf(fmaxf(x, -8.0f), 8.0f);
    const float x2 = xc * xc;
    float p = fmaf(-0.00031580769466f, x2, 0.036798259397f);
    p = fmaf(p, x2, 0.79771783151f);
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(xc * p));
    return x * fmaf(0.5f, t, 0.5f);
}

struct FastGemmArgs {
    uint32_t m, n, k;     // n % 128 == 0, k % 64 == 0
    const uint32_t* m_ptr;  // packed rows: the row count lives on the device (m is then the upper bound the grid was sized for)
    const float* bias;    // [n]
    int mode;             // 0: f16 out = acc + bias   1: f16 out = gelu(acc + bias)   2: f32 out = acc + bias
};

__host__ __device__ inline size_t fast_gemm_smem_bytes() {
    // ring | 8 warps x 8 KiB staging (f32 mode: two [32 x 32] boxes; f16 modes use the first 4 KiB) | barriers
    return 1024 + (size_t)kFastStages * 2 * kMmaTileBytes + 8 * 8192 + 256 + 384 * 4;
}

// tm_out: f16 modes: [M, N] f16, box [64 cols x 32 rows]; f32 mode: [M, N] f32, box [32 cols x 32 rows]; both
// 128-byte swizzled (row r of a box at r * 128 B, its 16-byte chunk c at position c ^ (r & 7)).
__global__ void __launch_bounds__(kFastThreads, 1)
gemm_f16_fast_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                     const __grid_constant__ CUtensorMap tm_out, const FastGemmArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    constexpr uint32_t per_stage = 2u * kMmaTileBytes;
    const uint32_t stage_smem = base + kFastStages * per_stage;  // 8 x 8 KiB, 1024-byte aligned
    uint8_t* stage_ptr = base_ptr + (size_t)kFastStages * per_stage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_ptr + 8 * 8192);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (20u + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);
    // bias of the whole layer in shared memory when it fits (n <= 384: this kernel's FFN-out / out-projection use);
    // with the L1 carved down to a few KiB every global bias load was an L2 round trip in the epilogue
    float* bias_s = reinterpret_cast<float*>(bars + 32);
    const bool bias_in_smem = args.n <= 384u;
    if (bias_in_smem)
        for (uint32_t i = threadIdx.x; i < args.n; i += blockDim.x) bias_s[i] = args.bias[i];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t m_rows = args.m_ptr ? min(*args.m_ptr, args.m) : args.m;
    const uint32_t tiles_m = (m_rows + 127u) / 128u, tiles_n = args.n / 128u;
    const uint32_t n_tiles = tiles_m * tiles_n, n_kb = args.k / kMmaKBlock;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
        for (uint32_t s = 0; s < kFastStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < kFastAcc; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        fence_barrier_init();
    } else if (warp == 0) {
        tmem_alloc(smem_u32(tmem_slot), kFastAcc * 128);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 8) {
        // ===== TMA producer: tile t = (m tile t / tiles_n, n tile t % tiles_n): CTAs that run together share the A rows
        uint32_t stage = 0, phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int32_t row_a = (int32_t)((t / tiles_n) * 128u), row_w = (int32_t)((t % tiles_n) * 128u);
            for (uint32_t kb = 0; kb < n_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    const uint32_t s0 = base + stage * per_stage;
                    const int32_t kc = (int32_t)(kb * kMmaKBlock);
                    mbar_expect_tx(full_bar(stage), per_stage);
                    tma_load_2d(s0, &tm_a, full_bar(stage), kc, row_a);
                    tma_load_2d(s0 + kMmaTileBytes, &tm_w, full_bar(stage), kc, row_w);
                }
                __syncwarp();
                if (++stage == kFastStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 9) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_f16(128, 128);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * 128u;
            for (uint32_t kb = 0; kb < n_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t s0 = base + stage * per_stage;
                    const uint64_t a_desc = umma_desc_sw128(s0), w_desc = umma_desc_sw128(s0 + kMmaTileBytes);
#pragma unroll
                    for (uint32_t k4 = 0; k4 < kMmaKBlock / 16; ++k4)
                        umma_f16(d_tmem, a_desc + 2u * k4, w_desc + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                    umma_commit(empty_bar(stage));
                    if (kb + 1 == n_kb) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == kFastStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (++acc == kFastAcc) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: TMEM lane = output row, column = output feature; warp = 32 rows x 64 columns =====
        const uint32_t quarter = warp & 3u, half = warp >> 2;
        const uint32_t my_stage = stage_smem + warp * 8192u;
        uint8_t* my_ptr = stage_ptr + (size_t)warp * 8192u;
        const uint32_t sw = lane & 7u;  // this row's swizzle phase
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const uint32_t row0 = (t / tiles_n) * 128u + quarter * 32u;
            const uint32_t col0 = (t % tiles_n) * 128u + half * 64u;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * 128u + half * 64u;
            uint32_t v[2][32];
            tmem_ld_x32(taddr, v[0]);
            tmem_ld_x32(taddr + 32u, v[1]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));  // the accumulator is in registers: hand it back
            if (++acc == kFastAcc) {
                acc = 0;
                acc_phase ^= 1u;
            }
            if (row0 >= m_rows) continue;  // (whole warp: a tile past the last row stores nothing)
            // the previous store of this warp must have finished reading the staging box
            if (lane == 0) tma_store_wait_read();
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const uint32_t col = col0 + c * 32 + j + i;
                        x[i] = __uint_as_float(v[c][j + i]) + (bias_in_smem ? bias_s[col] : __ldg(args.bias + col));
                        if (args.mode == 1) x[i] = gelu_tanh_fit(x[i]);
                    }
                    if (args.mode == 2) {  // f32: box c = columns [32c, 32c+32): 8 chunks of 4 floats per row
                        const uint32_t ch = (uint32_t)j / 4u;
                        float4* d0 = reinterpret_cast<float4*>(my_ptr + c * 4096 + lane * 128u + ((ch ^ sw) << 4));
                        float4* d1 = reinterpret_cast<float4*>(my_ptr + c * 4096 + lane * 128u + (((ch + 1u) ^ sw) << 4));
                        *d0 = make_float4(x[0], x[1], x[2], x[3]);
                        *d1 = make_float4(x[4], x[5], x[6], x[7]);
                    } else {  // f16: one box of 64 columns: 8 chunks of 8 halves per row
                        const uint32_t ch = (uint32_t)(c * 32 + j) / 8u;
                        __half2 h[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
                        *reinterpret_cast<uint4*>(my_ptr + lane * 128u + ((ch ^ sw) << 4)) = *reinterpret_cast<uint4*>(h);
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
                if (args.mode == 2) {
                    tma_store_2d(&tm_out, my_stage, (int32_t)col0, (int32_t)row0);
                    tma_store_2d(&tm_out, my_stage + 4096u, (int32_t)col0 + 32, (int32_t)row0);
                } else {
                    tma_store_2d(&tm_out, my_stage, (int32_t)col0, (int32_t)row0);
                }
                tma_store_commit();
            }
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kFastAcc * 128);
    }
}

// ─── the K <= 384 linears (QKV, out-projection, FFN-in) on CTA pairs with the activation tile RESIDENT ──────
// A 128 x 128 tile with both operands streamed asks the L2->SM fabric for 128 B per MMA clock and gets ~25: the
// single-CTA kernel above runs the tensor pipe at ~20 % (1.64 of the f16 form's 2.17 ms were GEMMs).  Here the
// two CTAs of a cluster share tcgen05.mma.cta_group::2 instructions of M = 256 rows (128 per CTA) x N = 256
// features (each CTA loads 128 weight rows), and a CTA's [128 x K] activation tile stays in shared memory while
// the pair walks over the feature blocks of its work items — only weights stream: 32 B per MMA clock, the ratio
// of the corpus scan's pair kernels.  Work items (256-row tile, 256-feature block) are dealt in contiguous
// m-major ranges, so a pair reloads its activation tile once or twice per GEMM; the reload is K-block by
// K-block behind per-block barriers, as soon as the last item of the old tile has consumed that block.
// N % 256 == 128 leaves a 128-wide block: the N = 128 instruction shape with 64 weight rows per CTA.
constexpr int kAresEpiWarps = 16;  // four per TMEM lane quarter, 64 columns each: the epilogue (bias, GELU, f16 pack) is
                                   // instruction-latency-bound — eight warps ran it at an IPC of ~0.35 per scheduler
constexpr int kAresThreads = 32 * (kAresEpiWarps + 2);  // + TMA producer, MMA issuer (highest warp ids)
constexpr uint32_t kAresMaxKb = 6;  // K <= 384
constexpr uint32_t kAresMaxN = 1536;  // the layer's bias vector is staged in shared memory

struct AresGemmArgs {
    uint32_t m, n, k;    // k = k_chunks * (<= 384)
    const uint32_t* m_ptr;  // packed rows: the row count lives on the device (m is then the upper bound the grid was sized for)
    uint32_t k_chunks;   // > 1 (FFN-out, K = 1536): the pair owns whole 256-row tiles and walks (K chunk, feature block)
                         // with the activation tile of the chunk resident; a block's accumulator lives through all
                         // chunks, so n <= 512 (the two TMEM accumulators = the tile's two feature blocks)
    uint32_t n_stages;   // W ring depth (16 KiB stages)
    const float* bias;
    int mode;            // as FastGemmArgs
};

__host__ __device__ inline size_t ares_gemm_smem_bytes(uint32_t n_kb, uint32_t n_stages) {
    return 1024 + (size_t)(n_kb + n_stages) * kMmaTileBytes + kAresEpiWarps * 4096 + 512 + kAresMaxN * 4;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAresThreads, 1)
gemm_f16_ares_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w,
                          const __grid_constant__ CUtensorMap tm_w64, const __grid_constant__ CUtensorMap tm_out,
                          const AresGemmArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t n_kb = args.k / args.k_chunks / kMmaKBlock;  // K-blocks of one chunk
    const uint32_t a_smem = base, w_smem = base + n_kb * kMmaTileBytes;
    const uint32_t stage_smem = w_smem + args.n_stages * kMmaTileBytes;  // one 4 KiB box per epilogue warp
    uint8_t* stage_ptr = base_ptr + (size_t)(n_kb + args.n_stages) * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage_ptr + kAresEpiWarps * 4096);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (18u + a); };
    auto afull_bar = [&](uint32_t kb) { return bar0 + 8u * (20u + kb); };
    auto aempty_bar = [&](uint32_t kb) { return bar0 + 8u * (26u + kb); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 32);
    // the bias vector in shared memory: with the L1 carved down to a few KiB every global bias load was an L2 round
    // trip inside the epilogue — 54 % of this kernel's stall samples (profiles/r02_minilm_ares_qkv_ncu.md)
    float* bias_s = reinterpret_cast<float*>(bars + 40);
    for (uint32_t i = threadIdx.x; i < args.n; i += blockDim.x) bias_s[i] = args.bias[i];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t m_rows = args.m_ptr ? min(*args.m_ptr, args.m) : args.m;
    const uint32_t m_tiles = (m_rows + 255u) / 256u, n_blocks = (args.n + 255u) / 256u;
    const uint32_t kch = args.k_chunks, per_mt = kch * n_blocks;  // items of one 256-row tile: (K chunk, feature block)
    // contiguous m-major item ranges; with K chunks a pair owns whole tiles (its accumulators live across the chunks)
    const uint32_t item0 = kch > 1 ? (uint32_t)((uint64_t)m_tiles * pair / n_pairs) * per_mt
                                   : (uint32_t)((uint64_t)m_tiles * n_blocks * pair / n_pairs);
    const uint32_t item1 = kch > 1 ? (uint32_t)((uint64_t)m_tiles * (pair + 1) / n_pairs) * per_mt
                                   : (uint32_t)((uint64_t)m_tiles * n_blocks * (pair + 1) / n_pairs);

    if (warp == kAresEpiWarps && lane == 0) {
        tma_prefetch_desc(&tm_a);
        tma_prefetch_desc(&tm_w);
        tma_prefetch_desc(&tm_out);
        for (uint32_t s = 0; s < args.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 2 * kAresEpiWarps);  // one arrival per epilogue warp of BOTH CTAs
        }
        for (uint32_t kb = 0; kb < kAresMaxKb; ++kb) {
            mbar_init(afull_bar(kb), 1);
            mbar_init(aempty_bar(kb), 1);
        }
        fence_barrier_init();
    } else if (warp == 0) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == kAresEpiWarps) {
        // ===== TMA producer (both CTAs; completion bytes land on the leader's barriers) =====
        uint32_t stage = 0, phase = 0, a_gen = 0;
        uint32_t cur_a = 0xFFFFFFFFu;
        for (uint32_t item = item0; item < item1; ++item) {
            const uint32_t mt = item / per_mt, kc = (item % per_mt) / n_blocks, nb = item % n_blocks;
            const bool tail = args.n - nb * 256u < 256u;  // 128-wide block: 64 weight rows per CTA
            const bool new_a = item / n_blocks != cur_a;  // a new (tile, K chunk)
            cur_a = item / n_blocks;
            const uint32_t k0 = kc * n_kb * kMmaKBlock;
            for (uint32_t kb = 0; kb < n_kb; ++kb) {
                if (new_a) {
                    if (a_gen) mbar_wait(aempty_bar(kb), (a_gen - 1u) & 1u);  // the old tile's last item has read this block
                    if (elect_one()) {
                        if (rank == 0) mbar_expect_tx(afull_bar(kb), 2u * kMmaTileBytes);
                        tma_load_2d_pair(a_smem + kb * kMmaTileBytes, &tm_a, afull_bar(kb), (int32_t)(k0 + kb * kMmaKBlock),
                                         (int32_t)(mt * 256u + rank * 128u));
                    }
                    __syncwarp();
                }
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(full_bar(stage), tail ? kMmaTileBytes : 2u * kMmaTileBytes);
                    if (tail)
                        tma_load_2d_pair(w_smem + stage * kMmaTileBytes, &tm_w64, full_bar(stage), (int32_t)(k0 + kb * kMmaKBlock),
                                         (int32_t)(nb * 256u + rank * 64u));
                    else
                        tma_load_2d_pair(w_smem + stage * kMmaTileBytes, &tm_w, full_bar(stage), (int32_t)(k0 + kb * kMmaKBlock),
                                         (int32_t)(nb * 256u + rank * 128u));
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (new_a) ++a_gen;
        }
    } else if (warp == kAresEpiWarps + 1) {
        if (rank == 0) {
            // ===== MMA issuer (leader CTA only) =====
            constexpr uint32_t idesc256 = umma_idesc_f16(256, 256), idesc128 = umma_idesc_f16(256, 128);
            const uint64_t a_desc0 = umma_desc_sw128(a_smem), w_desc0 = umma_desc_sw128(w_smem);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, a_gen = 0;
            uint32_t cur_a = 0xFFFFFFFFu;
            for (uint32_t item = item0; item < item1; ++item) {
                const uint32_t kc = (item % per_mt) / n_blocks, nb = item % n_blocks;
                const bool tail = args.n - nb * 256u < 256u;
                const bool new_a = item / n_blocks != cur_a;
                const bool last_of_a = item + 1 == item1 || (item + 1) / n_blocks != item / n_blocks;
                cur_a = item / n_blocks;
                if (kch > 1) {  // the block's accumulator, through every K chunk of the tile
                    acc = nb;
                    acc_phase = ((item - item0) / per_mt) & 1u;
                }
                if (kc == 0) {
                    mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                    tc_fence_after();
                }
                const uint32_t d_tmem = tmem_base + acc * 256u;
                for (uint32_t kb = 0; kb < n_kb; ++kb) {
                    if (new_a) mbar_wait(afull_bar(kb), a_gen & 1u);
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t a_desc = a_desc0 + (uint64_t)(kb * (kMmaTileBytes >> 4));
                        const uint64_t w_desc = w_desc0 + (uint64_t)(stage * (kMmaTileBytes >> 4));
#pragma unroll
                        for (uint32_t k4 = 0; k4 < 4; ++k4)
                            umma_f16_pair(d_tmem, a_desc + 2u * k4, w_desc + 2u * k4, tail ? idesc128 : idesc256,
                                          (kc | kb | k4) != 0u ? 1u : 0u);
                        umma_commit_pair(empty_bar(stage));
                        if (last_of_a) umma_commit_pair(aempty_bar(kb));
                        if (kb + 1 == n_kb && kc + 1 == kch) umma_commit_pair(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == args.n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (new_a) ++a_gen;
                if (kch == 1) {
                    acc ^= 1u;
                    if (acc == 0u) acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs): TMEM lane = output row of this CTA, column = feature of the block =====
        const uint32_t quarter = warp & 3u, part = warp >> 2;  // columns [part * 64, +64) of the block
        const uint32_t my_stage = stage_smem + warp * 4096u;
        uint8_t* my_ptr = stage_ptr + (size_t)warp * 4096u;
        const uint32_t sw = lane & 7u;
        const int mode = args.mode & 15, dbg = args.mode >> 4;  // dbg: timing experiments (FSGPU_MINILM_DBG)
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t item = item0; item < item1; ++item) {
            const uint32_t mt = item / per_mt, kc = (item % per_mt) / n_blocks, nb = item % n_blocks;
            if (kc + 1 != kch) continue;  // the block is complete after the tile's last K chunk
            if (kch > 1) {
                acc = nb;
                acc_phase = ((item - item0) / per_mt) & 1u;
            }
            const bool tail = args.n - nb * 256u < 256u;
            const uint32_t row0 = mt * 256u + rank * 128u + quarter * 32u;
            const bool idle = tail && part * 64u >= 128u;  // a 128-wide block has columns for two of the four parts
            const uint32_t col0 = nb * 256u + part * 64u;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * 256u + part * 64u;
            uint32_t v[2][32];
            if (!idle) {
                tmem_ld_x32(taddr, v[0]);
                tmem_ld_x32(taddr + 32u, v[1]);
                tmem_ld_wait();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);  // the accumulator is in registers
            if (kch == 1) {
                acc ^= 1u;
                if (acc == 0u) acc_phase ^= 1u;
            }
            if (idle) continue;
            if (row0 >= m_rows) continue;
#pragma unroll
            for (int c = 0; c < 2; ++c) {  // 32-column chunks; a staging box = 64 f16 columns or 32 f32 columns
                const bool new_box = mode == 2 || (c & 1) == 0;
                if (new_box && !(dbg & 2)) {
                    if (lane == 0) tma_store_wait_read();
                    __syncwarp();
                }
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        x[i] = __uint_as_float(v[c][j + i]) + ((dbg & 1) ? 0.f : bias_s[col0 + c * 32 + j + i]);
                        if (mode == 1) x[i] = gelu_tanh_fit(x[i]);
                    }
                    if (dbg & 4) {
                        if (x[0] + x[1] + x[2] + x[3] + x[4] + x[5] + x[6] + x[7] == 12345.678f) my_ptr[0] = 1;
                    } else if (mode == 2) {
                        const uint32_t ch = (uint32_t)j / 4u;
                        *reinterpret_cast<float4*>(my_ptr + lane * 128u + ((ch ^ sw) << 4)) = make_float4(x[0], x[1], x[2], x[3]);
                        *reinterpret_cast<float4*>(my_ptr + lane * 128u + (((ch + 1u) ^ sw) << 4)) = make_float4(x[4], x[5], x[6], x[7]);
                    } else {
                        const uint32_t ch = (uint32_t)((c & 1) * 32 + j) / 8u;
                        __half2 hh[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) hh[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
                        *reinterpret_cast<uint4*>(my_ptr + lane * 128u + ((ch ^ sw) << 4)) = *reinterpret_cast<uint4*>(hh);
                    }
                }
                const bool box_done = mode == 2 || (c & 1) == 1;
                if (box_done && !(dbg & 2)) {
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        const int32_t bc = (int32_t)(col0 + (mode == 2 ? c * 32 : (c & ~1) * 32));
                        tma_store_2d(&tm_out, my_stage, bc, (int32_t)row0);
                        tma_store_commit();
                    }
                }
            }
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ─── FFN-in -> erf-GELU -> FFN-out in ONE kernel (H = 384, I = 1536) ─────────────────────────────────────────
// The [rows x 1536] intermediate never leaves the SM.  A CTA pair owns 256-row tiles (128 rows per CTA: its h tile,
// 96 KiB, resident as the A operand of the first GEMM).  The 1536 intermediate features are walked in 12 chunks of
// 128: G1(c): acc1 [128 x 128] = h . W1[chunk c]^T  (K = 384; cta_group::2, M = 256, N = 128: 64 weight rows per CTA),
// epilogue(c): + bias, GELU, f16, written as the NEXT GEMM's A operand — straight into the 128-byte-swizzled K-major
// layout tcgen05 reads (two [128 x 64] K-blocks per chunk, double-buffered) —, G2(c): acc2 [128 x 384] += g_c . W2[:, chunk c]^T
// (K = 128; three N = 128 instructions per K-block).  The issuer interleaves G1(c+1) before G2(c), so the tensor pipe
// works on one while the sixteen epilogue warps finish the other; acc1 is handed back as soon as it is in registers.
// TMEM: 128 + 384 = 512 columns.  Only weights stream (32 B per MMA clock): every weight box of either matrix is
// [64 rows x 64 K] = 8 KiB per CTA, and one ring slot holds TWO of them behind one barrier (the two K-blocks that
// feed the same accumulator columns), i.e. eight MMAs = 512 tensor clocks per barrier round.  In-kernel timestamps
// (FSGPU_MINILM_FFN_TS) showed why: with one box per barrier the single issuing thread needed ~380 clocks per round
// (wait, fence, election, four MMAs, commits) for 256 clocks of tensor work — the ISSUER, not the epilogue or the
// weight stream, set the pace (tensor pipe 49 % active).  The whole issue loop now runs in one elected thread.
// After the last chunk the epilogue adds the FFN-out bias and stores acc2 as f32 through TMA (staging = the two g buffers).
constexpr int kFfnEpiWarps = 16;
constexpr int kFfnThreads = 32 * (kFfnEpiWarps + 2);
constexpr uint32_t kFfnStages = 4;                     // weight ring slots (three held one GEMM group: every refill waited for L2)
constexpr uint32_t kFfnBoxBytes = kMmaTileBytes / 2;   // [64 x 64] f16
constexpr uint32_t kFfnSlotBytes = 2 * kFfnBoxBytes;   // two boxes per slot
constexpr uint32_t kFfnChunks = 12;                    // 1536 / 128
constexpr uint32_t kFfnHKb = 6;                        // 384 / 64

struct FfnArgs {
    uint32_t m;
    const uint32_t* m_ptr;  // packed rows: the row count lives on the device (m is then the upper bound the grid was sized for)
    const float* bias1;  // [1536]
    const float* bias2;  // [384]
    uint32_t dbg;        // timing experiments (FSGPU_MINILM_FFN_DBG; results are wrong with any bit set): 1 no GELU math,
                         // 2 issuer ignores g_full, 4 G2 MMAs not issued, 8 weight boxes loaded once (no TMA after the
                         // first ring fill), 16 G1 MMAs not issued
    long long* ts;       // FSGPU_MINILM_FFN_TS: clock64 stamps of pair 0's issuer and first epilogue warp, 8 per chunk, 24 chunks
    // residual + LayerNorm in the final epilogue (ln_g != nullptr): h = LayerNorm(acc2 + bias2 + h) written in place as
    // f16 through tm_h_out (and as f32 to h32 when given: the last layer feeds the pooling kernel); nothing goes to tm_out then
    __half* h16;         // [m x 384] the layer input = the residual (read from the resident tile); tm_h_out stores over it
    const float* ln_g;   // [384]
    const float* ln_b;   // [384]
    float* h32;          // [m x 384] or nullptr
    float eps;
};

__host__ __device__ inline size_t ffn_fused_smem_bytes() {
    return 1024 + (size_t)(kFfnHKb + 4) * kMmaTileBytes + (size_t)kFfnStages * kFfnSlotBytes + 512 + 384 * 4;  // = 227 KiB exactly
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFfnThreads, 1)
ffn_fused_pair_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w1_64,
                      const __grid_constant__ CUtensorMap tm_w2_64, const __grid_constant__ CUtensorMap tm_out,
                      const __grid_constant__ CUtensorMap tm_h_out, const FfnArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t h_smem = base;                            // 6 x 16 KiB
    const uint32_t g_smem = base + kFfnHKb * kMmaTileBytes;  // 2 buffers x 2 K-blocks x 16 KiB
    const uint32_t w_smem = g_smem + 4u * kMmaTileBytes;     // ring
    uint8_t* g_ptr = base_ptr + (size_t)kFfnHKb * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + (size_t)(kFfnHKb + 4) * kMmaTileBytes + (size_t)kFfnStages * kFfnSlotBytes);
    const uint32_t bar0 = smem_u32(bars);
    auto w_full = [&](uint32_t s) { return bar0 + 8u * s; };
    auto w_empty = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto h_full = [&](uint32_t kb) { return bar0 + 8u * (16u + kb); };
    auto h_empty = [&](uint32_t kb) { return bar0 + 8u * (22u + kb); };
    const uint32_t acc1_full = bar0 + 8u * 28u, acc1_empty = bar0 + 8u * 29u;
    auto g_full = [&](uint32_t b) { return bar0 + 8u * (30u + b); };
    auto g_empty = [&](uint32_t b) { return bar0 + 8u * (32u + b); };
    const uint32_t acc2_full = bar0 + 8u * 34u, acc2_empty = bar0 + 8u * 35u;
    const uint32_t res_done = bar0 + 8u * 36u;  // fused LayerNorm: this CTA's epilogue warps have read the residual out of the h tile
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 40);
    float* bias2_s = reinterpret_cast<float*>(bars + 64);  // the FFN-in bias has no room here: see the epilogue
    for (uint32_t i = threadIdx.x; i < 384u; i += blockDim.x) bias2_s[i] = args.bias2[i];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t m_rows = args.m_ptr ? min(*args.m_ptr, args.m) : args.m;
    const uint32_t m_tiles = (m_rows + 255u) / 256u;

    if (warp == kFfnEpiWarps && lane == 0) {
        tma_prefetch_desc(&tm_h);
        tma_prefetch_desc(&tm_w1_64);
        tma_prefetch_desc(&tm_w2_64);
        tma_prefetch_desc(&tm_out);
        tma_prefetch_desc(&tm_h_out);
        for (uint32_t s = 0; s < kFfnStages; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), 1);
        }
        for (uint32_t kb = 0; kb < kFfnHKb; ++kb) {
            mbar_init(h_full(kb), 1);
            mbar_init(h_empty(kb), 1);
        }
        mbar_init(acc1_full, 1);
        mbar_init(acc1_empty, 2 * kFfnEpiWarps);
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(g_full(b), 2 * kFfnEpiWarps);
            mbar_init(g_empty(b), 1);
        }
        mbar_init(acc2_full, 1);
        mbar_init(acc2_empty, 2 * kFfnEpiWarps);
        mbar_init(res_done, kFfnEpiWarps);
        fence_barrier_init();
    } else if (warp == 0) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc2_col = 0u, acc1_col = 384u;

    if (warp == kFfnEpiWarps) {
        // ===== TMA producer (both CTAs): h tile per 256-row tile, then the weight sequence G1(0), G1(1), G2(0), G1(2), G2(1), ...
        uint32_t stage = 0, phase = 0, tiles_done = 0, n_loads = 0;
        auto load_w = [&](const CUtensorMap* tm, int32_t kcol, int32_t wrow) {  // two [64 x 64] boxes per CTA: K-columns kcol, kcol + 64
            mbar_wait(w_empty(stage), phase ^ 1u);
            if (elect_one()) {
                if ((args.dbg & 8u) && n_loads >= kFfnStages) {
                    if (rank == 0) mbar_arrive(w_full(stage));
                } else {
                    if (rank == 0) mbar_expect_tx(w_full(stage), 2u * kFfnSlotBytes);
                    tma_load_2d_pair(w_smem + stage * kFfnSlotBytes, tm, w_full(stage), kcol, wrow);
                    tma_load_2d_pair(w_smem + stage * kFfnSlotBytes + kFfnBoxBytes, tm, w_full(stage), kcol + (int32_t)kMmaKBlock, wrow);
                }
            }
            ++n_loads;
            __syncwarp();
            if (++stage == kFfnStages) {
                stage = 0;
                phase ^= 1u;
            }
        };
        auto load_g1 = [&](uint32_t c) {  // W1 rows [128c, +128): 64 per CTA, K-blocks (0,1) (2,3) (4,5)
            for (uint32_t i = 0; i < kFfnHKb / 2; ++i)
                load_w(&tm_w1_64, (int32_t)(2u * i * kMmaKBlock), (int32_t)(c * 128u + rank * 64u));
        };
        auto load_g2 = [&](uint32_t c) {  // W2 columns [128c, +128) (two K-blocks) for each block of 128 output features
            for (uint32_t part = 0; part < 3; ++part)
                load_w(&tm_w2_64, (int32_t)(c * 128u), (int32_t)(part * 128u + rank * 64u));
        };
        auto load_h = [&](uint32_t mt) {  // the tile's h blocks, each as soon as the previous tile's last G1 has read it
            if (tiles_done && args.ln_g) mbar_wait(res_done, (tiles_done - 1u) & 1u);
            for (uint32_t kb = 0; kb < kFfnHKb; ++kb) {
                if (tiles_done) mbar_wait(h_empty(kb), (tiles_done - 1u) & 1u);
                if (elect_one()) {
                    if (rank == 0) mbar_expect_tx(h_full(kb), 2u * kMmaTileBytes);
                    tma_load_2d_pair(h_smem + kb * kMmaTileBytes, &tm_h, h_full(kb), (int32_t)(kb * kMmaKBlock),
                                     (int32_t)(mt * 256u + rank * 128u));
                }
                __syncwarp();
            }
        };
        if (pair < m_tiles) load_h(pair);
        for (uint32_t mt = pair; mt < m_tiles; mt += n_pairs) {
            for (uint32_t c = 0; c < kFfnChunks; ++c) {
                load_g1(c);
                if (c) load_g2(c - 1u);
            }
            ++tiles_done;
            // unfused: before the last G2's weights, so that it lands under G2(10), G2(11); fused LayerNorm: the final
            // epilogue still reads the residual out of the tile (res_done), and that epilogue needs G2(11)'s weights first
            if (!args.ln_g && mt + n_pairs < m_tiles) load_h(mt + n_pairs);
            load_g2(kFfnChunks - 1u);
            if (args.ln_g && mt + n_pairs < m_tiles) load_h(mt + n_pairs);
        }
    } else if (warp == kFfnEpiWarps + 1) {
        if (rank == 0 && elect_one()) {
            // ===== MMA issuer (one thread of the leader CTA) =====
            constexpr uint32_t idesc128 = umma_idesc_f16(256, 128);
            const uint64_t h_desc0 = umma_desc_sw128(h_smem), g_desc0 = umma_desc_sw128(g_smem), w_desc0 = umma_desc_sw128(w_smem);
            const bool ts_on = args.ts != nullptr && blockIdx.x == 0;
            uint32_t stage = 0, phase = 0, tiles_done = 0;
            uint32_t n_g1 = 0;          // G1 groups issued so far (acc1 generation)
            uint32_t n_g2[2] = {0, 0};  // G2 groups issued per g buffer
            auto next_stage = [&]() {
                if (++stage == kFfnStages) {
                    stage = 0;
                    phase ^= 1u;
                }
            };
            auto issue_g1 = [&](uint32_t c, bool first_of_tile) {
                mbar_wait(acc1_empty, (n_g1 & 1u) ^ 1u);
                tc_fence_after();
                if (ts_on && n_g1 < 24) args.ts[n_g1 * 8 + 0] = clock64();
                for (uint32_t i = 0; i < kFfnHKb / 2; ++i) {
                    if (first_of_tile) {
                        mbar_wait(h_full(2u * i), tiles_done & 1u);
                        mbar_wait(h_full(2u * i + 1u), tiles_done & 1u);
                    }
                    mbar_wait(w_full(stage), phase);
                    tc_fence_after();
                    const uint64_t a = h_desc0 + (uint64_t)(2u * i * (kMmaTileBytes >> 4));
                    const uint64_t b = w_desc0 + (uint64_t)(stage * (kFfnSlotBytes >> 4));
                    if (!(args.dbg & 16u)) {
#pragma unroll
                        for (uint32_t half = 0; half < 2; ++half)
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4)
                                umma_f16_pair(tmem_base + acc1_col, a + half * (kMmaTileBytes >> 4) + 2u * k4,
                                              b + half * (kFfnBoxBytes >> 4) + 2u * k4, idesc128, (i | half | k4) != 0u ? 1u : 0u);
                    }
                    umma_commit_pair(w_empty(stage));
                    if (c + 1 == kFfnChunks) {  // the tile's last use of these h blocks
                        umma_commit_pair(h_empty(2u * i));
                        umma_commit_pair(h_empty(2u * i + 1u));
                    }
                    if (i + 1 == kFfnHKb / 2) umma_commit_pair(acc1_full);
                    next_stage();
                }
                if (ts_on && n_g1 < 24) args.ts[n_g1 * 8 + 1] = clock64();
                ++n_g1;
            };
            auto issue_g2 = [&](uint32_t c) {
                const uint32_t buf = c & 1u;
                if (!(args.dbg & 2u)) mbar_wait(g_full(buf), n_g2[buf] & 1u);
                tc_fence_after();
                const uint32_t gi = tiles_done * kFfnChunks + c;
                if (ts_on && gi < 24) args.ts[gi * 8 + 2] = clock64();
                if (c == 0) {
                    mbar_wait(acc2_empty, (tiles_done & 1u) ^ 1u);
                    tc_fence_after();
                }
                const uint64_t a = g_desc0 + (uint64_t)(buf * 2u * (kMmaTileBytes >> 4));
                for (uint32_t part = 0; part < 3; ++part) {  // output features [128 part, +128)
                    mbar_wait(w_full(stage), phase);
                    tc_fence_after();
                    const uint64_t b = w_desc0 + (uint64_t)(stage * (kFfnSlotBytes >> 4));
                    const uint32_t d = tmem_base + acc2_col + part * 128u;
                    if (!(args.dbg & 4u)) {
#pragma unroll
                        for (uint32_t j = 0; j < 2; ++j)
#pragma unroll
                            for (uint32_t k4 = 0; k4 < 4; ++k4)
                                umma_f16_pair(d, a + j * (kMmaTileBytes >> 4) + 2u * k4, b + j * (kFfnBoxBytes >> 4) + 2u * k4, idesc128,
                                              (c | j | k4) != 0u ? 1u : 0u);
                    }
                    umma_commit_pair(w_empty(stage));
                    if (part == 2) {
                        umma_commit_pair(g_empty(buf));
                        if (c + 1 == kFfnChunks) umma_commit_pair(acc2_full);
                    }
                    next_stage();
                }
                if (ts_on && gi < 24) args.ts[gi * 8 + 3] = clock64();
                ++n_g2[buf];
            };
            for (uint32_t mt = pair; mt < m_tiles; mt += n_pairs, ++tiles_done) {
                for (uint32_t c = 0; c < kFfnChunks; ++c) {
                    issue_g1(c, c == 0);
                    if (c) issue_g2(c - 1u);
                }
                issue_g2(kFfnChunks - 1u);
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue (both CTAs): TMEM lane = row of this CTA =====
        const uint32_t quarter = warp & 3u, part = warp >> 2;  // acc1: columns [32 part, +32); acc2: [96 part, +96)
        const uint32_t row_l = quarter * 32u + lane;           // local row 0..127
        const uint32_t sw = row_l & 7u;
        uint32_t n_acc1 = 0, n_g[2] = {0, 0}, tiles_done = 0;
        const bool ts_on = args.ts != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
        for (uint32_t mt = pair; mt < m_tiles; mt += n_pairs, ++tiles_done) {
            for (uint32_t c = 0; c < kFfnChunks; ++c) {
                const uint32_t buf = c & 1u;
                // this warp's 32 FFN-in bias values of the chunk (warp-uniform addresses), fetched ahead of the wait
                float4 bq[8];
                const float4* bsrc = reinterpret_cast<const float4*>(args.bias1 + c * 128u + part * 32u);
#pragma unroll
                for (int i = 0; i < 8; ++i) bq[i] = __ldg(bsrc + i);
                const float* bias1_r = reinterpret_cast<const float*>(bq);
                mbar_wait(acc1_full, n_acc1 & 1u);
                tc_fence_after();
                if (ts_on && n_acc1 < 24) args.ts[n_acc1 * 8 + 4] = clock64();
                uint32_t v[32];
                tmem_ld_x32(tmem_base + ((quarter * 32u) << 16) + acc1_col + part * 32u, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(acc1_empty, 0);
                if (ts_on && n_acc1 < 24) args.ts[n_acc1 * 8 + 5] = clock64();
                ++n_acc1;
                // the g buffer is free once G2 of the chunk that used it last has completed
                if (n_g[buf]) mbar_wait(g_empty(buf), (n_g[buf] - 1u) & 1u);
                if (ts_on && n_acc1 <= 24) args.ts[(n_acc1 - 1) * 8 + 6] = clock64();
                uint8_t* gk = g_ptr + (size_t)(buf * 2u + (part >> 1)) * kMmaTileBytes + row_l * 128u;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    __half2 hh[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float x0 = __uint_as_float(v[j + 2 * i]) + bias1_r[j + 2 * i];
                        float x1 = __uint_as_float(v[j + 2 * i + 1]) + bias1_r[j + 2 * i + 1];
                        if (!(args.dbg & 1u)) {
                            x0 = gelu_tanh_fit(x0);
                            x1 = gelu_tanh_fit(x1);
                        }
                        hh[i] = __floats2half2_rn(x0, x1);
                    }
                    const uint32_t ch = (part & 1u) * 4u + (uint32_t)j / 8u;
                    *reinterpret_cast<uint4*>(gk + ((ch ^ sw) << 4)) = *reinterpret_cast<uint4*>(hh);
                }
                fence_proxy_async_smem();  // generic-proxy writes -> tcgen05 operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(g_full(buf), 0);
                if (ts_on && n_acc1 <= 24) args.ts[(n_acc1 - 1) * 8 + 7] = clock64();
                ++n_g[buf];
            }
            mbar_wait(acc2_full, tiles_done & 1u);
            tc_fence_after();
            if (args.ln_g) {
                // final, fused form: x = acc2 + bias2 + residual, LayerNorm over the row's 384 features (two-pass statistics
                // in f32 as the stand-alone kernel computes them), h written in place.  A row's features sit in four threads
                // (same lane, the four column parts): partial sums meet in shared memory (the g buffers are free: acc2_full
                // implies every G2 of the tile has completed).  Pass 1 reads the residual out of the resident h tile (the
                // swizzled A operand), writes x BACK into the accumulator's TMEM columns and sums it; the h tile is then
                // released to the producer (res_done); passes 2 (variance) and 3 (output) read x from TMEM only — 96 values
                // per thread never sit in registers.  (Residual reads from global memory instead: 36 us per launch, latency.)
                float* red = reinterpret_cast<float*>(g_ptr);  // [2][128 rows][4 parts]
                float* gb_s = red + 1024;                      // gamma [384] | beta [384]
                const uint32_t et = threadIdx.x;               // 0 .. 511 (the epilogue warps are warps 0 .. 15)
                if (et < 384u) gb_s[et] = args.ln_g[et];  // 512 threads stage the 768 values in two steps
                else gb_s[et] = args.ln_b[et - 384u];
                if (et < 256u) gb_s[512u + et] = args.ln_b[128u + et];
                const uint32_t row = mt * 256u + rank * 128u + row_l;
                const bool row_ok = row < m_rows;
                const uint32_t tcol = tmem_base + ((quarter * 32u) << 16) + acc2_col + part * 96u;
                const uint8_t* h_row = base_ptr + row_l * 128u;  // this row inside every [128 x 64] K-block of the h tile
                float acc = 0.0f;
#pragma unroll 1
                for (uint32_t b = 0; b < 3; ++b) {
                    uint32_t w2[32];
                    tmem_ld_x32(tcol + b * 32u, w2);
                    uint4 r4[4];
#pragma unroll
                    for (uint32_t i = 0; i < 4; ++i) {
                        const uint32_t col0 = part * 96u + b * 32u + i * 8u;
                        r4[i] = *reinterpret_cast<const uint4*>(h_row + (size_t)(col0 >> 6) * kMmaTileBytes + ((((col0 >> 3) & 7u) ^ sw) << 4));
                    }
                    tmem_ld_wait();
                    const __half2* rh = reinterpret_cast<const __half2*>(r4);
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const float2 r = __half22float2(rh[j / 2]);
                        const uint32_t col = part * 96u + b * 32u + (uint32_t)j;
                        const float x0 = __uint_as_float(w2[j]) + bias2_s[col] + r.x;
                        const float x1 = __uint_as_float(w2[j + 1]) + bias2_s[col + 1u] + r.y;
                        acc += x0;
                        acc += x1;
                        w2[j] = __float_as_uint(x0);
                        w2[j + 1] = __float_as_uint(x1);
                    }
                    tmem_st_x32(tcol + b * 32u, w2);
                }
                tmem_st_wait();
                __syncwarp();
                if (lane == 0) mbar_arrive(res_done);  // the residual has been read: the producer may load the next h tile
                red[row_l * 4u + part] = acc;
                asm volatile("bar.sync 1, %0;" ::"r"(kFfnEpiWarps * 32) : "memory");
                const float mean = ((red[row_l * 4u] + red[row_l * 4u + 1u]) + (red[row_l * 4u + 2u] + red[row_l * 4u + 3u])) * (1.0f / 384.0f);
                acc = 0.0f;
#pragma unroll 1
                for (uint32_t b = 0; b < 3; ++b) {
                    uint32_t w2[32];
                    tmem_ld_x32(tcol + b * 32u, w2);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float d = __uint_as_float(w2[j]) - mean;
                        acc = fmaf(d, d, acc);
                    }
                }
                red[512u + row_l * 4u + part] = acc;
                asm volatile("bar.sync 1, %0;" ::"r"(kFfnEpiWarps * 32) : "memory");
                const float var_eps = ((red[512u + row_l * 4u] + red[512u + row_l * 4u + 1u]) + (red[512u + row_l * 4u + 2u] + red[512u + row_l * 4u + 3u])) *
                                          (1.0f / 384.0f) + args.eps;
                const float inv = rsqrtf(var_eps);
                const float inv2 = inv * (1.5f - 0.5f * var_eps * inv * inv);  // one Newton step (as layernorm_row)
                // output: twelve warps (column parts 0..2) take two [32 rows x 64 features] boxes each — f16 rows of 128 bytes,
                // written swizzled into 4 KiB of the g buffers and stored by TMA (full row segments; 16-byte stores of one
                // row per lane cost 6 k LSU cycles per tile); part 3 only hands the accumulator back
                const uint32_t trow = tmem_base + ((quarter * 32u) << 16) + acc2_col;
                if (part == 3u) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc2_empty, 0);
                } else {
                    uint8_t* st_ptr = g_ptr + 8192u + (size_t)(quarter * 3u + part) * 4096u;
                    const uint32_t st_smem = g_smem + 8192u + (quarter * 3u + part) * 4096u;
                    const uint32_t row0 = mt * 256u + rank * 128u + quarter * 32u;
#pragma unroll 1
                    for (uint32_t bi = 0; bi < 2; ++bi) {
                        const uint32_t c0 = (2u * part + bi) * 64u;  // first feature of the box
                        if (lane == 0) tma_store_wait_read();       // the previous box has left the staging buffer
                        __syncwarp();
#pragma unroll 1
                        for (uint32_t hf = 0; hf < 2; ++hf) {
                            uint32_t w2[32];
                            tmem_ld_x32(trow + c0 + hf * 32u, w2);
                            tmem_ld_wait();
                            if (bi == 1 && hf == 1) {  // last read of the accumulator: hand it back
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive_cluster(acc2_empty, 0);
                            }
                            const float* gs = gb_s + c0 + hf * 32u;
                            float x[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] = (__uint_as_float(w2[j]) - mean) * inv2 * gs[j] + gs[384 + j];
#pragma unroll
                            for (uint32_t i = 0; i < 4; ++i) {
                                __half2 hh[4];
#pragma unroll
                                for (uint32_t q = 0; q < 4; ++q) hh[q] = __floats2half2_rn(x[8 * i + 2 * q], x[8 * i + 2 * q + 1]);
                                *reinterpret_cast<uint4*>(st_ptr + lane * 128u + (((hf * 4u + i) ^ (lane & 7u)) << 4)) = *reinterpret_cast<const uint4*>(hh);
                            }
                            if (args.h32 && row_ok) {
                                float* o32 = args.h32 + (size_t)row * 384u + c0 + hf * 32u;
#pragma unroll
                                for (int i = 0; i < 8; ++i)
                                    *reinterpret_cast<float4*>(o32 + 4 * i) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                            }
                        }
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (lane == 0 && row0 < m_rows) {
                            tma_store_2d(&tm_h_out, st_smem, (int32_t)c0, (int32_t)row0);  // rows past m are clipped by the map
                            tma_store_commit();
                        }
                    }
                    if (lane == 0) tma_store_wait_read();
                }
                // every warp's reads of the statistics must be over before any warp writes g values of the next tile over them
                asm volatile("bar.sync 1, %0;" ::"r"(kFfnEpiWarps * 32) : "memory");
                continue;
            }
            // final, unfused form: acc2 + bias2 -> f32 pre, 96 columns per warp as three [32 x 32] boxes through this warp's 4 KiB of
            // the g buffers (free: acc2_full implies every G2 of the tile has completed)
            const uint32_t row0 = mt * 256u + rank * 128u + quarter * 32u;
            uint8_t* st_ptr = g_ptr + (size_t)warp * 4096u;
            const uint32_t st_smem = g_smem + warp * 4096u;
            const uint32_t swl = lane & 7u;
#pragma unroll 1
            for (uint32_t b = 0; b < 3; ++b) {  // one [32 x 32] box at a time (32 accumulator registers, not 96)
                uint32_t w2[32];
                tmem_ld_x32(tmem_base + ((quarter * 32u) << 16) + acc2_col + part * 96u + b * 32u, w2);
                tmem_ld_wait();
                if (b == 2) {  // the accumulator is in registers / already stored: hand it back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc2_empty, 0);
                }
                if (row0 >= m_rows) continue;
                if (lane == 0) tma_store_wait_read();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const uint32_t col = part * 96u + b * 32u + (uint32_t)j;
                    const float4 x = make_float4(__uint_as_float(w2[j]) + bias2_s[col], __uint_as_float(w2[j + 1]) + bias2_s[col + 1],
                                                 __uint_as_float(w2[j + 2]) + bias2_s[col + 2], __uint_as_float(w2[j + 3]) + bias2_s[col + 3]);
                    *reinterpret_cast<float4*>(st_ptr + lane * 128u + ((((uint32_t)j / 4u) ^ swl) << 4)) = x;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm_out, st_smem, (int32_t)(part * 96u + b * 32u), (int32_t)row0);
                    tma_store_commit();
                }
            }
            if (lane == 0) tma_store_wait_read();
            // every warp's staging must have been read before any warp writes g values of the next tile over it
            asm volatile("bar.sync 1, %0;" ::"r"(kFfnEpiWarps * 32) : "memory");
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ─── h = LayerNorm(pre + residual): one warp per token row of H = 384 ───────────────────────
// pre f32 (linear + bias), residual f16 (the layer input); h f16 in place of the residual, and an f32 copy on
// request (the last layer: the pooling kernel reads f32).
__global__ void __launch_bounds__(256)
minilm_fast_ln_kernel(const float* __restrict__ pre, __half* __restrict__ h, size_t rows, const float* __restrict__ g,
                      const float* __restrict__ b, float eps, float* __restrict__ out_f32, const uint32_t* __restrict__ m_ptr) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows || (m_ptr && row >= *m_ptr)) return;
    float x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const size_t o = row * kHidden + lane + 32u * i;
        x[i] = pre[o] + __half2float(h[o]);
    }
    layernorm_row(x, g, b, eps, lane);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const size_t o = row * kHidden + lane + 32u * i;
        h[o] = __float2half_rn(x[i]);
        if (out_f32) out_f32[o] = x[i];
    }
}

// Packed rows (f16 form): sequence b owns rows [offs[b], offs[b] + len_b) — no padding rows exist, so every row-wise
// kernel and GEMM of the forward works on sum(len) rows instead of batch * t_pad (queries are short and ragged).
// offs[b] = sum of min(max(lens[i], 0), t_pad) for i < b; offs[batch] = the row count every later kernel reads.
__global__ void __launch_bounds__(1024) minilm_offsets_kernel(const int32_t* __restrict__ lens, uint32_t batch, uint32_t t_pad,
                                                              uint32_t* __restrict__ offs) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < batch; base += 1024u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < batch ? min((uint32_t)max(lens[i], 0), t_pad) : 0u;
        uint32_t x = v;
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
            if ((int)lane >= o) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if ((int)lane >= o) w += y;
            }
            warp_sums[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const uint32_t before = carry + (warp ? warp_sums[warp - 1] : 0u) + x - v;  // exclusive prefix
        if (i < batch) offs[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offs[batch] = carry;
}

// embeddings -> LayerNorm -> h f16 (BertEmbeddings; the f32 sums are those of minilm_embed_kernel)
__global__ void __launch_bounds__(256)
minilm_fast_embed_kernel(const int32_t* __restrict__ ids, uint32_t batch, uint32_t t_pad, uint32_t vocab,
                         const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
                         const float* __restrict__ g, const float* __restrict__ b, float eps, __half* __restrict__ h,
                         const uint32_t* __restrict__ offs) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (size_t)batch * t_pad) return;
    uint32_t t = (uint32_t)(row % t_pad);
    size_t src = row;
    if (offs) {  // packed rows: row -> (sequence, token) through the prefix sums of the lengths
        if (row >= offs[batch]) return;
        uint32_t lo = 0, hi = batch;  // the last sequence with offs[seq] <= row (empty sequences share an offset: take the last)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (offs[mid] <= row) lo = mid;
            else hi = mid;
        }
        t = (uint32_t)row - offs[lo];
        src = (size_t)lo * t_pad + t;
    }
    int32_t id = ids[src];
    if (id < 0 || (uint32_t)id >= vocab) id = 0;
    float x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint32_t d = lane + 32u * i;
        x[i] = word[(size_t)id * kHidden + d] + pos[(size_t)t * kHidden + d] + type0[d];
    }
    layernorm_row(x, g, b, eps, lane);
#pragma unroll
    for (int i = 0; i < 12; ++i) h[row * kHidden + lane + 32u * i] = __float2half_rn(x[i]);
}

// ─── attention for t_pad <= 32 on mma.sync: one warp per (sequence, head) ───────────────────
// qkv [M, 1152] f16 = [q | k | v], head h at columns h * 32.  S = Q K^T / sqrt(32) (2 x 4 m16n8k16 tiles, K = 32),
// keys >= len masked, soft-max in f32 on the accumulator fragments, O = P V with P re-used from the
// accumulator registers as the A operand (f16), O / row sum -> ctx [M, 384] f16.  Tiles are staged in shared
// memory with an 80-byte row pitch (conflict-free ldmatrix).
constexpr int kAttPitch = 40;  // halves per staged row (32 + 8 padding)

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(128)
minilm_fast_attention_kernel(const __half* __restrict__ qkv, const int32_t* __restrict__ lens, uint32_t batch,
                             uint32_t t_pad, __half* __restrict__ ctx, const uint32_t* __restrict__ offs) {
    __shared__ __align__(16) __half tiles[4][3][32 * kAttPitch];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t item = blockIdx.x * 4u + warp;
    if (item >= batch * kHeads) return;
    const uint32_t b = item / kHeads, h = item % kHeads;
    const uint32_t len = min((uint32_t)max(lens[b], 0), t_pad);
    const size_t row0 = offs ? (size_t)offs[b] : (size_t)b * t_pad;
    const uint32_t t_rows = offs ? len : t_pad;  // packed rows: the sequence owns exactly len rows (the next ones are another sequence's)
    __half* qs = tiles[warp][0];
    __half* ks = tiles[warp][1];
    __half* vs = tiles[warp][2];
    // stage: 3 matrices x 32 rows x 4 chunks of 16 bytes; rows >= t_pad are zero
#pragma unroll
    for (int it = 0; it < 12; ++it) {
        const uint32_t idx = (uint32_t)it * 32u + lane;  // 0 .. 383
        const uint32_t mat = idx / 128u, r = (idx % 128u) / 4u, ch = idx % 4u;
        uint4 val = make_uint4(0u, 0u, 0u, 0u);
        if (r < t_rows)
            val = *reinterpret_cast<const uint4*>(qkv + (row0 + r) * (3 * kHidden) + mat * kHidden + h * kHeadDim + ch * 8u);
        *reinterpret_cast<uint4*>(tiles[warp][mat] + r * kAttPitch + ch * 8u) = val;
    }
    __syncwarp();
    const uint32_t g = lane >> 2, tig = lane & 3u;
    // S = Q K^T
    float s[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s[mt][nt][i] = 0.0f;
    uint32_t kf[4][4];  // per key tile nt: b0/b1 of k-step 0, b0/b1 of k-step 1
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
        ldmatrix_x4(kf[nt], smem_u32(ks + (nt * 8 + (lane & 7u)) * kAttPitch + (lane >> 3) * 8u));
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int kstep = 0; kstep < 2; ++kstep) {
            uint32_t a[4];
            ldmatrix_x4(a, smem_u32(qs + (mt * 16 + (lane & 7u) + ((lane >> 3) & 1u) * 8u) * kAttPitch + kstep * 16 + (lane >> 4) * 8u));
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_16816(s[mt][nt], a, kf[nt][2 * kstep], kf[nt][2 * kstep + 1]);
        }
    }
    // soft-max over the keys (columns): thread holds rows g (c0, c1) and g + 8 (c2, c3) of each m tile
    const float scale = 0.17677669529663687f;  // 1 / sqrt(32)
    uint32_t p[2][2][4];                         // P as A fragments: [m tile][key step of 16][a0..a3]
    float inv_sum[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {  // row g (hh = 0) or g + 8 (hh = 1)
            float mx = -INFINITY;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const uint32_t key = (uint32_t)nt * 8u + 2u * tig + (uint32_t)i;
                    float x = s[mt][nt][2 * hh + i] * scale;
                    x = key < len ? x : -INFINITY;
                    s[mt][nt][2 * hh + i] = x;
                    mx = fmaxf(mx, x);
                }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            float sum = 0.0f;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float e = __expf(s[mt][nt][2 * hh + i] - mx);  // len >= 1: mx is finite
                    s[mt][nt][2 * hh + i] = e;
                    sum += e;
                }
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            inv_sum[mt][hh] = 1.0f / sum;
        }
#pragma unroll
        for (int kt = 0; kt < 2; ++kt) {
            p[mt][kt][0] = pack_half2(s[mt][2 * kt][0], s[mt][2 * kt][1]);
            p[mt][kt][1] = pack_half2(s[mt][2 * kt][2], s[mt][2 * kt][3]);
            p[mt][kt][2] = pack_half2(s[mt][2 * kt + 1][0], s[mt][2 * kt + 1][1]);
            p[mt][kt][3] = pack_half2(s[mt][2 * kt + 1][2], s[mt][2 * kt + 1][3]);
        }
    }
    // O = P V
    float o[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) o[mt][nt][i] = 0.0f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {  // 8 output dims per tile
        uint32_t vf[4];                // b0/b1 of key step 0, b0/b1 of key step 1 (transposed 8 x 8 blocks of V)
        ldmatrix_x4_trans(vf, smem_u32(vs + lane * kAttPitch + nt * 8));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            mma_16816(o[mt][nt], p[mt][0], vf[0], vf[1]);
            mma_16816(o[mt][nt], p[mt][1], vf[2], vf[3]);
        }
    }
    // O / sum -> staged rows (over the Q tile) -> 64-byte row segments of ctx
    __syncwarp();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const uint32_t d = (uint32_t)nt * 8u + 2u * tig;
            *reinterpret_cast<uint32_t*>(qs + (mt * 16 + g) * kAttPitch + d) =
                pack_half2(o[mt][nt][0] * inv_sum[mt][0], o[mt][nt][1] * inv_sum[mt][0]);
            *reinterpret_cast<uint32_t*>(qs + (mt * 16 + g + 8) * kAttPitch + d) =
                pack_half2(o[mt][nt][2] * inv_sum[mt][1], o[mt][nt][3] * inv_sum[mt][1]);
        }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const uint32_t idx = (uint32_t)it * 32u + lane, r = idx / 4u, ch = idx % 4u;
        if (r < t_rows)
            *reinterpret_cast<uint4*>(ctx + (row0 + r) * kHidden + h * kHeadDim + ch * 8u) =
                *reinterpret_cast<const uint4*>(qs + r * kAttPitch + ch * 8u);
    }
}

}  // namespace fsgpu
