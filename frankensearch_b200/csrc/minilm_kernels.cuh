// minilm_kernels.cuh — the MiniLM-L6-v2 query encoder on sm_100a.
//
// Reference contract: FastEmbedEmbedder (crates/frankensearch-embed/src/fastembed_embedder.rs:317-353,
// :416-426) = BERT encoder (all-MiniLM-L6-v2: H = 384, L = 6, 12 heads x 32, FFN 1536, erf-GELU,
// post-LN, eps 1e-12 — the architecture stated in-repo at crates/frankensearch-rerank/src/native.rs:36-45,
// :366-432, :587-626) -> attention-mask mean pool -> L2 (eps 1e-12) -> adapter L2 with a zero-vector
// guard (crates/frankensearch-embed/src/model_manifest.rs:300-304).  The arithmetic itself lives in
// un-vendored third-party code (fastembed 6.0.0 -> ort -> ONNX Runtime) and the reference pins its
// values only through a SHA-256 digest, so VALUE PARITY IS UNPINNED: the checker is a PyTorch f32
// BertModel with the same (seeded) weights, tolerance 1e-3 on the 384 outputs (tests/test_gpu_minilm.py).
//
// Layout: the batch is flattened to M = batch * t_pad token rows.  Every linear layer is ONE
// persistent tcgen05 GEMM  C[M, N] = A[M, K] * W[N, K]^T  (PyTorch's [out, in] weight layout is
// already K-major): TMA -> mbarrier ring -> tcgen05.mma.kind::f16 -> TMEM -> fused epilogue
// (bias, residual, erf-GELU, f32 and/or split-f16 outputs).  To keep f32-level accuracy on f16
// tensor cores each f32 operand is carried as hi + lo f16 halves and three products are
// accumulated (A_hi W_hi + A_lo W_hi + A_hi W_lo; the dropped lo*lo term is < 2^-22 relative).
// Attention itself (12 heads x 32 dims, T <= 512 keys) is < 3 % of the FLOPs at query lengths
// and runs on CUDA cores in f32; LayerNorm / embedding / pooling are row-wise f32 kernels.
#pragma once

#include "fsgpu_common.cuh"
#include "tc_ptx.cuh"

namespace fsgpu {

constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quarter)
constexpr int kGemmTileM = 128, kGemmTileN = 128;
constexpr int kGemmAccStages = 2;

struct GemmArgs {
    uint32_t m, n, k;          // n % 128 == 0, k % 64 == 0
    uint32_t n_stages;
    uint32_t products;         // 3: hi/lo split of both operands, 1: hi halves only
    const float* bias;         // [n]
    const float* residual;     // [m, n] or nullptr
    float* out_f32;            // [m, n] or nullptr
    __half* out_hi;            // [m, n] or nullptr   (split-f16 copy of the result for the next GEMM)
    __half* out_lo;
    int gelu;                  // erf-GELU after bias
};

__host__ __device__ inline size_t gemm_smem_bytes(uint32_t n_stages, uint32_t products) {
    return 1024 + (size_t)n_stages * (products == 3 ? 4 : 2) * kMmaTileBytes + 256 + 8 * 16 * 36 * 4;
}

__device__ __forceinline__ void split_f16(float v, __half& hi, __half& lo) {
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// 0.5 x (1 + erf(x / sqrt 2)) with erf by Abramowitz-Stegun 7.1.26 — the form the reference's own
// native MiniLM uses (crates/frankensearch-rerank/src/native.rs:170-186): |erf error| <= 1.5e-7 plus
// ~2e-7 from the fast reciprocal / exp2, i.e. <= 2e-7 |x| on the activation.  ~18 instructions
// against ~30 for erff(): the FFN-in epilogue evaluates 50 M of these per layer at 1024 x 32 tokens.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float e = 1.0f - p * t * __expf(-z * z);  // erf(|x| / sqrt 2)
    return 0.5f * x * (1.0f + copysignf(e, x));
}

// Epilogue of one warp over its [32 rows x n_cols] part of an accumulator tile.  TMEM hands each
// lane one ROW; storing in that form issues 32-sector requests (one 16-byte piece per row) and the
// stores, not the MMAs, bound the GEMM (no-store experiment: 3.36 -> 2.00 ms per 1024 x 32-token
// batch).  So each 32 x 32 chunk goes through a padded shared-memory tile, 16 rows at a time, and
// comes back with 8 lanes per row: bias / residual / outputs then move as float4s of full 128-byte
// row segments, 4 rows per instruction.  Residual and bias for the chunk are in flight before the
// accumulator chunk is consumed (the L1 is a few KiB next to a 227 KiB TMA ring: every global
// load is an L2 round trip).
constexpr int kGemmXposePitch = 36;                            // floats: 16-byte aligned, conflict-free
constexpr int kGemmXposeFloats = 16 * kGemmXposePitch;         // per epilogue warp
__device__ __forceinline__ void gemm_epilogue_warp(const GemmArgs& args, uint32_t taddr, uint32_t row_base,
                                                   uint32_t col0, uint32_t n_cols, float* tile, uint32_t lane) {
    const uint32_t cq = lane & 7u, rq = lane >> 3;  // float4 column chunk / row within a 4-row step
#pragma unroll 1
    for (uint32_t c = 0; c < n_cols / 32; ++c) {
        const uint32_t col = col0 + c * 32u + 4u * cq;
        const float4 bias4 = *reinterpret_cast<const float4*>(args.bias + col);
        float4 res[8];
        if (args.residual) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t r = row_base + (uint32_t)(j >> 2) * 16u + (uint32_t)(j & 3) * 4u + rq;
                res[j] = r < args.m ? *reinterpret_cast<const float4*>(args.residual + (size_t)r * args.n + col)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        uint32_t v[32];
        tmem_ld_x32(taddr + c * 32u, v);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if ((int)(lane >> 4) == h) {
                float4* dst = reinterpret_cast<float4*>(tile + (lane & 15u) * kGemmXposePitch);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                         __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
            }
            __syncwarp();
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const uint32_t rl = (uint32_t)it * 4u + rq;
                const uint32_t r = row_base + (uint32_t)h * 16u + rl;
                float4 x = *reinterpret_cast<const float4*>(tile + rl * kGemmXposePitch + 4u * cq);
                x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
                if (args.residual) {
                    const float4 s4 = res[h * 4 + it];
                    x.x += s4.x; x.y += s4.y; x.z += s4.z; x.w += s4.w;
                }
                if (args.gelu) {
                    x.x = gelu_erf(x.x); x.y = gelu_erf(x.y); x.z = gelu_erf(x.z); x.w = gelu_erf(x.w);
                }
                if (r < args.m) {
                    const size_t o = (size_t)r * args.n + col;
                    if (args.out_f32) *reinterpret_cast<float4*>(args.out_f32 + o) = x;
                    if (args.out_hi) {
                        __half hh[4], ll[4];
                        split_f16(x.x, hh[0], ll[0]); split_f16(x.y, hh[1], ll[1]);
                        split_f16(x.z, hh[2], ll[2]); split_f16(x.w, hh[3], ll[3]);
                        *reinterpret_cast<uint2*>(args.out_hi + o) = *reinterpret_cast<uint2*>(hh);
                        *reinterpret_cast<uint2*>(args.out_lo + o) = *reinterpret_cast<uint2*>(ll);
                    }
                }
            }
            __syncwarp();
        }
    }
}

// Persistent GEMM: CTA c computes output tiles c, c + grid, ... (tile = m_tile * tiles_n + n_tile).
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16split_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                     const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                     const GemmArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t per_stage = (args.products == 3 ? 4u : 2u) * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + (size_t)args.n_stages * per_stage);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (20u + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles_m = (args.m + kGemmTileM - 1) / kGemmTileM, tiles_n = args.n / kGemmTileN;
    const uint32_t n_tiles = tiles_m * tiles_n, n_kb = args.k / kMmaKBlock;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_w_hi);
        for (uint32_t s = 0; s < args.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < kGemmAccStages; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), kGemmAccStages * kGemmTileN);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        uint32_t stage = 0, phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int32_t row_a = (int32_t)((t / tiles_n) * kGemmTileM), row_w = (int32_t)((t % tiles_n) * kGemmTileN);
            for (uint32_t kb = 0; kb < n_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    const uint32_t s0 = base + stage * per_stage;
                    const int32_t kc = (int32_t)(kb * kMmaKBlock);
                    mbar_expect_tx(full_bar(stage), per_stage);
                    tma_load_2d(s0, &tm_a_hi, full_bar(stage), kc, row_a);
                    tma_load_2d(s0 + kMmaTileBytes, &tm_w_hi, full_bar(stage), kc, row_w);
                    if (args.products == 3) {
                        tma_load_2d(s0 + 2 * kMmaTileBytes, &tm_a_lo, full_bar(stage), kc, row_a);
                        tma_load_2d(s0 + 3 * kMmaTileBytes, &tm_w_lo, full_bar(stage), kc, row_w);
                    }
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc_f16(kGemmTileM, kGemmTileN);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * kGemmTileN;
            for (uint32_t kb = 0; kb < n_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t s0 = base + stage * per_stage;
                    const uint64_t a_hi = umma_desc_sw128(s0), w_hi = umma_desc_sw128(s0 + kMmaTileBytes);
                    const uint64_t a_lo = umma_desc_sw128(s0 + 2 * kMmaTileBytes);
                    const uint64_t w_lo = umma_desc_sw128(s0 + 3 * kMmaTileBytes);
#pragma unroll
                    for (uint32_t k4 = 0; k4 < kMmaKBlock / 16; ++k4)
                        umma_f16(d_tmem, a_hi + 2u * k4, w_hi + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                    if (args.products == 3) {
#pragma unroll
                        for (uint32_t k4 = 0; k4 < kMmaKBlock / 16; ++k4) {
                            umma_f16(d_tmem, a_lo + 2u * k4, w_hi + 2u * k4, idesc, 1u);
                            umma_f16(d_tmem, a_hi + 2u * k4, w_lo + 2u * k4, idesc, 1u);
                        }
                    }
                    umma_commit(empty_bar(stage));
                    if (kb + 1 == n_kb) umma_commit(tfull_bar(acc));
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (++acc == kGemmAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    } else {
        // ===== epilogue: TMEM lane = output row, column = output feature =====
        const uint32_t quarter = warp & 3u;
        const uint32_t half = (warp - 2u) >> 2;  // which 64 of the tile's 128 columns this warp stores
        float* tile = reinterpret_cast<float*>(bars + 32) + (warp - 2u) * kGemmXposeFloats;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const uint32_t row_base = (t / tiles_n) * kGemmTileM + quarter * 32u;
            const uint32_t col0 = (t % tiles_n) * kGemmTileN + half * (kGemmTileN / 2);
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kGemmTileN + half * (kGemmTileN / 2);
            gemm_epilogue_warp(args, taddr, row_base, col0, kGemmTileN / 2, tile, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == kGemmAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kGemmAccStages * kGemmTileN);
    }
}

// ─── the same GEMM on CTA pairs ─────────────────────────────────────────────────────────────
// With K = 384 and 128 x 128 tiles the single-CTA GEMM re-reads ~21 GB of A/W tiles from L2 per
// 1024-query batch and sits at the L2->SM fabric limit (22-62 % tensor-pipe active).  Here two CTAs
// of a cluster share one tcgen05.mma.cta_group::2: M = 256 output rows (128 per CTA: its A tile in
// its own shared memory, its accumulators in its own TMEM) x N = 256 features (each CTA TMA-loads
// 128 weight rows), i.e. half the L2 traffic per flop.  A 128-wide tail tile (N % 256 == 128) uses
// the N = 128 instruction shape with 64 weight rows per CTA.  Barrier protocol as in
// mma_scan_pair_kernel: `full` / `tmem_empty` on the leader, `empty` / `tmem_full` multicast.
constexpr int kGemmPairN = 256;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemmThreads, 1)
gemm_f16split_pair_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                          const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
                          const __grid_constant__ CUtensorMap tm_w64_hi, const __grid_constant__ CUtensorMap tm_w64_lo,
                          const GemmArgs args) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t per_stage = (args.products == 3 ? 4u : 2u) * kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + (size_t)args.n_stages * per_stage);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (8u + s); };
    auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (16u + a); };
    auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (20u + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const uint32_t tiles_m = (args.m + 2 * kGemmTileM - 1) / (2 * kGemmTileM);
    const uint32_t tiles_n = (args.n + kGemmPairN - 1) / kGemmPairN;  // the last one may be 128 wide
    const uint32_t n_tiles = tiles_m * tiles_n, n_kb = args.k / kMmaKBlock;
    auto tile_width = [&](uint32_t nt) { return min((uint32_t)kGemmPairN, args.n - nt * kGemmPairN); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_a_hi);
        tma_prefetch_desc(&tm_w_hi);
        for (uint32_t s = 0; s < args.n_stages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t a = 0; a < kGemmAccStages; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 16);  // 8 epilogue warps of both CTAs
        }
        fence_barrier_init();
    } else if (warp == 2) {
        tmem_alloc_pair(smem_u32(tmem_slot), kGemmAccStages * kGemmPairN);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs; completion bytes land on the leader's barriers) =====
        uint32_t stage = 0, phase = 0;
        for (uint32_t t = pair; t < n_tiles; t += n_pairs) {
            const uint32_t nt = t % tiles_n, width = tile_width(nt);
            const int32_t row_a = (int32_t)((t / tiles_n) * 2 * kGemmTileM + rank * kGemmTileM);
            const int32_t row_w = (int32_t)(nt * kGemmPairN + rank * (width / 2));
            const bool wide = width == kGemmPairN;
            const uint32_t w_bytes = wide ? kMmaTileBytes : kMmaTileBytes / 2;
            const uint32_t bytes_per_cta = (args.products == 3 ? 2u : 1u) * (kMmaTileBytes + w_bytes);
            for (uint32_t kb = 0; kb < n_kb; ++kb) {
                mbar_wait(empty_bar(stage), phase ^ 1u);
                if (elect_one()) {
                    const uint32_t s0 = base + stage * per_stage;
                    const int32_t kc = (int32_t)(kb * kMmaKBlock);
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2u * bytes_per_cta);
                    tma_load_2d_pair(s0, &tm_a_hi, full_bar(stage), kc, row_a);
                    tma_load_2d_pair(s0 + kMmaTileBytes, wide ? &tm_w_hi : &tm_w64_hi, full_bar(stage), kc, row_w);
                    if (args.products == 3) {
                        tma_load_2d_pair(s0 + 2 * kMmaTileBytes, &tm_a_lo, full_bar(stage), kc, row_a);
                        tma_load_2d_pair(s0 + 3 * kMmaTileBytes, wide ? &tm_w_lo : &tm_w64_lo, full_bar(stage), kc, row_w);
                    }
                }
                __syncwarp();
                if (++stage == args.n_stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // ===== MMA issuer (leader CTA only) =====
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = pair; t < n_tiles; t += n_pairs) {
                const uint32_t idesc = tile_width(t % tiles_n) == kGemmPairN ? umma_idesc_f16(2 * kGemmTileM, kGemmPairN)
                                                                             : umma_idesc_f16(2 * kGemmTileM, kGemmPairN / 2);
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * kGemmPairN;
                for (uint32_t kb = 0; kb < n_kb; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t s0 = base + stage * per_stage;
                        const uint64_t a_hi = umma_desc_sw128(s0), w_hi = umma_desc_sw128(s0 + kMmaTileBytes);
                        const uint64_t a_lo = umma_desc_sw128(s0 + 2 * kMmaTileBytes);
                        const uint64_t w_lo = umma_desc_sw128(s0 + 3 * kMmaTileBytes);
#pragma unroll
                        for (uint32_t k4 = 0; k4 < kMmaKBlock / 16; ++k4)
                            umma_f16_pair(d_tmem, a_hi + 2u * k4, w_hi + 2u * k4, idesc, (kb | k4) != 0u ? 1u : 0u);
                        if (args.products == 3) {
#pragma unroll
                            for (uint32_t k4 = 0; k4 < kMmaKBlock / 16; ++k4) {
                                umma_f16_pair(d_tmem, a_lo + 2u * k4, w_hi + 2u * k4, idesc, 1u);
                                umma_f16_pair(d_tmem, a_hi + 2u * k4, w_lo + 2u * k4, idesc, 1u);
                            }
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (kb + 1 == n_kb) umma_commit_pair(tfull_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == args.n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++acc == kGemmAccStages) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue (both CTAs): TMEM lane = one of this CTA's 128 rows, column = feature =====
        const uint32_t quarter = warp & 3u;
        const uint32_t half = (warp - 2u) >> 2;
        float* tile = reinterpret_cast<float*>(bars + 32) + (warp - 2u) * kGemmXposeFloats;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = pair; t < n_tiles; t += n_pairs) {
            const uint32_t nt = t % tiles_n, width = tile_width(nt);
            const uint32_t row_base = (t / tiles_n) * 2 * kGemmTileM + rank * kGemmTileM + quarter * 32u;
            const uint32_t cols_per_warp = width / 2;  // 128 or 64
            const uint32_t col0 = nt * kGemmPairN + half * cols_per_warp;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + acc * kGemmPairN + half * cols_per_warp;
            gemm_epilogue_warp(args, taddr, row_base, col0, cols_per_warp, tile, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
            if (++acc == kGemmAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kGemmAccStages * kGemmPairN);
    }
}

// ─── row-wise f32 kernels ───────────────────────────────────────────────────────────────────
// One warp per token row of H = 384 (12 values per lane).  LayerNorm as PyTorch computes it:
// mean, then biased variance of (x - mean), eps inside the sqrt.
constexpr int kHidden = 384;

__device__ __forceinline__ void layernorm_row(float (&x)[12], const float* __restrict__ g, const float* __restrict__ b,
                                              float eps, uint32_t lane) {
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 12; ++i) s += x[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kHidden);
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const float d = x[i] - mean;
        v = fmaf(d, d, v);
    }
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float inv = rsqrtf(v * (1.0f / kHidden) + eps);
    // rsqrtf is within 2 ulp; one Newton step brings it to correctly-rounded quality
    const float var_eps = v * (1.0f / kHidden) + eps;
    const float inv2 = inv * (1.5f - 0.5f * var_eps * inv * inv);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint32_t d = lane + 32u * i;
        x[i] = (x[i] - mean) * inv2 * g[d] + b[d];
    }
}

__device__ __forceinline__ void store_row(const float (&x)[12], size_t row, float* out_f32, __half* out_hi,
                                          __half* out_lo, uint32_t lane) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const size_t o = row * kHidden + lane + 32u * i;
        out_f32[o] = x[i];
        __half h, l;
        split_f16(x[i], h, l);
        out_hi[o] = h;
        out_lo[o] = l;
    }
}

// embeddings: word[id] + position[t] + token_type[0] -> LayerNorm   (BertEmbeddings)
__global__ void __launch_bounds__(256)
minilm_embed_kernel(const int32_t* __restrict__ ids, uint32_t batch, uint32_t t_pad, uint32_t vocab,
                    const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type0,
                    const float* __restrict__ g, const float* __restrict__ b, float eps, float* out_f32,
                    __half* out_hi, __half* out_lo) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= (size_t)batch * t_pad) return;
    const uint32_t t = (uint32_t)(row % t_pad);
    int32_t id = ids[row];
    if (id < 0 || (uint32_t)id >= vocab) id = 0;  // pad slots / out-of-range ids read row 0 (masked later)
    float x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const uint32_t d = lane + 32u * i;
        x[i] = word[(size_t)id * kHidden + d] + pos[(size_t)t * kHidden + d] + type0[d];
    }
    layernorm_row(x, g, b, eps, lane);
    store_row(x, row, out_f32, out_hi, out_lo, lane);
}

// post-LN residual block tail: LayerNorm(pre) where `pre` already holds linear + bias + residual
__global__ void __launch_bounds__(256)
minilm_layernorm_kernel(const float* __restrict__ pre, size_t rows, const float* __restrict__ g,
                        const float* __restrict__ b, float eps, float* out_f32, __half* out_hi, __half* out_lo) {
    const uint32_t lane = threadIdx.x & 31;
    const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = pre[row * kHidden + lane + 32u * i];
    layernorm_row(x, g, b, eps, lane);
    store_row(x, row, out_f32, out_hi, out_lo, lane);
}

// ─── attention: one CTA per (sequence, head); f32 on CUDA cores ─────────────────────────────
// qkv [M, 1152] f32 = [q | k | v], head h at columns h*32.  softmax(q k^T / sqrt(32)) over the
// sequence's valid keys (native.rs:82-133), context written as split f16 for the out-projection.
constexpr int kHeads = 12, kHeadDim = 32;

__global__ void __launch_bounds__(128)
minilm_attention_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ lens, uint32_t t_pad,
                        __half* __restrict__ ctx_hi, __half* __restrict__ ctx_lo) {
    extern __shared__ __align__(16) float att_smem[];
    const uint32_t b = blockIdx.x / kHeads, h = blockIdx.x % kHeads;
    const uint32_t len = min((uint32_t)max(lens[b], 0), t_pad);
    float* ks = att_smem;                       // [len][33]  (padded rows: conflict-free column reads)
    float* vs = ks + (size_t)t_pad * 33;        // [len][33]
    float* ps = vs + (size_t)t_pad * 33;        // [4 warps][t_pad] probabilities of the row in flight
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t row0 = (size_t)b * t_pad;
    for (uint32_t i = threadIdx.x; i < len * kHeadDim; i += blockDim.x) {
        const uint32_t t = i / kHeadDim, d = i % kHeadDim;
        const float* src = qkv + (row0 + t) * (3 * kHidden) + h * kHeadDim + d;
        ks[t * 33 + d] = src[kHidden];
        vs[t * 33 + d] = src[2 * kHidden];
    }
    __syncthreads();
    const float scale = 0.17677669529663688110f;  // 1 / sqrt(32)
    float* p = ps + (size_t)warp * t_pad;
    for (uint32_t t = warp; t < t_pad; t += 4) {
        const size_t o = (row0 + t) * kHidden + h * kHeadDim + lane;
        if (t >= len) {  // pad rows: defined output, never read by the pooling
            ctx_hi[o] = __float2half_rn(0.0f);
            ctx_lo[o] = __float2half_rn(0.0f);
            continue;
        }
        const float qd = qkv[(row0 + t) * (3 * kHidden) + h * kHeadDim + lane];  // lane = dim
        float mx = -INFINITY;
        for (uint32_t j0 = 0; j0 < len; j0 += 32) {
            const uint32_t j = j0 + lane;
            float s = 0.0f;
#pragma unroll
            for (int d = 0; d < kHeadDim; ++d) {
                const float qv = __shfl_sync(0xffffffffu, qd, d);
                s = fmaf(qv, j < len ? ks[j * 33 + d] : 0.0f, s);
            }
            s *= scale;
            if (j < len) {
                p[j] = s;
                mx = fmaxf(mx, s);
            }
        }
        for (int o2 = 16; o2 > 0; o2 >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o2));
        __syncwarp();
        float sum = 0.0f;
        for (uint32_t j = lane; j < len; j += 32) {
            const float e = expf(p[j] - mx);
            p[j] = e;
            sum += e;
        }
        for (int o2 = 16; o2 > 0; o2 >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o2);
        __syncwarp();
        float acc = 0.0f;  // lane = output dim
        for (uint32_t j = 0; j < len; ++j) acc = fmaf(p[j], vs[j * 33 + lane], acc);
        acc /= sum;
        __half hi, lo;
        split_f16(acc, hi, lo);
        ctx_hi[o] = hi;
        ctx_lo[o] = lo;
        __syncwarp();
    }
}

// Short sequences (t_pad <= 32, the query-encoding case): one WARP per (sequence, head), one LANE
// per query row.  K and V rows sit in a per-warp shared tile and are read as broadcast float4s;
// each lane keeps its q row, its 32 scores and its 32 outputs in registers, so the soft-max needs
// no shuffles and the two contractions are plain FMA streams (~2.9 k instructions per head against
// ~4.8 k for the shuffle-broadcast form: 150 -> ~90 us per layer at 1024 x 32 tokens).
__global__ void __launch_bounds__(128)
minilm_attention_short_kernel(const float* __restrict__ qkv, const int32_t* __restrict__ lens, uint32_t batch,
                              uint32_t t_pad, __half* __restrict__ ctx_hi, __half* __restrict__ ctx_lo) {
    __shared__ __align__(16) float kv_all[4][2][32 * kHeadDim];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t unit = blockIdx.x * 4 + warp;
    if (unit >= batch * kHeads) return;
    const uint32_t b = unit / kHeads, h = unit % kHeads;
    const uint32_t len = min((uint32_t)max(lens[b], 0), t_pad);
    float* ks = kv_all[warp][0];
    float* vs = kv_all[warp][1];
    const size_t row0 = (size_t)b * t_pad;
    // stage K and V rows [len][32]: 8 lanes per row, 4 rows per instruction
    for (uint32_t r = lane >> 3; r < len; r += 4) {
        const float* src = qkv + (row0 + r) * (3 * kHidden) + h * kHeadDim + 4 * (lane & 7);
        *reinterpret_cast<float4*>(ks + r * kHeadDim + 4 * (lane & 7)) = *reinterpret_cast<const float4*>(src + kHidden);
        *reinterpret_cast<float4*>(vs + r * kHeadDim + 4 * (lane & 7)) = *reinterpret_cast<const float4*>(src + 2 * kHidden);
    }
    const uint32_t t = lane;  // this lane's query row
    const size_t o = (row0 + t) * kHidden + h * kHeadDim;
    float q[kHeadDim];
    if (t < len) {
        const float4* qs = reinterpret_cast<const float4*>(qkv + (row0 + t) * (3 * kHidden) + h * kHeadDim);
#pragma unroll
        for (int d = 0; d < kHeadDim / 4; ++d) {
            const float4 v4 = qs[d];
            q[4 * d] = v4.x; q[4 * d + 1] = v4.y; q[4 * d + 2] = v4.z; q[4 * d + 3] = v4.w;
        }
    } else {
#pragma unroll
        for (int d = 0; d < kHeadDim; ++d) q[d] = 0.0f;
    }
    __syncwarp();
    const float scale = 0.17677669529663688110f;  // 1 / sqrt(32)
    float sc[32];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        float a = -INFINITY;
        if ((uint32_t)j < len) {  // warp-uniform
            const float4* kr = reinterpret_cast<const float4*>(ks + j * kHeadDim);
            a = 0.0f;
#pragma unroll
            for (int d = 0; d < kHeadDim / 4; ++d) {
                const float4 k4 = kr[d];
                a = fmaf(q[4 * d], k4.x, a);
                a = fmaf(q[4 * d + 1], k4.y, a);
                a = fmaf(q[4 * d + 2], k4.z, a);
                a = fmaf(q[4 * d + 3], k4.w, a);
            }
            a *= scale;
            mx = fmaxf(mx, a);
        }
        sc[j] = a;
    }
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        sc[j] = (uint32_t)j < len ? expf(sc[j] - mx) : 0.0f;
        sum += sc[j];
    }
    float out[kHeadDim];
#pragma unroll
    for (int d = 0; d < kHeadDim; ++d) out[d] = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        if ((uint32_t)j < len) {
            const float4* vr = reinterpret_cast<const float4*>(vs + j * kHeadDim);
            const float pj = sc[j];
#pragma unroll
            for (int d = 0; d < kHeadDim / 4; ++d) {
                const float4 v4 = vr[d];
                out[4 * d] = fmaf(pj, v4.x, out[4 * d]);
                out[4 * d + 1] = fmaf(pj, v4.y, out[4 * d + 1]);
                out[4 * d + 2] = fmaf(pj, v4.z, out[4 * d + 2]);
                out[4 * d + 3] = fmaf(pj, v4.w, out[4 * d + 3]);
            }
        }
    }
    if (t >= t_pad) return;
    const float inv = t < len ? 1.0f / sum : 0.0f;  // pad rows: defined zeros, never read by the pooling
#pragma unroll
    for (int d = 0; d < kHeadDim; d += 4) {
        __half hh[4], ll[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_f16(t < len ? out[d + i] * inv : 0.0f, hh[i], ll[i]);
        *reinterpret_cast<uint2*>(ctx_hi + o + d) = *reinterpret_cast<uint2*>(hh);
        *reinterpret_cast<uint2*>(ctx_lo + o + d) = *reinterpret_cast<uint2*>(ll);
    }
}

// ─── pooling: masked mean -> L2 (eps 1e-12) -> adapter L2 with zero-vector guard ─────────────
__global__ void __launch_bounds__(128)
minilm_pool_kernel(const float* __restrict__ hidden, const int32_t* __restrict__ lens, uint32_t t_pad,
                   float* __restrict__ out, const uint32_t* __restrict__ offs = nullptr) {
    const uint32_t b = blockIdx.x;
    const uint32_t len = min((uint32_t)max(lens[b], 0), t_pad);
    const size_t row0 = offs ? (size_t)offs[b] : (size_t)b * t_pad;  // packed rows (f16 form) or batch-longest padding
    __shared__ float red[4];
    float v[3];
    float part = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const uint32_t d = threadIdx.x + 128u * i;
        float s = 0.0f;
        for (uint32_t t = 0; t < len; ++t) s += hidden[(row0 + t) * kHidden + d];
        v[i] = len ? s / (float)len : 0.0f;  // sum(mask * h) / clamp(sum(mask), 1e-9)
        part = fmaf(v[i], v[i], part);
    }
    auto block_sum = [&](float x) {
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
        __syncthreads();
        return (red[0] + red[1]) + (red[2] + red[3]);
    };
    const float n1 = fmaxf(sqrtf(block_sum(part)), 1e-12f);  // fastembed normalize
    float part2 = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        v[i] = v[i] / n1;
        part2 = fmaf(v[i], v[i], part2);
    }
    const float norm_sq = block_sum(part2);  // adapter normalize_in_place (fastembed_embedder.rs:416-426)
    const bool ok = isfinite(norm_sq) && norm_sq > 1.1920929e-07f;
    const float inv = ok ? 1.0f / sqrtf(norm_sq) : 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) out[(size_t)b * kHidden + threadIdx.x + 128u * i] = ok ? v[i] * inv : 0.0f;
}

// f32 -> split f16 (weights at load time)
__global__ void split_f16_kernel(const float* __restrict__ src, size_t n, __half* __restrict__ hi,
                                 __half* __restrict__ lo) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        __half h, l;
        split_f16(src[i], h, l);
        hi[i] = h;
        lo[i] = l;
    }
}

}  // namespace fsgpu
