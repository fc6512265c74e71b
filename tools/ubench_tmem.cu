// ubench_tmem.cu — microbenchmark behind DESIGN.md's epilogue budget of the int8 scan:
//   (1) TMEM read throughput (tcgen05.ld) per SM for 1 / 4 / 8 reading warps and several shapes,
//   (2) the same while one thread issues back-to-back tcgen05.mma.cta_group::2 kind::i8 (M=256, N=256 or
//       N=128) — does the MMA rate hold, and how many accumulator bytes can the epilogue read per MMA cycle.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I frankensearch_b200/csrc \
//        -o tools/_bin/ubench_tmem tools/ubench_tmem.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"

using namespace fsgpu;

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e_ = (x);                                                              \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));     \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

__device__ __forceinline__ void tmem_ld_x64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 256 bits per repetition: x8 = 32 registers per thread (4 KiB per warp, like 32x32b.x32)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]),
          "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// non-blocking probe (try_wait may suspend the thread up to a time limit, which would throttle the readers)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ uint32_t fold32(const uint32_t (&v)[32]) {
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= v[i];
    return x;
}

// One pass over 128 accumulator columns of this warp's lane quarter, in the given shape.
// shape 0: 4 x (32x32b.x32) then one wait   shape 1: 2 x (32x32b.x64) then one wait
// shape 2: x32, wait, x32, wait ...          shape 3: 4 x (16x256b.x8) (16 lanes each) then one wait
template <int SHAPE>
__device__ __forceinline__ uint32_t read128(uint32_t taddr) {
    uint32_t x = 0;
    if constexpr (SHAPE == 0) {
        uint32_t v[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_x32(taddr + c * 32u, v[c]);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) x ^= fold32(v[c]);
    } else if constexpr (SHAPE == 1) {
        uint32_t v[2][64];
        tmem_ld_x64(taddr, v[0]);
        tmem_ld_x64(taddr + 64u, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 64; ++i) x ^= v[0][i] ^ v[1][i];
    } else if constexpr (SHAPE == 2) {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            tmem_ld_x32(taddr + c * 32u, v);
            tmem_ld_wait();
            x ^= fold32(v);
        }
    } else {
        uint32_t v[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_16x256b_x8(taddr + ((uint32_t)(c >> 1) << 20) + (c & 1) * 64u, v[c]);  // 16 lanes x 64 columns each
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) x ^= fold32(v[c]);
    }
    return x;
}

// (1) readers only.  n_warps in {1, 4, 8}: warp w reads lane quarter w % 4, column half (w / 4).
template <int SHAPE>
__global__ void __launch_bounds__(256, 1) ldtm_kernel(uint32_t iters, uint32_t n_warps, long long* out_clk, uint32_t* sink) {
    __shared__ uint32_t slot;
    const uint32_t warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = slot;
    uint32_t x = 0;
    const long long t0 = clock64();
    if (warp < n_warps) {
        const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 128u;
        for (uint32_t it = 0; it < iters; ++it) x ^= read128<SHAPE>(taddr + (it & 1u) * 256u);
    }
    const long long t1 = clock64();
    if (x == 0x12345u) sink[0] = x;
    if (warp < n_warps && (threadIdx.x & 31) == 0) out_clk[blockIdx.x * 8 + warp] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// (2) MMA issue loop (leader CTA, warp 0) + n_epi reader warps in BOTH CTAs (warps 2..9).
// N = 256: two alternating 256-column accumulators; N = 128: four 128-column ones.
template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
mma_ldtm_kernel(uint32_t iters, uint32_t n_epi, long long* out_clk, uint32_t* out_loops, uint32_t* sink) {
    extern __shared__ uint8_t smem_dyn[];
    const uint32_t raw = smem_u32(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_dyn + (base - raw);
    const uint32_t a_smem = base, b_smem = base + kMmaTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(base_ptr + 2 * kMmaTileBytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const uint32_t done_bar = smem_u32(bars);
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1u);
    for (uint32_t i = threadIdx.x; i < 2 * kMmaTileBytes / 4; i += blockDim.x) {
        s = s * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t*>(base_ptr)[i] = s;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(done_bar, 1);
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc_pair(smem_u32(tmem_slot), 512);
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp == 0) {
        if (rank == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(256, N);
            const uint64_t a_desc = umma_desc_sw128(a_smem), b_desc = umma_desc_sw128(b_smem);
            const long long t0 = clock64();
            for (uint32_t it = 0; it < iters; ++it) {
                if (elect_one()) {
                    const uint32_t d = tmem_base + (it % (512u / N)) * N;
#pragma unroll
                    for (uint32_t k4 = 0; k4 < 4; ++k4) umma_i8_pair(d, a_desc + 2u * k4, b_desc + 2u * k4, idesc, 1u);
                }
                __syncwarp();
            }
            if (elect_one()) umma_commit_pair(done_bar);
            __syncwarp();
            mbar_wait(done_bar, 0);
            const long long t1 = clock64();
            if ((threadIdx.x & 31) == 0) out_clk[blockIdx.x >> 1] = t1 - t0;
        }
    } else if (warp >= 2 && warp - 2 < n_epi) {
        const uint32_t w = warp - 2;
        const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16) + (w >> 2) * 128u;
        uint32_t x = 0, loops = 0;
        while (!mbar_test_wait(done_bar, 0)) {
            x ^= read128<0>(taddr + (loops & 1u) * 256u);
            ++loops;
        }
        if (x == 0x12345u) sink[0] = x;
        if ((threadIdx.x & 31) == 0) out_loops[blockIdx.x * 8 + w] = loops;
    }
    if (warp != 0 || rank != 0) mbar_wait(done_bar, 0);
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int grid = (sms / 2) * 2;
    long long* d_clk;
    uint32_t *d_loops, *d_sink;
    CK(cudaMalloc(&d_clk, sizeof(long long) * grid * 8));
    CK(cudaMalloc(&d_loops, sizeof(uint32_t) * grid * 8));
    CK(cudaMalloc(&d_sink, 16));
    std::vector<long long> clk(grid * 8);
    std::vector<uint32_t> loops(grid * 8);
    const uint32_t iters = 20000;
    printf("# device %s, %d SMs\n", prop.name, sms);
    printf("# (1) tcgen05.ld only: bytes per SM clock per SM (each pass = 128 columns x 32 lanes x 4 B = 16 KiB per warp)\n");
    const char* names[4] = {"4 x 32x32b.x32 + wait", "2 x 32x32b.x64 + wait", "(32x32b.x32 + wait) x 4", "4 x 16x256b.x8 + wait"};
    for (int shape = 0; shape < 4; ++shape) {
        for (uint32_t nw : {1u, 4u, 8u}) {
            CK(cudaMemset(d_clk, 0, sizeof(long long) * grid * 8));
            for (int rep = 0; rep < 2; ++rep) {
                switch (shape) {
                    case 0: ldtm_kernel<0><<<grid, 256>>>(iters, nw, d_clk, d_sink); break;
                    case 1: ldtm_kernel<1><<<grid, 256>>>(iters, nw, d_clk, d_sink); break;
                    case 2: ldtm_kernel<2><<<grid, 256>>>(iters, nw, d_clk, d_sink); break;
                    default: ldtm_kernel<3><<<grid, 256>>>(iters, nw, d_clk, d_sink); break;
                }
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(clk.data(), d_clk, sizeof(long long) * grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int c = 0; c < grid; ++c)
                for (uint32_t w = 0; w < nw; ++w) mx = std::max(mx, clk[c * 8 + w]);
            const double bytes = (double)iters * 16384.0 * nw;
            printf("shape %-26s warps %u: %8.1f clk per pass, %7.1f B/clk/SM\n", names[shape], nw, (double)mx / iters, bytes / mx);
        }
    }
    printf("# (2) kind::i8 cta_group::2 M=256 issue loop (4 MMAs of K=32 per iteration) with n reader warps per CTA\n");
    const size_t smem = 1024 + 2 * kMmaTileBytes + 64;
    CK(cudaFuncSetAttribute(mma_ldtm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(mma_ldtm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int n : {256, 128}) {
        for (uint32_t ne : {0u, 4u, 8u}) {
            CK(cudaMemset(d_loops, 0, sizeof(uint32_t) * grid * 8));
            for (int rep = 0; rep < 2; ++rep) {
                if (n == 256)
                    mma_ldtm_kernel<256><<<grid, 320, smem>>>(iters, ne, d_clk, d_loops, d_sink);
                else
                    mma_ldtm_kernel<128><<<grid, 320, smem>>>(iters, ne, d_clk, d_loops, d_sink);
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(clk.data(), d_clk, sizeof(long long) * grid * 8, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(loops.data(), d_loops, sizeof(uint32_t) * grid * 8, cudaMemcpyDeviceToHost));
            long long mx = 0;
            for (int p = 0; p < grid / 2; ++p) mx = std::max(mx, clk[p]);
            double lp = 0;
            for (int c = 0; c < grid; ++c)
                for (uint32_t w = 0; w < ne; ++w) lp += loops[c * 8 + w];
            const double macs_per_clk_sm = (double)iters * 4.0 * 256.0 * n * 32.0 / 2.0 / mx;  // per SM of the pair
            const double ideal = 4.0 * 128.0 * n * 32.0 / 8192.0;                             // clocks per iteration at 8192 MAC/clk/SM
            printf("N=%3d readers %u: %7.1f clk per iteration (ideal %5.1f), %7.0f MAC/clk/SM, epilogue read %6.1f B/clk/SM\n", n,
                   ne, (double)mx / iters, ideal, macs_per_clk_sm, lp * 16384.0 / grid / mx);
        }
    }
    return 0;
}
