// fsgpu_peak.cu — measurement utility: the tensor-pipe issue ceiling of THIS GPU for the MMA shapes the
// batched scan uses (tcgen05.mma.cta_group::2, M = 256 x N = 256, kind::f16 or kind::i8), so that
// bench.py can quote the scan kernel against a measured int8 figure (MEASURED_PEAKS.json only carries a
// cuBLAS bf16 number).  The kernel is the scan's MMA loop with everything else removed: operands sit in
// shared memory (128-byte swizzled K-blocks, pseudo-random bytes), one elected thread of every leader CTA
// issues back-to-back MMAs into two alternating TMEM accumulators, nothing is loaded or read back.
#include <algorithm>
#include <cstdint>

#include "fsgpu.h"
#include "fsgpu_host.cuh"
#include "tc_ptx.cuh"

using namespace fsgpu;

template <bool I8>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) tensor_peak_kernel(uint32_t iters) {
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
    // operands: small pseudo-random values (f16: bit patterns of finite numbers below 2; int8: any byte)
    uint32_t s = 0x9E3779B9u * (blockIdx.x * blockDim.x + threadIdx.x + 1u);
    for (uint32_t i = threadIdx.x; i < 2 * kMmaTileBytes / 4; i += blockDim.x) {
        s = s * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t*>(base_ptr)[i] = I8 ? s : (s & 0x3BFF3BFFu);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async (tensor) proxy reads
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
    if (warp == 0 && rank == 0) {
        constexpr uint32_t idesc = I8 ? umma_idesc_i8(256, 256) : umma_idesc_f16(256, 256);
        const uint64_t a_desc = umma_desc_sw128(a_smem), b_desc = umma_desc_sw128(b_smem);
        for (uint32_t it = 0; it < iters; ++it) {
            if (elect_one()) {
                const uint32_t d = tmem_base + (it & 1u) * 256u;
#pragma unroll
                for (uint32_t k4 = 0; k4 < 4; ++k4) {
                    if constexpr (I8)
                        umma_i8_pair(d, a_desc + 2u * k4, b_desc + 2u * k4, idesc, 1u);
                    else
                        umma_f16_pair(d, a_desc + 2u * k4, b_desc + 2u * k4, idesc, 1u);
                }
            }
            __syncwarp();
        }
        if (elect_one()) umma_commit_pair(done_bar);
        __syncwarp();
    }
    mbar_wait(done_bar, 0);
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

extern "C" int fsgpu_measure_tensor_peak(int device, int kind, uint32_t target_ms, double* out_ops_per_s) {
    if (!out_ops_per_s) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out_ops_per_s = 0.0;
    if (kind != 0 && kind != 1) return fail(FSGPU_ERR_INVALID_CONFIG, "kind must be 0 (f16) or 1 (int8)");
    int ndev = 0;
    int rc = fsgpu_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(FSGPU_ERR_INVALID_CONFIG, "device %d not present", device);
    DeviceGuard g(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: device %d is not sm_100", device);
    const int grid = (prop.multiProcessorCount / 2) * 2;
    const size_t smem = 1024 + 2 * kMmaTileBytes + 64;
    auto kernel = kind == 1 ? tensor_peak_kernel<true> : tensor_peak_kernel<false>;
    CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaStream_t s;
    CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    // ops per MMA: 2 * M * N * K with K = 32 bytes of operand per instruction (16 f16 / 32 int8)
    const double ops_per_iter = 4.0 * 2.0 * 256.0 * 256.0 * (kind == 1 ? 32.0 : 16.0) * (grid / 2);
    uint32_t iters = 2000;
    double best = 0.0;
    cudaError_t err = cudaSuccess;
    for (int rep = 0; rep < 5 && err == cudaSuccess; ++rep) {
        cudaEventRecord(e0, s);
        kernel<<<grid, 128, smem, s>>>(iters);
        cudaEventRecord(e1, s);
        err = cudaStreamSynchronize(s);
        if (err != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep == 0) {  // size the timed launches for ~target_ms
            iters = (uint32_t)std::min(4.0e6, std::max(2000.0, iters * (double)std::max(1u, target_ms) / std::max(ms, 0.01f) / 3.0));
            continue;
        }
        best = std::max(best, ops_per_iter * iters / (ms * 1e-3));
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(s);
    if (err != cudaSuccess) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: tensor peak kernel failed: %s", cudaGetErrorString(err));
    *out_ops_per_s = best;
    return FSGPU_OK;
}
