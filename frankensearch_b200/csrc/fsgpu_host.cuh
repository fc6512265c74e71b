// fsgpu_host.cuh — host-side helpers shared by the translation units of libfsgpu.so.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "fsgpu.h"

// Records the thread-local message returned by fsgpu_last_error() and returns `code`.
int fsgpu_fail(int code, const char* fmt, ...);
#define fail fsgpu_fail

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(FSGPU_ERR_SUBSYSTEM, "gpu: %s failed: %s (%s:%d)", #expr,              \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                           \
    } while (0)

// Host -> device copy that has LANDED when it returns.  A plain cudaMemcpy from pageable memory may
// return once the bytes are staged, with the DMA still in flight on the legacy stream, and this
// library's streams are non-blocking (they do not order against the legacy stream): a kernel
// launched right after on one of them could read the destination too early.
static inline cudaError_t h2d_complete(void* dst, const void* src, size_t bytes) {
    cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamLegacy);
    return e;
}

// Small RAII-free device buffer that only grows.
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    cudaError_t reserve(size_t want) {
        if (want <= bytes) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) bytes = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// One-shot device staging for the host-buffer fusion / encoder entry points.
struct Staging {
    std::vector<void*> ptrs;
    ~Staging() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class T>
    cudaError_t up(const T* host, size_t count, T** dev) {
        *dev = nullptr;
        if (!host || count == 0) return cudaSuccess;
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, count * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        *dev = reinterpret_cast<T*>(p);
        return h2d_complete(p, host, count * sizeof(T));
    }
    template <class T>
    cudaError_t alloc(size_t count, T** dev) {
        void* p = nullptr;
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(1, count) * sizeof(T));
        if (e != cudaSuccess) return e;
        ptrs.push_back(p);
        *dev = reinterpret_cast<T*>(p);
        return cudaSuccess;
    }
};

inline int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

inline uint32_t host_next_pow2(uint32_t x) {
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// TMA descriptor of a row-major [rows, dim] f16 matrix read as [box_rows x 64 elements] boxes with
// the 128-byte swizzle the UMMA shared-memory descriptors expect (defined in fsgpu_api.cu).
bool fsgpu_tma_available();
bool make_f16_tile_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t dim, uint32_t box_rows = 128);
// The same for any 2- or 4-byte element type and box: row-major [rows, cols], box [box_cols x box_rows] with
// box_cols * elem_bytes == 128 (the 128-byte swizzle).  Used for TMA stores of GEMM outputs.
bool make_tile_map_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t cols, uint32_t elem_bytes,
                      uint32_t box_cols, uint32_t box_rows);
