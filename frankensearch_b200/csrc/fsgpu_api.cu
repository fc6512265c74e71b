// fsgpu_api.cu — C ABI (include/fsgpu.h) over the sm_100a kernels: index lifetime, launch
// planning, host<->device staging, FSVI v1 reader.  No CPU compute path exists in this file:
// every search/fusion/embedding entry point launches kernels or fails.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "fsgpu.h"
#include "fsgpu_common.cuh"
#include "fsgpu_host.cuh"
#include "mma_scan_kernels.cuh"
#include "scan_kernels.cuh"
#include "select_kernels.cuh"
#include "two_pass_kernels.cuh"
#include "synth_kernels.cuh"

using namespace fsgpu;

// ─── errors ─────────────────────────────────────────────────────────────────────────────────
static thread_local std::string g_last_error;

int fsgpu_fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

extern "C" const char* fsgpu_last_error(void) { return g_last_error.c_str(); }
extern "C" int fsgpu_abi_version(void) { return FSGPU_ABI_VERSION; }

extern "C" int fsgpu_device_count(int* out_count) {
    if (!out_count) return fail(FSGPU_ERR_INVALID_CONFIG, "out_count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *out_count = 0;
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: no usable CUDA device: %s", cudaGetErrorString(e));
    }
    *out_count = n;
    return FSGPU_OK;
}

extern "C" void fsgpu_index_options_default(fsgpu_index_options* o) {
    if (!o) return;
    o->device = 0;
    o->reduce_order = FSGPU_REDUCE_HALVES_PAIRWISE;
    o->tail_fma = 1;
    o->slab_is_device = 0;
    o->row_base = 0;
    o->int8_codes = 1;
    o->flags = 0;
}

// ─── the index handle ───────────────────────────────────────────────────────────────────────
struct fsgpu_index {
    int device = 0;
    int num_sms = 0;
    uint64_t n_rows = 0, row_base = 0;
    uint32_t dim = 0;
    int reduce_order = 0, tail_fma = 1;
    uint16_t* d_slab = nullptr;
    bool owns_slab = false;
    // an f32-quantised FSVI file (quantization 0, lib.rs:6-43) keeps its slab as f32 and is scored with the
    // reference's f32 kernel (dot_product_f32_bytes_f32, search.rs:1300-1321): exact score of every row +
    // radix select; the f16 / int8 scan forms do not apply
    float* d_slab_f32 = nullptr;
    const void* slab_any() const { return d_slab_f32 ? static_cast<const void*>(d_slab_f32) : static_cast<const void*>(d_slab); }
    int is_f32() const { return d_slab_f32 ? 1 : 0; }
    uint8_t* d_tomb = nullptr;
    mutable const uint8_t* d_excl = nullptr;  // per-call exclusion bitmap (tombstones | !filter), else nullptr
    mutable DevBuf ws_excl, ws_allow;
    // resident WAL rows (f32, appended since the last compaction): scored beside the slab and merged
    // into the same top-k (search.rs:488-491, :1449-1475)
    DevBuf d_wal;
    uint32_t n_wal = 0;
    uint64_t wal_base = 0;                       // hit row of WAL entry 0 (record_count, search.rs:1583)
    mutable const uint8_t* d_wal_allow = nullptr;  // per-call allow bitmap (bit n_rows + w), else nullptr
    mutable bool wal_allow_bit0_is_zero = false;   // ... or a bitmap over the WAL rows alone (bit w)
    mutable DevBuf ws_wal_main, ws_wal_keys;
    // doc-id hashes of the rows (record table field 0, lib.rs:130-174) for hash filters on the device
    DevBuf d_hashes;
    bool has_hashes = false;
    // per-call selective gather (try_gather_filtered, search.rs:1114-1161): listed rows replace the scan
    mutable const uint32_t* d_gather_pos = nullptr;
    mutable const uint32_t* d_gather_count = nullptr;
    mutable uint32_t gather_cap = 0;
    mutable DevBuf ws_allowed, ws_gather_pos, ws_gather_count, ws_gather_keys;
    cudaStream_t stream = nullptr;
    mutable std::mutex mu;
    // workspaces (grow-only, guarded by mu)
    mutable DevBuf ws_partial, ws_queries, ws_keys, ws_hits, ws_counts, ws_sort_a, ws_sort_b,
        ws_cub, ws_rows, ws_scores, ws_present;
    // d_error[4]: {0: contract-violation flag, 1: "some query needs the exact path" (set by the refine
    // kernel, cleared per sub-batch), 2: bails (a redo launch found more flagged queries than its limit),
    // 3: flagged queries served by device-side redo launches since the call began}
    uint32_t* d_error = nullptr;
    uint32_t* h_flags = nullptr;    // pinned mirror of d_error[0..4): copied at the end of every batched search
    // device-side redo (scan_kernels.cuh RedoArgs): slot -> query table and per-launch count
    mutable DevBuf ws_redo_slots, ws_redo_partial;
    // calls may arrive on different streams; the workspaces above are shared, so every call first waits
    // for the previous one (cudaStreamWaitEvent: device-side ordering, the host never blocks)
    cudaEvent_t ev_last = nullptr;
    mutable cudaStream_t last_stream = nullptr;
    mutable bool have_last = false;
    // adaptive form choice for stream-asynchronous callers: when the int8 form's candidate lists
    // overflowed for more than a few queries of a call, the next calls use the f16 form
    mutable cudaEvent_t ev_flags = nullptr;
    mutable bool flags_pending = false;
    mutable int i8_penalty = 0;
    mutable bool force_f16 = false;  // set by a synchronous caller's retry
    mutable bool sync_caller = true;  // the entry point synchronises before returning (it may retry)
    // batched tensor-core path (mma_scan_kernels.cuh): slab statistics for the error bound, the
    // slab's TMA descriptor, workspaces
    bool mma_ok = false;            // dim % 64 == 0, dim <= 512, every element finite, TMA usable
    float max_row_norm = 0.0f;      // upper bound on ||row||_2 over the slab
    CUtensorMap tm_slab;
    mutable CUtensorMap tm_qhat;
    mutable const void* tm_qhat_ptr = nullptr;
    mutable uint32_t tm_qhat_rows = 0;
    mutable DevBuf ws_qhat, ws_margin, ws_gate, ws_redo, ws_cand, ws_cand_count, ws_progress;
    // int8 form of the batched path (FSGPU_MMA_I8=1 at index creation): corpus codes with the
    // reference's corpus-wide scale (simd.rs:1842-1859), their TMA descriptor, the measured bound on
    // the per-row quantisation error
    DevBuf d_slab_i8;
    bool i8_ok = false;
    bool want_i8 = true;  // fsgpu_index_options.int8_codes
    float i8_sx = 0.0f, i8_max_ex = 0.0f;
    CUtensorMap tm_slab_i8;
    mutable CUtensorMap tm_qhat_i8;
    mutable const void* tm_qhat_i8_ptr = nullptr;
    mutable uint32_t tm_qhat_i8_rows = 0;
    mutable DevBuf ws_qscale;
    // single-query int8 pass 1 (host API only: it needs the end-of-call synchronisation)
    mutable bool use_i8_single = false;
    mutable DevBuf ws_approx, ws_i8_top, ws_i8_cnt;
    // large-k radix select (select_kernels.cuh): position lists and the two select states
    mutable DevBuf ws_sel_pos, ws_sel_pos2, ws_sel_state;
    // the reference's quantised two-pass searches (two_pass_kernels.cuh): lazily built code slabs
    mutable DevBuf d_tp_codes8, d_tp_codes4;
    mutable bool tp8_ready = false, tp4_ready = false;
    mutable float tp_max_abs = -1.0f;  // corpus-wide max |element| (ignoring NaN), -1 = not computed yet
    // launch accounting (guarded by mu)
    mutable bool profiling = false;
    mutable fsgpu_profile prof{};
    mutable std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
    // host-side doc-id table
    std::vector<uint8_t> doc_bytes;
    std::vector<uint64_t> doc_off;
};

// Dynamic shared memory of the kernels that stage one f32 query: rounded up with a spare vector, because the
// compiler may fetch the scalar tail of warp_exact_dot (dim % 8 != 0) as one 12- or 16-byte load that reaches a few
// bytes past element dim - 1 (compute-sanitizer: benign but out of bounds at dim = 70).
static inline size_t query_smem_bytes(uint32_t dim) { return (((size_t)dim * 4 + 15) & ~(size_t)15) + 16; }

// ─── launch planning ────────────────────────────────────────────────────────────────────────
constexpr uint32_t kFusedMaxK = 1024;

static uint32_t cand_capacity(uint32_t k) { return host_next_pow2(std::max(2 * k, k + 512)); }

typedef void (*ScanKernel)(const ScanArgs);

template <int NJ, int QB, int R>
static ScanKernel fast_kernel() { return scan_topk_fast_kernel<NJ, QB, R>; }

template <int NJ>
static ScanKernel pick_fast_qb(int qb, int r) {
    if (r == 2) {
        switch (qb) {
            case 1: return fast_kernel<NJ, 1, 2>();
            case 2: return fast_kernel<NJ, 2, 2>();
            case 4: return fast_kernel<NJ, 4, 2>();
        }
    } else {
        switch (qb) {
            case 1: return fast_kernel<NJ, 1, 1>();
            case 2: return fast_kernel<NJ, 2, 1>();
            case 4: return fast_kernel<NJ, 4, 1>();
            case 8: return fast_kernel<NJ, 8, 1>();
        }
    }
    return nullptr;
}
static ScanKernel pick_fast(uint32_t dim, int qb, int r) {
    switch (dim) {
        case 128: return pick_fast_qb<4>(qb, r);
        case 256: return pick_fast_qb<8>(qb, r);
        case 384: return pick_fast_qb<12>(qb, r);
    }
    return nullptr;
}

struct ScanPlan {
    ScanKernel kernel = nullptr;
    int qb = 1, r = 1;
    uint32_t tile_rows = 8;
    uint32_t cap = 1024, sync_every = 1;
    size_t smem = 0;
    int grid = 1;
};

static int make_plan(const fsgpu_index* ix, uint32_t k, int qb_want, ScanPlan* plan) {
    ScanPlan p;
    p.cap = cand_capacity(k);
    int r = env_int("FSGPU_SCAN_R", 2);
    if (r != 1 && r != 2) r = 1;
    if (qb_want == 8) r = 1;
    p.kernel = pick_fast(ix->dim, qb_want, r);
    if (p.kernel) {
        p.qb = qb_want;
        p.r = r;
        p.tile_rows = kScanWarps * 8 * r;
    } else {
        p.kernel = scan_topk_generic_kernel;
        p.qb = 1;
        p.r = 1;
        p.tile_rows = kScanWarps;
    }
    p.sync_every = std::max<uint32_t>(1, std::min<uint32_t>(16, (p.cap - k) / (2 * p.tile_rows)));
    p.smem = scan_smem_bytes(p.qb, p.cap, ix->dim);
    CUDA_TRY(cudaFuncSetAttribute(p.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, p.kernel, kScanThreads, p.smem));
    if (per_sm < 1)
        return fail(FSGPU_ERR_INVALID_CONFIG, "scan kernel does not fit (k=%u dim=%u smem=%zu)", k,
                    ix->dim, p.smem);
    const int cap_per_sm = env_int("FSGPU_SCAN_CTAS_PER_SM", 0);
    if (cap_per_sm > 0) per_sm = std::min(per_sm, cap_per_sm);
    const uint64_t n_tiles = (ix->n_rows + p.tile_rows - 1) / p.tile_rows;
    p.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)per_sm * ix->num_sms, n_tiles));
    *plan = p;
    return FSGPU_OK;
}

static int launch_merge(const MergeArgs& m, uint32_t batch, cudaStream_t stream) {
    const size_t smem = (size_t)m.cap * 8 + 16;
    CUDA_TRY(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem));
    merge_topk_kernel<<<batch, kScanThreads, smem, stream>>>(m);
    CUDA_TRY(cudaGetLastError());
    return FSGPU_OK;
}

// Exact top-k for `batch` device-resident queries on the CUDA-core kernels; outputs may be NULL.
// Caller holds ix->mu.
static int search_exact_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch,
                               uint32_t k, uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                               uint32_t* d_out_counts, cudaStream_t stream) {
    if (batch == 0) return FSGPU_OK;
    if (k == 0 || ix->n_rows == 0) {  // search.rs:438-440
        if (d_out_counts) CUDA_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)batch * 4, stream));
        return FSGPU_OK;
    }
    if (k > kFusedMaxK || ix->d_slab_f32) {
        // `limit >= n` / very large k arm (search.rs:449-473): score every row, radix sort.
        const uint64_t n = ix->n_rows;
        CUDA_TRY(ix->ws_sort_a.reserve(n * 8));
        CUDA_TRY(ix->ws_sort_b.reserve(n * 8));
        size_t cub_bytes = 0;
        CUDA_TRY(cub::DeviceRadixSort::SortKeysDescending(nullptr, cub_bytes, ix->ws_sort_a.as<uint64_t>(),
                                                         ix->ws_sort_b.as<uint64_t>(), n, 0, 64, stream));
        CUDA_TRY(ix->ws_cub.reserve(cub_bytes));
        const uint32_t k_eff = (uint32_t)std::min<uint64_t>(k, n);
        const int grid = (int)std::min<uint64_t>((n + kScanWarps - 1) / kScanWarps, (uint64_t)ix->num_sms * 8);
        for (uint32_t b = 0; b < batch; ++b) {
            score_all_kernel<<<grid, kScanThreads, query_smem_bytes(ix->dim), stream>>>(
                ix->slab_any(), ix->is_f32(), ix->d_excl ? ix->d_excl : ix->d_tomb, n, ix->row_base, ix->dim, d_queries + (size_t)b * ix->dim,
                ix->reduce_order, ix->tail_fma, ix->ws_sort_a.as<uint64_t>());
            CUDA_TRY(cudaGetLastError());
            CUDA_TRY(cub::DeviceRadixSort::SortKeysDescending(
                ix->ws_cub.p, cub_bytes, ix->ws_sort_a.as<uint64_t>(), ix->ws_sort_b.as<uint64_t>(), n, 0,
                64, stream));
            emit_sorted_prefix_kernel<<<std::max(1u, std::min(1024u, (k + kScanWarps - 1) / kScanWarps)),
                                        kScanThreads, 0, stream>>>(
                ix->ws_sort_b.as<uint64_t>(), k_eff, k, ix->slab_any(), ix->is_f32(), d_queries + (size_t)b * ix->dim,
                ix->n_rows, ix->row_base, ix->dim, ix->reduce_order, ix->tail_fma,
                d_out_keys ? d_out_keys + (size_t)b * k : nullptr,
                d_out_hits ? d_out_hits + (size_t)b * k : nullptr, d_out_counts ? d_out_counts + b : nullptr);
            CUDA_TRY(cudaGetLastError());
            ix->prof.other_launches += 5;  // score-all + radix sort passes + emit
        }
        return FSGPU_OK;
    }

    const int qb_max = std::max(1, env_int("FSGPU_SCAN_QB", 4));
    uint32_t done = 0;
    while (done < batch) {
        const uint32_t left = batch - done;
        int qb = 1;
        for (int c : {8, 4, 2, 1})
            if (c <= qb_max && (uint32_t)c <= left) {
                qb = c;
                break;
            }
        ScanPlan plan;
        int rc = make_plan(ix, k, qb, &plan);
        if (rc) return rc;
        qb = plan.qb;
        CUDA_TRY(ix->ws_partial.reserve((size_t)plan.grid * qb * k * 8));
        ScanArgs a{};
        a.slab = ix->d_slab;
        a.tombstones = ix->d_excl ? ix->d_excl : ix->d_tomb;
        a.queries = d_queries + (size_t)done * ix->dim;
        a.n_rows = ix->n_rows;
        a.row_base = ix->row_base;
        a.dim = ix->dim;
        a.k = k;
        a.cap = plan.cap;
        a.sync_every = plan.sync_every;
        a.reduce_order = ix->reduce_order;
        a.tail_fma = ix->tail_fma;
        a.allow_packed = env_int("FSGPU_SCAN_PACKED", 1);
        a.partial = ix->ws_partial.as<uint64_t>();
        a.error_flag = ix->d_error;
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        if (ix->profiling) {
            if (!ix->ev_free.empty()) {
                ev = ix->ev_free.back();
                ix->ev_free.pop_back();
            } else {
                CUDA_TRY(cudaEventCreate(&ev.first));
                CUDA_TRY(cudaEventCreate(&ev.second));
            }
            CUDA_TRY(cudaEventRecord(ev.first, stream));
        }
        plan.kernel<<<plan.grid, kScanThreads, plan.smem, stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        if (ix->profiling) {
            CUDA_TRY(cudaEventRecord(ev.second, stream));
            ix->ev_pending.push_back(ev);
        }
        ix->prof.scan_launches += 1;
        ix->prof.merge_launches += 1;
        ix->prof.scan_bytes += ix->n_rows * ix->dim * 2ull;

        MergeArgs m{};
        m.keys = a.partial;
        m.list_stride = (uint64_t)qb * k;
        m.query_stride = k;
        m.n_lists = (uint32_t)plan.grid;
        m.k_in = k;
        m.k_out = k;
        m.cap = plan.cap;
        m.out_keys = d_out_keys ? d_out_keys + (size_t)done * k : nullptr;
        m.out_hits = d_out_hits ? d_out_hits + (size_t)done * k : nullptr;
        m.out_counts = d_out_counts ? d_out_counts + done : nullptr;
        m.slab = ix->slab_any();
        m.slab_is_f32 = ix->is_f32();
        m.queries = a.queries;
        m.n_rows = ix->n_rows;
        m.row_base = ix->row_base;
        m.dim = ix->dim;
        m.reduce_order = ix->reduce_order;
        m.tail_fma = ix->tail_fma;
        m.error_flag = ix->d_error;
        rc = launch_merge(m, (uint32_t)qb, stream);
        if (rc) return rc;
        done += (uint32_t)qb;
    }
    return FSGPU_OK;
}

// ─── batched tensor-core path ───────────────────────────────────────────────────────────────
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// TMA descriptor of a row-major [rows, dim] f16 matrix read as [128 rows x 64 elements] boxes with
// the 128-byte swizzle the UMMA shared-memory descriptors expect.
bool fsgpu_tma_available() { return encode_tiled_fn() != nullptr; }

bool make_f16_tile_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t dim, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rows == 0) return false;
    const cuuint64_t gdim[2] = {dim, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kMmaKBlock, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_tile_map_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t cols, uint32_t elem_bytes,
                      uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rows == 0 || (elem_bytes != 2 && elem_bytes != 4) || box_cols * elem_bytes != 128) return false;
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * elem_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
               const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The same boxes over a row-major [rows, dim] matrix of 8-bit codes: 128 codes per 128-byte row.
static bool make_u8_tile_map(CUtensorMap* tm, const void* base, uint64_t rows, uint32_t dim) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rows == 0) return false;
    const cuuint64_t gdim[2] = {dim, rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)dim};
    const cuuint32_t box[2] = {128, 128};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), gdim, gstride, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Slab statistics + TMA descriptor; decides whether the batched tensor-core path may serve this
// index (otherwise every batch runs on the exact CUDA-core kernels).
static int index_finish_setup(fsgpu_index* ix) {
    ix->mma_ok = false;
    if (ix->d_slab_f32) return FSGPU_OK;  // f32-quantised slab: scored exactly row by row, no tensor-core forms
    if (ix->n_rows == 0 || ix->dim % kMmaKBlock != 0 || ix->dim > kMmaMaxDim ||
        ix->n_rows > 0x7FFFFF00ull || (reinterpret_cast<uintptr_t>(ix->d_slab) & 15u) != 0)
        return FSGPU_OK;
    uint32_t* d_stats = nullptr;
    CUDA_TRY(cudaMalloc(&d_stats, 16));
    CUDA_TRY(cudaMemsetAsync(d_stats, 0, 16, ix->stream));
    const int grid = (int)std::min<uint64_t>((ix->n_rows + 7) / 8, (uint64_t)ix->num_sms * 16);
    slab_stats_kernel<<<grid, 256, 0, ix->stream>>>(ix->d_slab, ix->n_rows, ix->dim, d_stats);
    uint32_t stats[4] = {0, 0, 0, 0};
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, ix->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) {
        cudaFree(d_stats);
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: slab statistics failed: %s", cudaGetErrorString(e));
    }
    float norm;
    memcpy(&norm, &stats[0], 4);
    ix->max_row_norm = norm * 1.0001f;
    const bool finite = stats[1] == 0 && std::isfinite(ix->max_row_norm);
    ix->mma_ok = finite && make_f16_tile_map(&ix->tm_slab, ix->d_slab, ix->n_rows, ix->dim);

    // int8 codes for the kind::i8 form (opt-in: half as many bytes again in HBM)
    ix->i8_ok = false;
    if (ix->mma_ok && ix->dim % 128 == 0 && ix->want_i8 && env_int("FSGPU_MMA_I8", 1) != 0) {
        // largest |element|: f16 magnitude bits -> f32
        const uint16_t hb = (uint16_t)stats[2];
        const uint32_t exp = (hb >> 10) & 0x1F, man = hb & 0x3FF;
        const float max_abs = exp == 0 ? std::ldexp((float)man, -24) : std::ldexp((float)(man | 0x400), (int)exp - 25);
        if (max_abs > 0.0f) {
            const float scale = 127.0f / max_abs;  // simd.rs:1850
            ix->i8_sx = max_abs / 127.0f;
            e = ix->d_slab_i8.reserve(ix->n_rows * ix->dim);
            if (e == cudaErrorMemoryAllocation) {  // no room for the codes: the f16 forms serve everything
                cudaGetLastError();
                cudaFree(d_stats);
                return FSGPU_OK;
            }
            if (e == cudaSuccess) e = cudaMemsetAsync(d_stats, 0, 16, ix->stream);
            if (e == cudaSuccess) {
                quantize_slab_i8_kernel<<<grid, 256, 0, ix->stream>>>(ix->d_slab, ix->n_rows, ix->dim, scale, ix->i8_sx,
                                                                      ix->d_slab_i8.as<int8_t>(), d_stats);
                e = cudaGetLastError();
            }
            if (e == cudaSuccess) e = cudaMemcpyAsync(stats, d_stats, 4, cudaMemcpyDeviceToHost, ix->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
            if (e != cudaSuccess) {
                cudaFree(d_stats);
                return fail(FSGPU_ERR_SUBSYSTEM, "gpu: int8 slab build failed: %s", cudaGetErrorString(e));
            }
            float ex;
            memcpy(&ex, &stats[0], 4);
            ix->i8_max_ex = ex * 1.0001f;
            ix->i8_ok = std::isfinite(ix->i8_max_ex) &&
                        make_u8_tile_map(&ix->tm_slab_i8, ix->d_slab_i8.p, ix->n_rows, ix->dim);
        }
    }
    cudaFree(d_stats);
    return FSGPU_OK;
}

static int search_exact_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                               uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                               cudaStream_t stream);

// Device-side redo of the queries a batched search flagged (ws_redo[b] != 0; d_error[1] != 0 iff any):
// exact CUDA-core scan + merge launches on the same stream that exit at once when nothing is flagged —
// the call needs no host round trip.  `limit`: with more flagged queries than this the launches do
// nothing and count a bail in d_error[2] (a synchronous caller then re-runs the batch in the f16 form).
static int enqueue_redo_locked(const fsgpu_index* ix, const float* d_queries, uint32_t sub, uint32_t k,
                               uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                               cudaStream_t stream, uint32_t limit) {
    ScanPlan plan;
    int rc = make_plan(ix, k, 1, &plan);
    if (rc) return rc;
    const int grid = std::min(plan.grid, ix->num_sms * 2);
    const size_t slot_bytes = (size_t)grid * k * 8;
    const uint32_t per = (uint32_t)std::max<size_t>(1, std::min<size_t>({(size_t)kRedoMaxPerLaunch, (size_t)sub,
                                                                         ((size_t)128 << 20) / slot_bytes}));
    const uint32_t worst = std::min(sub, limit);
    const uint32_t rounds = std::max(1u, (worst + per - 1) / per);
    CUDA_TRY(ix->ws_redo_partial.reserve((size_t)per * slot_bytes));
    CUDA_TRY(ix->ws_redo_slots.reserve((size_t)(per + 1) * 4));
    for (uint32_t r = 0; r < rounds; ++r) {
        ScanArgs a{};
        a.redo.flags = ix->ws_redo.as<uint32_t>();
        a.redo.any = ix->d_error + 1;
        a.redo.n = sub;
        a.redo.first = r * per;
        a.redo.max = per;
        a.redo.limit = limit;
        a.redo.bail = ix->d_error + 2;
        a.redo.slots = ix->ws_redo_slots.as<uint32_t>();
        a.redo.round_n = ix->ws_redo_slots.as<uint32_t>() + per;
        a.redo.served = ix->d_error + 3;
        a.slab = ix->d_slab;
        a.tombstones = ix->d_excl ? ix->d_excl : ix->d_tomb;
        a.queries = d_queries;
        a.n_rows = ix->n_rows;
        a.row_base = ix->row_base;
        a.dim = ix->dim;
        a.k = k;
        a.cap = plan.cap;
        a.sync_every = plan.sync_every;
        a.reduce_order = ix->reduce_order;
        a.tail_fma = ix->tail_fma;
        a.allow_packed = env_int("FSGPU_SCAN_PACKED", 1);
        a.partial = ix->ws_redo_partial.as<uint64_t>();
        a.error_flag = ix->d_error;
        plan.kernel<<<grid, kScanThreads, plan.smem, stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        MergeArgs m{};
        m.keys = a.partial;
        m.list_stride = k;
        m.query_stride = (uint64_t)grid * k;
        m.n_lists = (uint32_t)grid;
        m.k_in = k;
        m.k_out = k;
        m.cap = plan.cap;
        m.out_keys = d_out_keys;
        m.out_hits = d_out_hits;
        m.out_counts = d_out_counts;
        m.slab = ix->slab_any();
        m.slab_is_f32 = ix->is_f32();
        m.queries = d_queries;
        m.n_rows = ix->n_rows;
        m.row_base = ix->row_base;
        m.dim = ix->dim;
        m.reduce_order = ix->reduce_order;
        m.tail_fma = ix->tail_fma;
        m.error_flag = ix->d_error;
        m.redo_slots = a.redo.slots;
        m.redo_round_n = a.redo.round_n;
        rc = launch_merge(m, per, stream);
        if (rc) return rc;
        ix->prof.other_launches += 2;
    }
    return FSGPU_OK;
}

// Flags of the call in flight -> pinned host words.  A synchronous entry point reads them after its own
// synchronisation; a stream-asynchronous one leaves them for the next call (consume_flags_locked).
static int copy_flags_locked(const fsgpu_index* ix, cudaStream_t stream) {
    CUDA_TRY(cudaMemcpyAsync(ix->h_flags, ix->d_error, 16, cudaMemcpyDeviceToHost, stream));
    CUDA_TRY(cudaEventRecord(ix->ev_flags, stream));
    ix->flags_pending = true;
    return FSGPU_OK;
}

// Book-keeping from the flags of a finished call: redo accounting, and the adaptive choice of form —
// when the int8 form sent more than a few queries to the exact path (its candidate lists overflowed:
// a corpus whose score spread is narrower than the int8 bound), the next calls take the f16 form.
static void consume_flags_locked(const fsgpu_index* ix, bool wait) {
    if (!ix->flags_pending) return;
    if (wait) {
        if (cudaEventSynchronize(ix->ev_flags) != cudaSuccess) return;
    } else if (cudaEventQuery(ix->ev_flags) != cudaSuccess) {
        cudaGetLastError();  // not ready: look again at the next call
        return;
    }
    ix->flags_pending = false;
    ix->prof.redo_queries += ix->h_flags[3];
    if (ix->h_flags[3] > 4u || ix->h_flags[2] != 0u) ix->i8_penalty = 32;
}

// Start of every entry point that touches the index's workspaces (ix->mu held): take in the flags of the
// last asynchronous call, order this call's stream after the previous call's work (the workspaces and
// flag words are shared between streams), clear the flag words.
static int begin_call_locked(const fsgpu_index* ix, cudaStream_t s, bool sync_caller) {
    consume_flags_locked(ix, false);
    if (ix->have_last && ix->last_stream != s) CUDA_TRY(cudaStreamWaitEvent(s, ix->ev_last, 0));
    ix->sync_caller = sync_caller;
    CUDA_TRY(cudaMemsetAsync(ix->d_error, 0, 16, s));
    return FSGPU_OK;
}
// End of the enqueue: flags -> pinned host words, and the event the next call (on any stream) waits on.
static int end_call_locked(const fsgpu_index* ix, cudaStream_t s) {
    int rc = copy_flags_locked(ix, s);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(ix->ev_last, s));
    ix->last_stream = s;
    ix->have_last = true;
    if (ix->i8_penalty > 0) --ix->i8_penalty;
    return FSGPU_OK;
}
// A synchronous entry point after end_call_locked: wait, then read the verdict of the call.
// *retry_f16 = the int8 form bailed (too many overflowed lists): run the call again in the f16 form.
static int finish_sync_call_locked(const fsgpu_index* ix, cudaStream_t s, bool* retry_f16) {
    CUDA_TRY(cudaStreamSynchronize(s));
    const uint32_t violated = ix->h_flags[0], bails = ix->h_flags[2];
    consume_flags_locked(ix, true);
    if (violated) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: top-k candidate buffer contract violated");
    if (retry_f16) *retry_f16 = bails != 0u;
    return FSGPU_OK;
}

// The sample cascade of one batched search (mma_scan_kernels.cuh header): tiles per level, the
// selection rank k' used for intermediate gates, and the per-query list capacity.
struct MmaCascade {
    uint64_t n_tiles = 0, t0 = 0, t1 = 0;  // t1 == 0: two levels only; t0 == n_tiles: one level
    uint32_t k_sel = 16, tile_rows = kMmaN;
    bool group_max = false;  // level 0 keeps the best score of every 8-row group of a sample 8x as large
    double slack = 6.0, random_part = 0.0;  // expected gate-clearing rows per query and level
    // capacity of one (CTA, query) list when `g` CTAs (or CTA pairs) share a query block
    // (each CTA keeps `parts` lists per query, one per epilogue warp of a lane quarter: equal shares of a tile's columns)
    uint32_t list_cap(uint32_t g, uint32_t parts) const {
        // level 0 keeps every score of its tiles (or one score per 8 rows: group_max)
        const uint64_t dump = (t0 + g - 1) / g * (tile_rows / (group_max ? 8 * parts : parts));
        const uint64_t rnd = (uint64_t)std::min(slack * random_part / ((double)parts * g), 4194304.0) + 64;
        return (uint32_t)((std::max(dump, rnd) + 63) / 64 * 64);
    }
};

// FSGPU_MMA_TRACE=1: CUDA-event timestamps between the launches of one batched search, printed to
// stderr after the call's synchronisation (a debugging aid: where the fixed per-call time goes).
struct MmaTrace {
    bool on = false;
    cudaStream_t stream = nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    void mark(const char* what) {
        if (!on) return;
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, stream);
        marks.emplace_back(what, e);
    }
    void report(uint32_t batch, uint64_t rows) {
        if (!on || marks.size() < 2) return;
        fprintf(stderr, "[fsgpu trace] batch=%u rows=%llu:", batch, (unsigned long long)rows);
        for (size_t i = 1; i < marks.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
            fprintf(stderr, " %s=%.1fus", marks[i].first, ms * 1e3f);
        }
        float total = 0.f;
        cudaEventElapsedTime(&total, marks.front().second, marks.back().second);
        fprintf(stderr, " total=%.1fus\n", total * 1e3f);
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }
};

static MmaCascade plan_cascade(uint64_t n_rows, uint32_t k, uint32_t tile_rows, bool i8) {
    // Appends per query: level 0 keeps all t0*tile_rows sample scores (deterministic); level 1
    // expects k' * t1/t0 and level 2 k' * n_tiles/t1, both minimised by t1 = sqrt(t0 * n_tiles) at
    // L = k' * sqrt(n_tiles/t0).
    MmaCascade c;
    c.tile_rows = tile_rows;
    c.n_tiles = (n_rows + tile_rows - 1) / tile_rows;
    // k' >= k keeps the sample thresholds valid; k' >= 8 keeps their spread small (the number of
    // corpus rows above a sample's k'-th best is ~Gamma(k') times its mean: 6x slack is never reached)
    c.k_sel = std::max(k, 8u);
    c.slack = std::max(2, env_int("FSGPU_MMA_LIST_SLACK", 6));
    // t0 minimises the total number of appended candidates per query, t0*tile_rows + 2 k' sqrt(n_tiles/t0)
    // (appends, not MMAs, are what the sample levels cost: ~0.1 us of warp time each)
    const double t0_bal = std::pow(c.k_sel * std::sqrt((double)c.n_tiles) / tile_rows, 2.0 / 3.0);
    const uint64_t t0_min = ((uint64_t)8 * c.k_sel + tile_rows - 1) / tile_rows;
    c.t0 = std::min<uint64_t>(c.n_tiles, std::max<uint64_t>({(uint64_t)2048 / tile_rows, t0_min, (uint64_t)std::ceil(t0_bal)}));
    if (c.n_tiles > 8 * c.t0) {
        c.t1 = (uint64_t)std::llround(std::sqrt((double)c.t0 * (double)c.n_tiles));
        // int8 form: ~3x the rows clear a gate lowered by its wider bound, and every appended row costs the full pass's
        // epilogue and both refine steps — a second sample up to twice as large pays from ~5 M rows on (10 M x 384,
        // batch 1024: sample pass +70 us, exact-gate step -27, full pass -95, refine -70; neutral at 1.25 M rows;
        // four times as large overflows the gate kernel's staging).  FSGPU_MMA_T1_PCT overrides (100 = the f16 rule).
        int t1_pct = env_int("FSGPU_MMA_T1_PCT", 0);
        if (t1_pct <= 0) t1_pct = i8 ? (int)(100.0 * std::min(2.0, std::max(1.0, std::cbrt((double)c.n_tiles / 4883.0)))) : 100;
        c.t1 = c.t1 * (uint64_t)t1_pct / 100;
        c.t1 = std::min(c.n_tiles, std::max(c.t1, 2 * c.t0));
        // the first level only feeds a k'-th-best selection: with one score per 8-row group it samples
        // 8x as many tiles for the same number of appends, and its gate is ~8x tighter
        if (env_int("FSGPU_MMA_GROUP_MAX", 1) != 0) {
            // ... but the next level must still be >= 8x as large, or it too often catches fewer than
            // k' rows above this gate (their count is ~ Poisson(k' * Gamma(k')/k' * t1/t0))
            const uint64_t cap0 = (uint64_t)std::max(1, env_int("FSGPU_MMA_T0_CAP", 8));
            const uint64_t wide = std::min<uint64_t>(cap0 * c.t0, std::max<uint64_t>(c.t0, c.t1 / 8));
            if (wide > c.t0) {
                c.t0 = wide;
                c.group_max = true;
            }
        }
        c.random_part = std::max((double)c.k_sel * (double)c.t1 / (double)c.t0,
                                 (double)c.k_sel * (double)c.n_tiles / (double)c.t1);
    } else if (c.t0 < c.n_tiles) {
        c.random_part = (double)c.k_sel * (double)c.n_tiles / (double)c.t0;
    }
    return c;
}

// Up to num_sms*128 queries in ONE full pass over the slab (+ ~1.5 % of sample passes):
// prep -> [scan sample, gate] x 2 -> scan all -> exact refine.  Synchronises `stream` once (to learn
// which queries, if any, must be redone exactly).
static int search_mma_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                             uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                             cudaStream_t stream) {
    // the int8 form pays while the candidate volume stays small (k <= 32: top-50 loses) and the pass is
    // long enough to amortise its extra refine step (>= 1 M rows: at 1.25 M rows it is 5 % ahead, 0.82 vs
    // 0.87 ms at batch 1024) — profiles/r01_sweep_i8_form.txt
    const bool i8 = !ix->force_f16 && ix->i8_penalty == 0 && ix->i8_ok && env_int("FSGPU_MMA_I8", 1) != 0 &&
                    k <= (uint32_t)std::max(0, env_int("FSGPU_I8_MAX_K", 32)) &&
                    ix->n_rows >= (uint64_t)std::max(0, env_int("FSGPU_I8_MIN_ROWS", 1000000));
    const uint32_t n_kb = ix->dim / (i8 ? 128 : kMmaKBlock);  // 128-byte K-blocks
    const size_t smem_limit = 227 * 1024;
    // quad form (int8, more than 256 queries): two query blocks per CTA share every B stage
    const bool quad = i8 && env_int("FSGPU_MMA_QUAD", 1) != 0 && env_int("FSGPU_MMA_PAIR", 1) != 0 && ix->num_sms >= 2 &&
                      batch > 2 * kMmaM && mma_scan_smem_bytes(2 * n_kb, 4) <= smem_limit;
    const uint32_t a_tiles = quad ? 2 * n_kb : n_kb;  // 16 KiB query tiles resident per CTA
    const size_t fixed = mma_scan_smem_bytes(a_tiles, 0);
    uint32_t n_stages = (uint32_t)std::min<size_t>(kMmaMaxStages, (smem_limit - fixed) / kMmaTileBytes);
    const int stage_cap = env_int("FSGPU_MMA_STAGES", 0);
    if (stage_cap >= 2) n_stages = std::min<uint32_t>(n_stages, (uint32_t)stage_cap);
    if (n_stages < 2) return fail(FSGPU_ERR_INVALID_CONFIG, "batched scan does not fit in shared memory (dim=%u)", ix->dim);
    const size_t smem = mma_scan_smem_bytes(a_tiles, n_stages);
    // CTA-pair form (cta_group::2, 256 queries x 256 rows per MMA) unless FSGPU_MMA_PAIR=0
    // ... and unless the batch is a single 128-query block: that regime is HBM-bound and the
    // single-CTA form streams it faster (6.7 vs 5.8 TB/s at 10 M x 384, profiles/r01_sweep_mma_v6.txt)
    const bool pair = env_int("FSGPU_MMA_PAIR", 1) != 0 && ix->num_sms >= 2 && batch > kMmaM;
    // quad form: 16 epilogue warps (FSGPU_MMA_QUAD_WARPS=8: the first form, two per TMEM lane quarter)
    const uint32_t quad_warps = quad && env_int("FSGPU_MMA_QUAD_WARPS", 16) >= 16 ? 16u : 8u;
    const uint32_t parts = quad ? quad_warps / 4 : 2;  // private candidate lists per (CTA, query)
    const uint32_t scan_threads = quad ? 64 + 32 * quad_warps : kMmaThreads;
    auto scan_kernel = quad ? (quad_warps == 16 ? mma_scan_quad_kernel<16> : mma_scan_quad_kernel<8>)
                     : pair ? (i8 ? mma_scan_pair_kernel<true> : mma_scan_pair_kernel<false>)
                            : (i8 ? mma_scan_kernel<true> : mma_scan_kernel<false>);
    CUDA_TRY(cudaFuncSetAttribute(scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const uint32_t units = pair ? (uint32_t)ix->num_sms / 2 : (uint32_t)ix->num_sms;  // CTAs or CTA pairs
    const uint32_t unit_queries = quad ? 4 * kMmaM : pair ? 2 * kMmaM : kMmaM;
    MmaCascade cas = plan_cascade(ix->n_rows, k, pair ? kPairN : kMmaN, i8);
    // the int8 bound lowers every gate by ~0.25-0.5 sigma of the score distribution: ~4x the rows clear it
    if (i8) cas.random_part *= (double)std::max(1, env_int("FSGPU_I8_LIST_SCALE", 4));
    const uint32_t fin_cap = cand_capacity(k);
    // shared-memory staging of the candidate lists, sized from the expected list lengths (smaller
    // CTAs -> one wave of gate/refine CTAs); longer lists take the kernels' unstaged path
    auto stage_cap_for = [](double expected, uint32_t limit) {
        return std::min(limit, std::max(1024u, host_next_pow2((uint32_t)std::min(2.5 * expected + 256.0, 1.0e6))));
    };
    const uint32_t gate0_cap = stage_cap_for((double)cas.t0 * cas.tile_rows / (cas.group_max ? 8.0 : 1.0) / 2.5,
                                             kMmaStageScores);  // exact dump size
    const uint32_t gate1_cap = stage_cap_for(cas.random_part, kMmaStageScores);
    const uint32_t refine_cap = stage_cap_for(cas.t0 < cas.n_tiles ? cas.random_part : (double)cas.t0 * cas.tile_rows / 2.5,
                                              kMmaStagePairs);
    const size_t gate_smem_max = mma_stage_smem_bytes(kMmaStageScores, false);
    const size_t refine_smem = (size_t)fin_cap * 8 + 16 + mma_stage_smem_bytes(refine_cap, true) + (size_t)ix->dim * 4 +
                               (size_t)kMmaTopListCap * 4;
    CUDA_TRY(cudaFuncSetAttribute(mma_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gate_smem_max));
    CUDA_TRY(cudaFuncSetAttribute(mma_refine_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)refine_smem));

    const uint32_t max_queries = units * unit_queries;
    // a synchronous caller can re-run the batch in the f16 form when the int8 lists overflow for more than
    // a few queries (100x tighter bound); a stream-asynchronous one has the device redo everything flagged
    const uint32_t redo_limit = (i8 && ix->sync_caller) ? 4u : 0xFFFFFFFFu;
    for (uint32_t done = 0; done < batch; done += max_queries) {
        const uint32_t sub = std::min(max_queries, batch - done);
        if (done) CUDA_TRY(cudaMemsetAsync(ix->d_error + 1, 0, 4, stream));  // redo_any of this sub-batch (begin_call cleared the first)
        const uint32_t n_units = (sub + unit_queries - 1) / unit_queries;  // query blocks or query pairs
        const uint32_t n_qb = quad ? 4 * n_units : pair ? 2 * n_units : n_units;  // 128-query blocks incl. padding
        const uint32_t g = units / n_units;
        const uint32_t grid = g * n_units * (pair ? 2 : 1);
        const uint32_t slots = n_qb * kMmaM;
        CUDA_TRY(ix->ws_qhat.reserve((size_t)slots * ix->dim * 2));
        CUDA_TRY(ix->ws_margin.reserve((size_t)slots * 4));
        CUDA_TRY(ix->ws_gate.reserve((size_t)slots * 4));
        CUDA_TRY(ix->ws_redo.reserve((size_t)slots * 4));
        const uint32_t cap = cas.list_cap(g, parts);
        const size_t lists_per_cta = quad ? 2 * parts : 2;  // (sub-block x) column part
        CUDA_TRY(ix->ws_cand.reserve((size_t)grid * lists_per_cta * kMmaM * cap * sizeof(MmaCand)));
        CUDA_TRY(ix->ws_cand_count.reserve((size_t)grid * lists_per_cta * kMmaM * 4));
        if (i8) {
            CUDA_TRY(ix->ws_qscale.reserve((size_t)slots * 4));
            if (ix->tm_qhat_i8_ptr != ix->ws_qhat.p || ix->tm_qhat_i8_rows != slots) {
                if (!make_u8_tile_map(&ix->tm_qhat_i8, ix->ws_qhat.p, slots, ix->dim))
                    return fail(FSGPU_ERR_SUBSYSTEM, "gpu: cuTensorMapEncodeTiled failed for the query tile");
                ix->tm_qhat_i8_ptr = ix->ws_qhat.p;
                ix->tm_qhat_i8_rows = slots;
            }
        } else if (ix->tm_qhat_ptr != ix->ws_qhat.p || ix->tm_qhat_rows != slots) {
            if (!make_f16_tile_map(&ix->tm_qhat, ix->ws_qhat.p, slots, ix->dim))
                return fail(FSGPU_ERR_SUBSYSTEM, "gpu: cuTensorMapEncodeTiled failed for the query tile");
            ix->tm_qhat_ptr = ix->ws_qhat.p;
            ix->tm_qhat_rows = slots;
        }
        const float* q = d_queries + (size_t)done * ix->dim;
        MmaTrace trace;
        trace.on = env_int("FSGPU_MMA_TRACE", 0) != 0;
        trace.stream = stream;
        trace.mark("start");
        if (i8)
            mma_prep_queries_i8_kernel<<<slots, 128, 0, stream>>>(q, sub, ix->dim, ix->max_row_norm, ix->i8_max_ex,
                                                                  ix->i8_sx, ix->ws_qhat.as<int8_t>(),
                                                                  ix->ws_margin.as<float>(), ix->ws_qscale.as<float>(),
                                                                  ix->ws_redo.as<uint32_t>());
        else
            mma_prep_queries_kernel<<<slots, 128, 0, stream>>>(q, sub, ix->dim, ix->max_row_norm, ix->ws_qhat.as<__half>(),
                                                               ix->ws_margin.as<float>(), ix->ws_redo.as<uint32_t>());
        CUDA_TRY(cudaGetLastError());
        ix->prof.other_launches += 1;

        MmaScanArgs a{};
        a.n_rows = ix->n_rows;
        a.row_base = ix->row_base;
        a.tombstones = ix->d_excl ? ix->d_excl : ix->d_tomb;
        a.n_kblocks = n_kb;
        a.n_qblocks = n_qb;
        a.ctas_per_qblock = g;
        a.batch = sub;
        a.n_stages = n_stages;
        a.redo = ix->ws_redo.as<uint32_t>();
        a.qscale = i8 ? ix->ws_qscale.as<float>() : nullptr;
        a.cand = ix->ws_cand.as<MmaCand>();
        a.cand_count = ix->ws_cand_count.as<uint32_t>();
        a.cap = cap;
        // pacing of the pairs that share a tile stream (only meaningful with several query pairs)
        // (int8 quad form at 10 M x 384, batch 1024: lead 32 / 16 / 8 / 4 / 2 -> 4.24 / 4.16 / 3.90 / 3.87 / 3.87 GB read from
        // DRAM for 3.84 GB of codes, the same 2.35 ms)
        // f16 pair form: lead 16 -> 4: 12.08 -> 7.70 GB for 7.68 GB of slab and 5.12 -> 4.92 ms (less DRAM power, higher clock)
        const uint32_t lead = (uint32_t)std::max(0, env_int("FSGPU_MMA_LEAD", 4));
        const bool paced = pair && n_units > 1 && lead > 0;
        // one zeroed block per sub-batch, ONE memset: [pacing counters of the three passes | the exact-gate
        // refine's per-query counts] (every separate memset is ~2 us of stream time)
        const size_t progress_bytes = (((size_t)g * n_units * 4) + 15) & ~(size_t)15;
        const size_t zero_bytes = 3 * progress_bytes + (size_t)slots * 4;
        CUDA_TRY(ix->ws_progress.reserve(zero_bytes));
        CUDA_TRY(cudaMemsetAsync(ix->ws_progress.p, 0, zero_bytes, stream));
        auto progress_of = [&](int pass) { return reinterpret_cast<uint32_t*>(static_cast<char*>(ix->ws_progress.p) + pass * progress_bytes); };
        uint32_t* i8_cnt = reinterpret_cast<uint32_t*>(static_cast<char*>(ix->ws_progress.p) + 3 * progress_bytes);
        a.lead = lead;
        MmaGateArgs ga{};
        ga.lists = MmaLists{a.cand, a.cand_count, n_qb, g, cap, quad ? 2u : pair ? 1u : 0u, parts};
        ga.margin2 = ix->ws_margin.as<float>();
        ga.redo = a.redo;
        ga.gate = ix->ws_gate.as<float>();
        ga.k_sel = cas.k_sel;

        // sample levels: strided tiles, each level's k'-th best gates the next
        const uint64_t level_tiles[2] = {cas.t0 < cas.n_tiles ? cas.t0 : 0, cas.t1};
        bool have_gate = false;
        uint64_t stride1 = 0;  // tile stride of the second sample level (its lists can be carried into the full pass)
        for (int lvl = 0; lvl < 2; ++lvl) {
            if (!level_tiles[lvl]) continue;
            a.tile_stride = env_int("FSGPU_MMA_SAMPLE_CONTIG", 0) ? 1 : cas.n_tiles / level_tiles[lvl];
            if (lvl == 1) stride1 = a.tile_stride;
            a.tile_count = level_tiles[lvl];
            a.dump_group_max = (lvl == 0 && cas.group_max) ? 1u : 0u;
            a.gate = have_gate ? ix->ws_gate.as<float>() : nullptr;
            a.progress = paced ? progress_of(lvl) : nullptr;
            trace.mark(lvl ? "gate0" : "prep");
            scan_kernel<<<grid, scan_threads, smem, stream>>>(i8 ? ix->tm_qhat_i8 : ix->tm_qhat, i8 ? ix->tm_slab_i8 : ix->tm_slab, a);
            CUDA_TRY(cudaGetLastError());
            trace.mark(lvl ? "scan1" : "scan0");
            ga.stage_cap = have_gate ? gate1_cap : gate0_cap;
            ga.keep_prev = have_gate ? 1u : 0u;
            const bool exact_gate = i8 && lvl == 1 && env_int("FSGPU_I8_EXACT_GATE", 1) != 0;
            if (!exact_gate) {  // (the exact gate below supersedes this level's approximate one)
                mma_gate_kernel<<<sub, 256, mma_stage_smem_bytes(ga.stage_cap, false), stream>>>(ga);
                CUDA_TRY(cudaGetLastError());
                ix->prof.other_launches += 1;
            }
            ix->prof.other_launches += 1;
            have_gate = true;
            if (exact_gate) {
                // int8 form: tighten the gate of the full pass with an EXACT k-th best of this sample
                CUDA_TRY(ix->ws_i8_top.reserve((size_t)slots * k * 8));

                MmaRefineArgs ri{};
                ri.lists = ga.lists;
                ri.margin2 = ga.margin2;
                ri.redo = ix->ws_redo.as<uint32_t>();
                ri.k = k;
                ri.buf_cap = fin_cap;
                ri.stage_cap = refine_cap;
                ri.slab = ix->d_slab;
                ri.queries = q;
                ri.n_rows = ix->n_rows;
                ri.row_base = ix->row_base;
                ri.dim = ix->dim;
                ri.reduce_order = ix->reduce_order;
                ri.tail_fma = ix->tail_fma;
                ri.out_keys = ix->ws_i8_top.as<uint64_t>();
                ri.out_counts = i8_cnt;
                ri.error_flag = ix->d_error;
                ri.redo_any = ix->d_error + 1;
                ri.intermediate = 1;
                mma_refine_kernel<<<sub, 256, refine_smem, stream>>>(ri);
                CUDA_TRY(cudaGetLastError());
                mma_exact_gate_kernel<<<(sub + 255) / 256, 256, 0, stream>>>(ix->ws_i8_top.as<uint64_t>(),
                                                                           i8_cnt, k, sub,
                                                                           ix->ws_margin.as<float>(), ix->ws_gate.as<float>());
                CUDA_TRY(cudaGetLastError());
                ix->prof.other_launches += 2;
            }
        }
        // the full pass
        a.dump_group_max = 0;
        a.tile_stride = 1;
        a.tile_count = cas.n_tiles;
        a.dbg = (uint32_t)env_int("FSGPU_MMA_DBG", 0);  // timing experiments on the full pass (wrong results)
        long long* d_ts = nullptr;
        if (trace.on && quad && env_int("FSGPU_MMA_TS", 0) != 0) {  // debugging aid: pipeline timestamps of CTA 0
            CUDA_TRY(cudaMalloc(&d_ts, (8 + 64 * 2 * 72) * sizeof(long long)));
            CUDA_TRY(cudaMemsetAsync(d_ts, 0, (8 + 64 * 2 * 72) * sizeof(long long), stream));
            a.ts = d_ts;
        }
        a.gate = have_gate ? ix->ws_gate.as<float>() : nullptr;
        a.progress = paced ? progress_of(2) : nullptr;
        // quad form: the second sample level's tiles are not scanned twice — its lists stay in place, every thread
        // compacts its list against the final gate and the full pass appends behind it (FSGPU_MMA_CARRY=0: off)
        const bool carry = quad && level_tiles[1] && stride1 >= 1 && cas.n_tiles < (1ull << 31) && level_tiles[1] < cas.n_tiles &&
                           env_int("FSGPU_MMA_CARRY", 1) != 0;
        if (carry) {
            a.skip_stride = (uint32_t)stride1;
            a.skip_count = (uint32_t)level_tiles[1];
            a.carry = 1;
        }
        const double scanned = carry ? (double)(cas.n_tiles - level_tiles[1]) / (double)cas.n_tiles : 1.0;
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        if (ix->profiling) {
            if (!ix->ev_free.empty()) {
                ev = ix->ev_free.back();
                ix->ev_free.pop_back();
            } else {
                CUDA_TRY(cudaEventCreate(&ev.first));
                CUDA_TRY(cudaEventCreate(&ev.second));
            }
            CUDA_TRY(cudaEventRecord(ev.first, stream));
        }
        trace.mark("gate_last");
        scan_kernel<<<grid, scan_threads, smem, stream>>>(i8 ? ix->tm_qhat_i8 : ix->tm_qhat, i8 ? ix->tm_slab_i8 : ix->tm_slab, a);
        CUDA_TRY(cudaGetLastError());
        trace.mark("scan_full");
        if (ix->profiling) {
            CUDA_TRY(cudaEventRecord(ev.second, stream));
            ix->ev_pending.push_back(ev);
        }
        ix->prof.scan_launches += 1;
        ix->prof.scan_bytes += (uint64_t)((double)(ix->n_rows * ix->dim * (i8 ? 1ull : 2ull)) * scanned);
        ix->prof.mma_launches += 1;
        ix->prof.i8_launches += i8 ? 1 : 0;
        ix->prof.quad_launches += quad ? 1 : 0;
        ix->prof.pair_launches += (pair && !quad) ? 1 : 0;
        ix->prof.mma_flops += 2.0 * (double)slots * (double)ix->n_rows * (double)ix->dim * scanned;
        ix->prof.merge_launches += 1;  // refine

        MmaRefineArgs r{};
        r.lists = ga.lists;
        r.margin2 = ga.margin2;
        r.redo = ix->ws_redo.as<uint32_t>();
        r.k = k;
        r.buf_cap = fin_cap;
        r.stage_cap = refine_cap;
        r.slab = ix->d_slab;
        r.queries = q;
        r.n_rows = ix->n_rows;
        r.row_base = ix->row_base;
        r.dim = ix->dim;
        r.reduce_order = ix->reduce_order;
        r.tail_fma = ix->tail_fma;
        r.out_keys = d_out_keys ? d_out_keys + (size_t)done * k : nullptr;
        r.out_hits = d_out_hits ? d_out_hits + (size_t)done * k : nullptr;
        r.out_counts = d_out_counts ? d_out_counts + done : nullptr;
        r.error_flag = ix->d_error;
        r.redo_any = ix->d_error + 1;
        mma_refine_kernel<<<sub, 256, refine_smem, stream>>>(r);
        CUDA_TRY(cudaGetLastError());
        trace.mark("refine");

        // flagged queries (non-finite / overflowing components, overflowed candidate lists) are re-run by
        // the exact kernels on the same stream; those launches exit at once when nothing is flagged
        {
            const size_t o = (size_t)done;
            int rc = enqueue_redo_locked(ix, q, sub, k, d_out_keys ? d_out_keys + o * k : nullptr,
                                         d_out_hits ? d_out_hits + o * k : nullptr, d_out_counts ? d_out_counts + o : nullptr,
                                         stream, redo_limit);
            if (rc) return rc;
        }
        trace.mark("redo");
        if (trace.on) {
            CUDA_TRY(cudaStreamSynchronize(stream));
            trace.report(sub, ix->n_rows);
            if (d_ts) {
                std::vector<long long> ts(8 + 64 * 2 * 72);
                CUDA_TRY(cudaMemcpy(ts.data(), d_ts, ts.size() * sizeof(long long), cudaMemcpyDeviceToHost));
                cudaFree(d_ts);
                const long long t0 = ts[8], off1 = ts[1] - ts[0];  // CTA 1's clock minus CTA 0's at the cluster barrier
                fprintf(stderr, "[fsgpu ts] unit: issuer(wait_tempty got issued) | per epilogue warp (CTA 0's, then CTA 1's): tfull_seen hotbits released done  [CTA 0 SM clocks since the first stamp]\n");
                for (int u = 0; u < 128; ++u) {
                    fprintf(stderr, "[fsgpu ts] %3d.%d:", u / 2, u % 2);
                    for (int k = 0; k < 68; ++k) {
                        if (k == 3) continue;
                        const long long v = ts[8 + u * 72 + k];
                        fprintf(stderr, "%s%7lld", (k >= 4 && k % 4 == 0) ? " |" : "", v ? v - t0 - (k >= 36 ? off1 : 0) : -1);
                    }
                    fprintf(stderr, "\n");
                }
            }
        }
    }
    return FSGPU_OK;
}

// Exact top-k for `batch` device-resident queries; outputs may be NULL.  Caller holds ix->mu.
// Small batches stream the slab once per <= 4 queries on the CUDA cores (HBM-bound); batches of
// FSGPU_MMA_MIN_BATCH (default 3) or more take the tensor-core pass (one pass of the slab for the
// whole batch beats two CUDA-core passes from 3 queries on: profiles/r01_sweep_k.txt).  Both give
// identical results.
static int search_gather_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                                uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                                cudaStream_t stream);

// Positions the int8 pass may list per query before the call falls back to the f16 scan.
constexpr uint32_t kI8ListCap = 1u << 13;

// A few queries, int8 pass 1 + exact re-score (scan_kernels.cuh "single-query int8 pass 1"); up to
// four queries share one pass over the codes.  The caller (host API) has checked that the queries
// are finite and synchronises at the end: word 1 of d_error reports a position list that overflowed.
static int search_i8_single_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                                   uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                                   cudaStream_t stream) {
    const uint32_t cap = cand_capacity(k);
    const uint64_t n_tiles = (ix->n_rows + kI8RowsPerIter - 1) / kI8RowsPerIter;
    CUDA_TRY(ix->ws_qhat.reserve((size_t)4 * ix->dim));
    CUDA_TRY(ix->ws_margin.reserve(16));
    CUDA_TRY(ix->ws_qscale.reserve(16));
    CUDA_TRY(ix->ws_redo.reserve(16));
    CUDA_TRY(ix->ws_gate.reserve(4));
    CUDA_TRY(ix->ws_i8_top.reserve((size_t)4 * k * 8));
    CUDA_TRY(ix->ws_gather_pos.reserve((size_t)kI8ListCap * 4));
    CUDA_TRY(ix->ws_gather_count.reserve(4));
    // more than one query per pass makes the dp4a pass ALU-bound (1.12 ms for 2, 1.38 ms for 4 queries at
    // 10 M x 384, against 0.59 ms for one): off by default
    const int qb_max = std::max(1, env_int("FSGPU_I8_QB", 1));
    uint32_t done = 0;
    while (done < batch) {
        const uint32_t left = batch - done;
        const uint32_t qb = (left >= 4 && qb_max >= 4) ? 4u : (left >= 2 && qb_max >= 2) ? 2u : 1u;
        auto kernel = qb == 4 ? scan_i8_kernel<4> : qb == 2 ? scan_i8_kernel<2> : scan_i8_kernel<1>;
        const size_t smem = (size_t)qb * cap * 8 + (size_t)qb * 12 + 16;
        CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kScanThreads, smem));
        if (per_sm < 1) return fail(FSGPU_ERR_INVALID_CONFIG, "int8 scan kernel does not fit (k=%u)", k);
        const uint32_t grid = (uint32_t)std::min<uint64_t>(n_tiles, (uint64_t)ix->num_sms * per_sm);
        CUDA_TRY(ix->ws_approx.reserve((size_t)qb * ix->n_rows * 4));
        CUDA_TRY(ix->ws_partial.reserve((size_t)grid * qb * k * 8));
        const float* q = d_queries + (size_t)done * ix->dim;
        mma_prep_queries_i8_kernel<<<qb, 128, 0, stream>>>(q, qb, ix->dim, ix->max_row_norm, ix->i8_max_ex, ix->i8_sx,
                                                           ix->ws_qhat.as<int8_t>(), ix->ws_margin.as<float>(),
                                                           ix->ws_qscale.as<float>(), ix->ws_redo.as<uint32_t>());
        CUDA_TRY(cudaGetLastError());
        I8ScanArgs a{};
        a.codes = ix->d_slab_i8.as<int8_t>();
        a.tombstones = ix->d_excl ? ix->d_excl : ix->d_tomb;
        a.q_codes = ix->ws_qhat.as<int8_t>();
        a.qscale = ix->ws_qscale.as<float>();
        a.n_rows = ix->n_rows;
        a.row_base = ix->row_base;
        a.dim = ix->dim;
        a.k = k;
        a.cap = cap;
        a.sync_every = std::max<uint32_t>(1, std::min<uint32_t>(16, (cap - k) / (2 * kI8RowsPerIter)));
        a.approx = ix->ws_approx.as<float>();
        a.partial = ix->ws_partial.as<uint64_t>();
        a.error_flag = ix->d_error;
        std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
        if (ix->profiling) {
            if (!ix->ev_free.empty()) {
                ev = ix->ev_free.back();
                ix->ev_free.pop_back();
            } else {
                CUDA_TRY(cudaEventCreate(&ev.first));
                CUDA_TRY(cudaEventCreate(&ev.second));
            }
            CUDA_TRY(cudaEventRecord(ev.first, stream));
        }
        kernel<<<grid, kScanThreads, smem, stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        if (ix->profiling) {
            CUDA_TRY(cudaEventRecord(ev.second, stream));
            ix->ev_pending.push_back(ev);
        }
        ix->prof.scan_launches += 1;
        ix->prof.scan_bytes += ix->n_rows * ix->dim;  // int8 codes
        ix->prof.i8_launches += 1;
        MergeArgs m{};  // approximate top-k of every query of the group
        m.keys = a.partial;
        m.list_stride = (uint64_t)qb * k;
        m.query_stride = k;
        m.n_lists = grid;
        m.k_in = k;
        m.k_out = k;
        m.cap = cap;
        m.out_keys = ix->ws_i8_top.as<uint64_t>();
        m.error_flag = ix->d_error;
        int rc = launch_merge(m, qb, stream);
        if (rc) return rc;
        ix->prof.merge_launches += 1;
        for (uint32_t qi = 0; qi < qb; ++qi) {
            const size_t o = (size_t)done + qi;
            const float* qq = d_queries + o * ix->dim;
            const float* approx = ix->ws_approx.as<float>() + (size_t)qi * ix->n_rows;
            i8_gate_kernel<<<1, kScanThreads, 0, stream>>>(ix->ws_i8_top.as<uint64_t>() + (size_t)qi * k, k, ix->d_slab,
                                                           ix->row_base, ix->dim, qq, ix->ws_margin.as<float>() + qi,
                                                           ix->reduce_order, ix->tail_fma, ix->ws_gate.as<float>(),
                                                           ix->ws_gather_count.as<uint32_t>());
            CUDA_TRY(cudaGetLastError());
            i8_select_kernel<<<(unsigned)std::min<uint64_t>((ix->n_rows + 255) / 256, (uint64_t)ix->num_sms * 16), 256, 0,
                               stream>>>(approx, ix->n_rows, ix->ws_gate.as<float>(), ix->ws_gather_pos.as<uint32_t>(),
                                         kI8ListCap, ix->ws_gather_count.as<uint32_t>(), ix->d_error + 1);
            CUDA_TRY(cudaGetLastError());
            ix->prof.other_launches += 2;
            ix->d_gather_pos = ix->ws_gather_pos.as<uint32_t>();
            ix->d_gather_count = ix->ws_gather_count.as<uint32_t>();
            ix->gather_cap = kI8ListCap;
            rc = search_gather_locked(ix, qq, 1, k, d_out_keys ? d_out_keys + o * k : nullptr,
                                      d_out_hits ? d_out_hits + o * k : nullptr, d_out_counts ? d_out_counts + o : nullptr,
                                      stream);
            ix->d_gather_pos = nullptr;
            ix->d_gather_count = nullptr;
            ix->gather_cap = 0;
            if (rc) return rc;
        }
        ix->prof.other_launches += 1;  // prep
        done += qb;
    }
    return FSGPU_OK;
}

// Large k (FSGPU_SELECT_MIN_K <= k <= kSelMaxK), one query at a time: one pass that writes a score per
// row + a grid-wide radix select (select_kernels.cuh).  Fully asynchronous, nothing can overflow,
// works for any query (a query the int8 bound cannot cover sends every live row to the exact stage).
static int search_select_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                                uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                                cudaStream_t stream) {
    using u64 = unsigned long long;
    const uint64_t n = ix->n_rows;
    const bool i8 = ix->i8_ok && !ix->d_slab_f32 && env_int("FSGPU_MMA_I8", 1) != 0 && env_int("FSGPU_SELECT_I8", 1) != 0;
    CUDA_TRY(ix->ws_sel_state.reserve(2 * sizeof(SelState)));
    CUDA_TRY(ix->ws_sort_a.reserve(n * 8));
    CUDA_TRY(ix->ws_sel_pos2.reserve((size_t)kSelMaxK * 4));
    if (i8) {
        CUDA_TRY(ix->ws_approx.reserve(n * 4));
        CUDA_TRY(ix->ws_sel_pos.reserve(n * 4));
        CUDA_TRY(ix->ws_qhat.reserve((size_t)ix->dim));
        CUDA_TRY(ix->ws_margin.reserve(4));
        CUDA_TRY(ix->ws_qscale.reserve(4));
        CUDA_TRY(ix->ws_redo.reserve(4));
    }
    SelState* st0 = ix->ws_sel_state.as<SelState>();
    SelState* st1 = st0 + 1;
    uint32_t* n1 = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(st0) + offsetof(SelState, n_out));
    uint32_t* n2 = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(st1) + offsetof(SelState, n_out));
    u64* keys = ix->ws_sort_a.as<u64>();
    const uint8_t* tomb = ix->d_excl ? ix->d_excl : ix->d_tomb;
    const int wide_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 2047) / 2048, (uint64_t)ix->num_sms * 8));
    const int small_grid = ix->num_sms;  // lists whose length only the device knows (a few k rows)
    CUDA_TRY(cudaFuncSetAttribute(sel_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelMaxK * 8)));
    for (uint32_t b = 0; b < batch; ++b) {
        const float* q = d_queries + (size_t)b * ix->dim;
        CUDA_TRY(cudaMemsetAsync(st0, 0, 2 * sizeof(SelState), stream));
        const uint32_t* n_exact = nullptr;  // length of `keys` (nullptr = n)
        if (i8) {
            mma_prep_queries_i8_kernel<<<1, 128, 0, stream>>>(q, 1, ix->dim, ix->max_row_norm, ix->i8_max_ex, ix->i8_sx,
                                                              ix->ws_qhat.as<int8_t>(), ix->ws_margin.as<float>(),
                                                              ix->ws_qscale.as<float>(), ix->ws_redo.as<uint32_t>());
            CUDA_TRY(cudaGetLastError());
            std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
            if (ix->profiling) {
                if (!ix->ev_free.empty()) {
                    ev = ix->ev_free.back();
                    ix->ev_free.pop_back();
                } else {
                    CUDA_TRY(cudaEventCreate(&ev.first));
                    CUDA_TRY(cudaEventCreate(&ev.second));
                }
                CUDA_TRY(cudaEventRecord(ev.first, stream));
            }
            const int scan_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 63) / 64, (uint64_t)ix->num_sms * 8));
            scan_i8_all_kernel<<<scan_grid, 256, 0, stream>>>(ix->d_slab_i8.as<int8_t>(), tomb, ix->ws_qhat.as<int8_t>(),
                                                              ix->ws_qscale.as<float>(), n, ix->dim,
                                                              ix->ws_approx.as<uint32_t>());
            CUDA_TRY(cudaGetLastError());
            if (ix->profiling) {
                CUDA_TRY(cudaEventRecord(ev.second, stream));
                ix->ev_pending.push_back(ev);
            }
            ix->prof.scan_launches += 1;
            ix->prof.i8_launches += 1;
            ix->prof.scan_bytes += n * ix->dim;
            for (int p = 0; p < SelTraits<uint32_t>::kPasses; ++p)
                sel_hist_kernel<uint32_t><<<wide_grid, 256, 0, stream>>>(ix->ws_approx.as<uint32_t>(), n, nullptr, st0, p, k);
            sel_compact_kernel<uint32_t><<<wide_grid, 256, 0, stream>>>(ix->ws_approx.as<uint32_t>(), n, nullptr, st0,
                                                                        ix->ws_margin.as<float>(), ix->ws_redo.as<uint32_t>(),
                                                                        ix->ws_sel_pos.as<uint32_t>(), (uint32_t)n, n1);
            gather_list_keys_kernel<<<ix->num_sms * 4, 256, query_smem_bytes(ix->dim), stream>>>(
                ix->d_slab, ix->row_base, ix->dim, q, ix->ws_sel_pos.as<uint32_t>(), n1, ix->reduce_order, ix->tail_fma, keys);
            CUDA_TRY(cudaGetLastError());
            n_exact = n1;
            ix->prof.other_launches += 6;
        } else {
            const int grid = (int)std::min<uint64_t>((n + kScanWarps - 1) / kScanWarps, (uint64_t)ix->num_sms * 8);
            score_all_kernel<<<grid, kScanThreads, query_smem_bytes(ix->dim), stream>>>(ix->slab_any(), ix->is_f32(), tomb, n, ix->row_base, ix->dim, q,
                                                                                 ix->reduce_order, ix->tail_fma,
                                                                                 reinterpret_cast<uint64_t*>(keys));
            CUDA_TRY(cudaGetLastError());
            ix->prof.scan_launches += 1;
            ix->prof.scan_bytes += n * ix->dim * 2ull;
        }
        const int g2 = n_exact ? small_grid : wide_grid;
        for (int p = 0; p < SelTraits<u64>::kPasses; ++p)
            sel_hist_kernel<u64><<<g2, 256, 0, stream>>>(keys, n, n_exact, st1, p, k);
        sel_compact_kernel<u64><<<g2, 256, 0, stream>>>(keys, n, n_exact, st1, nullptr, nullptr, ix->ws_sel_pos2.as<uint32_t>(),
                                                        kSelMaxK, n2);
        sel_emit_kernel<<<1, 1024, kSelMaxK * 8, stream>>>(keys, ix->ws_sel_pos2.as<uint32_t>(), n2, k, ix->slab_any(), ix->is_f32(), q, n,
                                                           ix->row_base, ix->dim, ix->reduce_order, ix->tail_fma,
                                                           d_out_keys ? d_out_keys + (size_t)b * k : nullptr,
                                                           d_out_hits ? d_out_hits + (size_t)b * k : nullptr,
                                                           d_out_counts ? d_out_counts + b : nullptr, ix->d_error);
        CUDA_TRY(cudaGetLastError());
        ix->prof.other_launches += 7;
        ix->prof.merge_launches += 1;
    }
    return FSGPU_OK;
}

static int search_main_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                              uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                              cudaStream_t stream) {
    if (batch == 0) return FSGPU_OK;
    if (ix->d_slab_f32) {  // f32-quantised slab: exact score of every row, then radix select (or the sort arm)
        if (k >= 1 && k <= kSelMaxK && ix->n_rows > 0 && ix->n_rows <= 0xFFFFFFF0ull)
            return search_select_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
        return search_exact_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
    }
    if (ix->use_i8_single) return search_i8_single_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
    const int min_batch = env_int("FSGPU_MMA_MIN_BATCH", 3);
    const bool mma = ix->mma_ok && min_batch > 0 && batch >= (uint32_t)min_batch && k >= 1 && k <= kMmaMaxK &&
                     ix->n_rows > 0;
    if (mma) return search_mma_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
    // large k off the tensor-core path (few queries, or k past its ceiling): radix select; it needs 16-byte
    // rows for the int8 codes only, any dim otherwise
    const uint32_t select_min_k = (uint32_t)std::max(1, env_int("FSGPU_SELECT_MIN_K", 129));
    if (k >= select_min_k && k <= kSelMaxK && ix->n_rows > 0 && ix->n_rows <= 0xFFFFFFF0ull)
        return search_select_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
    return search_exact_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
}

// Merge of ONE contiguous list per query (main top-k + WAL keys, or the keys of a selective gather) into
// the best k_out.  The shared-memory merge holds cand_capacity(k_out) keys; past that (k_out > 8192:
// `limit >= record_count` searches on an index with WAL rows, search.rs:449-493) the list is sorted
// whole, like the score-all arm.
static int merge_single_list_locked(const fsgpu_index* ix, const MergeArgs& m, uint32_t batch, cudaStream_t stream) {
    if ((size_t)m.cap * 8 + 16 <= 200 * 1024) return launch_merge(m, batch, stream);
    const uint64_t n = m.k_in;
    CUDA_TRY(ix->ws_sort_b.reserve(n * 8));
    size_t cub_bytes = 0;
    CUDA_TRY(cub::DeviceRadixSort::SortKeysDescending(nullptr, cub_bytes, m.keys, ix->ws_sort_b.as<uint64_t>(), n, 0, 64, stream));
    CUDA_TRY(ix->ws_cub.reserve(cub_bytes));
    const uint32_t k_eff = (uint32_t)std::min<uint64_t>(m.k_out, n);
    for (uint32_t b = 0; b < batch; ++b) {
        CUDA_TRY(cub::DeviceRadixSort::SortKeysDescending(ix->ws_cub.p, cub_bytes, m.keys + (size_t)b * m.query_stride,
                                                         ix->ws_sort_b.as<uint64_t>(), n, 0, 64, stream));
        emit_sorted_prefix_kernel<<<std::max(1u, std::min(1024u, (m.k_out + kScanWarps - 1) / kScanWarps)), kScanThreads, 0,
                                    stream>>>(ix->ws_sort_b.as<uint64_t>(), k_eff, m.k_out, m.slab, m.slab_is_f32,
                                              m.queries + (size_t)b * m.dim, m.n_rows, m.row_base, m.dim, m.reduce_order,
                                              m.tail_fma, m.out_keys ? m.out_keys + (size_t)b * m.k_out : nullptr,
                                              m.out_hits ? m.out_hits + (size_t)b * m.k_out : nullptr,
                                              m.out_counts ? m.out_counts + b : nullptr);
        CUDA_TRY(cudaGetLastError());
        ix->prof.other_launches += 4;
    }
    return FSGPU_OK;
}

// The selective arm of a filtered search: score only the listed rows (scan_gather_positions,
// search.rs:1178-1255) and keep the best k.  Same outputs as search_main_locked.
static int search_gather_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                                uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                                cudaStream_t stream) {
    const uint32_t cap = ix->gather_cap;
    CUDA_TRY(ix->ws_gather_keys.reserve((size_t)batch * cap * 8));
    const dim3 grid((cap + kScanWarps - 1) / kScanWarps, batch);
    gather_keys_kernel<<<grid, kScanThreads, 0, stream>>>(ix->d_slab, ix->row_base, ix->dim, d_queries,
                                                          ix->d_gather_pos, ix->d_gather_count, cap,
                                                          ix->reduce_order, ix->tail_fma,
                                                          ix->ws_gather_keys.as<uint64_t>());
    CUDA_TRY(cudaGetLastError());
    ix->prof.other_launches += 1;
    ix->prof.merge_launches += 1;
    MergeArgs m{};
    m.keys = ix->ws_gather_keys.as<uint64_t>();
    m.list_stride = 0;
    m.query_stride = cap;
    m.n_lists = 1;
    m.k_in = cap;
    m.k_in_used = ix->d_gather_count;  // skip the empty tail of the list
    m.k_out = k;
    m.cap = cand_capacity(k);
    m.out_keys = d_out_keys;
    m.out_hits = d_out_hits;
    m.out_counts = d_out_counts;
    m.slab = ix->slab_any();
        m.slab_is_f32 = ix->is_f32();
    m.queries = d_queries;
    m.n_rows = ix->n_rows;
    m.row_base = ix->row_base;
    m.dim = ix->dim;
    m.reduce_order = ix->reduce_order;
    m.tail_fma = ix->tail_fma;
    m.error_flag = ix->d_error;
    return merge_single_list_locked(ix, m, batch, stream);
}

// Main slab + resident WAL rows (VectorIndex::search_top_k_internal, search.rs:476-493): the slab's
// top-k keys and one key per WAL row form one list per query, reduced by the same merge kernel.
static int search_device_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                                uint64_t* d_out_keys, fsgpu_hit* d_out_hits, uint32_t* d_out_counts,
                                cudaStream_t stream) {
    if (batch == 0) return FSGPU_OK;
    auto main_search = (ix->d_gather_pos && k > 0 && ix->n_rows > 0) ? search_gather_locked : search_main_locked;
    if (ix->n_wal == 0 || k == 0)
        return main_search(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, stream);
    const uint64_t* main_keys = nullptr;
    if (ix->n_rows > 0) {
        CUDA_TRY(ix->ws_wal_main.reserve((size_t)batch * k * 8));
        CUDA_TRY(cudaMemsetAsync(ix->ws_wal_main.p, 0, (size_t)batch * k * 8, stream));
        int rc = main_search(ix, d_queries, batch, k, ix->ws_wal_main.as<uint64_t>(), nullptr, nullptr, stream);
        if (rc) return rc;
        main_keys = ix->ws_wal_main.as<uint64_t>();
    }
    const size_t stride = (size_t)k + ix->n_wal;
    CUDA_TRY(ix->ws_wal_keys.reserve((size_t)batch * stride * 8));
    const dim3 grid((ix->n_wal + kScanWarps - 1) / kScanWarps, batch);
    wal_keys_kernel<<<grid, kScanThreads, 0, stream>>>(ix->d_wal.as<float>(), ix->n_wal, ix->wal_base, ix->dim,
                                                       d_queries, ix->d_wal_allow, ix->wal_allow_bit0_is_zero ? 0 : ix->n_rows, main_keys, k,
                                                       ix->reduce_order, ix->ws_wal_keys.as<uint64_t>());
    CUDA_TRY(cudaGetLastError());
    ix->prof.other_launches += 1;
    MergeArgs m{};
    m.keys = ix->ws_wal_keys.as<uint64_t>();
    m.list_stride = 0;
    m.query_stride = stride;
    m.n_lists = 1;
    m.k_in = (uint32_t)stride;
    m.k_out = k;
    m.cap = cand_capacity(k);
    m.out_keys = d_out_keys;
    m.out_hits = d_out_hits;
    m.out_counts = d_out_counts;
    m.slab = ix->slab_any();
        m.slab_is_f32 = ix->is_f32();  // raw NaN scores of main rows (WAL rows are always finite)
    m.queries = d_queries;
    m.n_rows = ix->n_rows;
    m.row_base = ix->row_base;
    m.dim = ix->dim;
    m.reduce_order = ix->reduce_order;
    m.tail_fma = ix->tail_fma;
    m.error_flag = ix->d_error;
    ix->prof.merge_launches += 1;
    return merge_single_list_locked(ix, m, batch, stream);
}

// ─── index creation ─────────────────────────────────────────────────────────────────────────
static int index_alloc_common(fsgpu_index* ix, const fsgpu_index_options* o, uint64_t n_rows,
                              uint32_t dim) {
    if (dim == 0) return fail(FSGPU_ERR_INVALID_CONFIG, "dimension must be non-zero");
    if (o->reduce_order < 0 || o->reduce_order > 4)
        return fail(FSGPU_ERR_INVALID_CONFIG, "reduce_order %d out of range", o->reduce_order);
    if (o->row_base + n_rows > 0xFFFFFFFFull)
        return fail(FSGPU_ERR_INVALID_CONFIG, "row_base + n_rows exceeds the u32 range of VectorHit.index");
    int ndev = 0;
    int rc = fsgpu_device_count(&ndev);
    if (rc) return rc;
    if (o->device < 0 || o->device >= ndev)
        return fail(FSGPU_ERR_INVALID_CONFIG, "device %d not present (%d devices)", o->device, ndev);
    ix->device = o->device;
    ix->n_rows = n_rows;
    ix->dim = dim;
    ix->row_base = o->row_base;
    ix->reduce_order = o->reduce_order;
    ix->tail_fma = o->tail_fma ? 1 : 0;
    ix->want_i8 = o->int8_codes != 0;
    CUDA_TRY(cudaSetDevice(ix->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, ix->device));
    if (prop.major < 10)
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: device %d is sm_%d%d; this library is built for sm_100a only",
                    ix->device, prop.major, prop.minor);
    ix->num_sms = prop.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaMalloc(&ix->d_error, 16));
    CUDA_TRY(cudaMemset(ix->d_error, 0, 16));
    CUDA_TRY(cudaHostAlloc(&ix->h_flags, 16, cudaHostAllocDefault));
    memset(ix->h_flags, 0, 16);
    CUDA_TRY(cudaEventCreateWithFlags(&ix->ev_last, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&ix->ev_flags, cudaEventDisableTiming));
    return FSGPU_OK;
}

static int upload_tombstones(fsgpu_index* ix, const uint8_t* bitmap) {
    if (ix->d_tomb) {
        cudaFree(ix->d_tomb);
        ix->d_tomb = nullptr;
    }
    if (!bitmap || ix->n_rows == 0) return FSGPU_OK;
    const size_t bytes = (ix->n_rows + 7) / 8;
    CUDA_TRY(cudaMalloc(&ix->d_tomb, bytes));
    CUDA_TRY(h2d_complete(ix->d_tomb, bitmap, bytes));
    return FSGPU_OK;
}

extern "C" void fsgpu_index_destroy(fsgpu_index* ix) {
    if (!ix) return;
    {
        DeviceGuard g(ix->device);
        if (ix->stream) cudaStreamSynchronize(ix->stream);
        if (ix->owns_slab && ix->d_slab) cudaFree(ix->d_slab);
        if (ix->d_slab_f32) cudaFree(ix->d_slab_f32);
        if (ix->d_tomb) cudaFree(ix->d_tomb);
        if (ix->d_error) cudaFree(ix->d_error);
        if (ix->h_flags) cudaFreeHost(ix->h_flags);
        if (ix->ev_last) cudaEventDestroy(ix->ev_last);
        if (ix->ev_flags) cudaEventDestroy(ix->ev_flags);
        for (auto* v : {&ix->ev_pending, &ix->ev_free})
            for (auto& ev : *v) {
                cudaEventDestroy(ev.first);
                cudaEventDestroy(ev.second);
            }
        for (DevBuf* b : {&ix->ws_partial, &ix->ws_queries, &ix->ws_keys, &ix->ws_hits, &ix->ws_counts,
                          &ix->ws_sort_a, &ix->ws_sort_b, &ix->ws_cub, &ix->ws_rows, &ix->ws_scores,
                          &ix->ws_present, &ix->ws_excl, &ix->ws_allow, &ix->ws_progress, &ix->ws_qhat, &ix->ws_margin, &ix->ws_gate, &ix->ws_redo, &ix->ws_cand,
                          &ix->ws_cand_count, &ix->d_wal, &ix->ws_wal_main, &ix->ws_wal_keys, &ix->d_hashes,
                          &ix->ws_allowed, &ix->ws_gather_pos, &ix->ws_gather_count, &ix->ws_gather_keys,
                          &ix->d_slab_i8, &ix->ws_qscale, &ix->ws_approx, &ix->ws_i8_top, &ix->ws_i8_cnt,
                          &ix->ws_sel_pos, &ix->ws_sel_pos2, &ix->ws_sel_state, &ix->ws_redo_slots, &ix->ws_redo_partial,
                          &ix->d_tp_codes8, &ix->d_tp_codes4})
            b->release();
        if (ix->stream) cudaStreamDestroy(ix->stream);
    }
    delete ix;
}

extern "C" int fsgpu_index_create_f16(const uint16_t* slab, uint64_t n_rows, uint32_t dim,
                                      const uint8_t* tombstones, const fsgpu_index_options* opts,
                                      fsgpu_index** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    fsgpu_index_options o;
    if (opts) o = *opts; else fsgpu_index_options_default(&o);
    if (n_rows && !slab) return fail(FSGPU_ERR_INVALID_CONFIG, "slab is NULL");
    DeviceGuard g(o.device);
    fsgpu_index* ix = new fsgpu_index();
    int rc = index_alloc_common(ix, &o, n_rows, dim);
    if (rc) { fsgpu_index_destroy(ix); return rc; }
    const size_t bytes = (size_t)n_rows * dim * 2;
    cudaError_t e = cudaSuccess;
    if (o.slab_is_device) {
        ix->d_slab = const_cast<uint16_t*>(slab);
        ix->owns_slab = false;
        if ((reinterpret_cast<uintptr_t>(slab) & 15u) != 0) {
            fsgpu_index_destroy(ix);
            return fail(FSGPU_ERR_INVALID_CONFIG, "device slab must be 16-byte aligned");
        }
    } else if (bytes) {
        e = cudaMalloc(&ix->d_slab, bytes);
        if (e == cudaSuccess) e = h2d_complete(ix->d_slab, slab, bytes);
        ix->owns_slab = true;
    }
    if (e != cudaSuccess) {
        fsgpu_index_destroy(ix);
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: slab upload failed: %s", cudaGetErrorString(e));
    }
    rc = upload_tombstones(ix, tombstones);
    if (!rc) rc = index_finish_setup(ix);
    if (rc) { fsgpu_index_destroy(ix); return rc; }
    *out = ix;
    return FSGPU_OK;
}

extern "C" int fsgpu_index_create_f32(const float* rows, uint64_t n_rows, uint32_t dim,
                                      const uint8_t* tombstones, const fsgpu_index_options* opts,
                                      fsgpu_index** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    fsgpu_index_options o;
    if (opts) o = *opts; else fsgpu_index_options_default(&o);
    if (n_rows && !rows) return fail(FSGPU_ERR_INVALID_CONFIG, "rows is NULL");
    DeviceGuard g(o.device);
    fsgpu_index* ix = new fsgpu_index();
    int rc = index_alloc_common(ix, &o, n_rows, dim);
    if (rc) { fsgpu_index_destroy(ix); return rc; }
    const uint64_t count = n_rows * dim;
    if (count) {
        const float* d_src = rows;
        float* staged = nullptr;
        cudaError_t e = cudaMalloc(&ix->d_slab, count * 2);
        ix->owns_slab = true;
        if (e == cudaSuccess && !o.slab_is_device) {
            e = cudaMalloc(&staged, count * 4);
            if (e == cudaSuccess) e = h2d_complete(staged, rows, count * 4);
            d_src = staged;
        }
        if (e == cudaSuccess) {
            encode_f16_kernel<<<ix->num_sms * 8, 256, 0, ix->stream>>>(d_src, count, ix->d_slab);
            e = cudaGetLastError();
            if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
        }
        if (staged) cudaFree(staged);
        if (e != cudaSuccess) {
            fsgpu_index_destroy(ix);
            return fail(FSGPU_ERR_SUBSYSTEM, "gpu: f32->f16 encode failed: %s", cudaGetErrorString(e));
        }
    }
    rc = upload_tombstones(ix, tombstones);
    if (!rc) rc = index_finish_setup(ix);
    if (rc) { fsgpu_index_destroy(ix); return rc; }
    *out = ix;
    return FSGPU_OK;
}

extern "C" uint64_t fsgpu_index_rows(const fsgpu_index* ix) { return ix ? ix->n_rows : 0; }
extern "C" uint32_t fsgpu_index_dim(const fsgpu_index* ix) { return ix ? ix->dim : 0; }
extern "C" uint64_t fsgpu_index_row_base(const fsgpu_index* ix) { return ix ? ix->row_base : 0; }
extern "C" int fsgpu_index_device(const fsgpu_index* ix) { return ix ? ix->device : -1; }
extern "C" const void* fsgpu_index_device_slab(const fsgpu_index* ix) { return ix ? ix->d_slab : nullptr; }

extern "C" int fsgpu_index_set_doc_ids(fsgpu_index* ix, const uint8_t* bytes, const uint64_t* offsets) {
    if (!ix || !offsets) return fail(FSGPU_ERR_INVALID_CONFIG, "index or offsets is NULL");
    std::lock_guard<std::mutex> lock(ix->mu);
    ix->doc_off.assign(offsets, offsets + ix->n_rows + 1);
    const uint64_t total = ix->doc_off.back();
    if (total && !bytes) return fail(FSGPU_ERR_INVALID_CONFIG, "doc-id bytes is NULL");
    ix->doc_bytes.assign(bytes, bytes + total);
    return FSGPU_OK;
}

extern "C" int fsgpu_index_doc_id(const fsgpu_index* ix, uint64_t global_row, const uint8_t** out_ptr,
                                  uint32_t* out_len) {
    if (!ix || !out_ptr || !out_len) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    if (ix->doc_off.empty()) return fail(FSGPU_ERR_INVALID_CONFIG, "index has no doc-id table");
    if (global_row < ix->row_base || global_row - ix->row_base >= ix->n_rows)
        return fail(FSGPU_ERR_INVALID_CONFIG, "row %llu is not in this shard", (unsigned long long)global_row);
    const uint64_t r = global_row - ix->row_base;
    *out_ptr = ix->doc_bytes.data() + ix->doc_off[r];
    *out_len = (uint32_t)(ix->doc_off[r + 1] - ix->doc_off[r]);
    return FSGPU_OK;
}

extern "C" int fsgpu_index_read_rows_f16(const fsgpu_index* ix, uint64_t row_start, uint64_t n,
                                         uint16_t* out_bits) {
    if (!ix || (n && !out_bits)) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    if (row_start > ix->n_rows || n > ix->n_rows - row_start)
        return fail(FSGPU_ERR_INVALID_CONFIG, "row range outside the index");
    if (n == 0) return FSGPU_OK;
    if (ix->d_slab_f32) return fail(FSGPU_ERR_INVALID_CONFIG, "index holds an f32-quantised slab: no f16 rows to read");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    CUDA_TRY(cudaMemcpy(out_bits, ix->d_slab + row_start * ix->dim, (size_t)n * ix->dim * 2,
                        cudaMemcpyDeviceToHost));
    return FSGPU_OK;
}

extern "C" int fsgpu_index_set_tombstones(fsgpu_index* ix, const uint8_t* bitmap) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    return upload_tombstones(ix, bitmap);
}

extern "C" int fsgpu_index_int8_ready(const fsgpu_index* ix) { return ix && ix->i8_ok ? 1 : 0; }

extern "C" int fsgpu_index_read_codes_i8(const fsgpu_index* ix, uint64_t row_start, uint64_t n, int8_t* out_codes,
                                         float* out_scale) {
    if (!ix || !out_codes) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    if (!ix->i8_ok) return fail(FSGPU_ERR_INVALID_CONFIG, "index holds no int8 codes (FSGPU_MMA_I8=1 at creation, dim %% 128 == 0)");
    if (row_start > ix->n_rows || n > ix->n_rows - row_start)
        return fail(FSGPU_ERR_INVALID_CONFIG, "row range [%llu, +%llu) outside the index",
                    (unsigned long long)row_start, (unsigned long long)n);
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    CUDA_TRY(cudaMemcpy(out_codes, ix->d_slab_i8.as<int8_t>() + row_start * ix->dim, (size_t)n * ix->dim,
                        cudaMemcpyDeviceToHost));
    if (out_scale) *out_scale = ix->i8_sx;
    return FSGPU_OK;
}

extern "C" int fsgpu_index_read_tombstones(const fsgpu_index* ix, uint8_t* out_bitmap) {
    if (!ix || !out_bitmap) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    const size_t bytes = (ix->n_rows + 7) / 8;
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    if (!ix->d_tomb) {
        memset(out_bitmap, 0, bytes);
        return FSGPU_OK;
    }
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    CUDA_TRY(cudaMemcpy(out_bitmap, ix->d_tomb, bytes, cudaMemcpyDeviceToHost));
    return FSGPU_OK;
}

extern "C" int fsgpu_index_zero_signal_state(const fsgpu_index* ix, uint64_t* out_state) {
    if (!ix || !out_state) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    unsigned long long counts[2] = {0, 0};
    if (ix->n_rows) {
        int rc = begin_call_locked(ix, ix->stream, true);
        if (rc) return rc;
        unsigned long long* d_counts = nullptr;
        CUDA_TRY(cudaMalloc(&d_counts, 16));
        cudaError_t e = cudaMemsetAsync(d_counts, 0, 16, ix->stream);
        if (e == cudaSuccess) {
            const int grid = (int)std::min<uint64_t>((ix->n_rows + 255) / 256, (uint64_t)ix->num_sms * 8);
            census_kernel<<<grid, 256, 0, ix->stream>>>(ix->slab_any(), ix->is_f32(), ix->n_rows, ix->dim, ix->d_tomb, d_counts);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d_counts, 16, cudaMemcpyDeviceToHost, ix->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
        cudaFree(d_counts);
        if (e != cudaSuccess) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: census failed: %s", cudaGetErrorString(e));
        rc = end_call_locked(ix, ix->stream);
        if (rc) return rc;
    }
    out_state[0] = ix->n_rows;               // record_count
    out_state[1] = ix->n_rows - counts[0];   // live_count
    out_state[2] = counts[0];                // tombstone_count
    out_state[3] = ix->n_wal;                // wal_count
    out_state[4] = counts[1];                // usable_vector_count
    return FSGPU_OK;
}

extern "C" int fsgpu_index_set_wal(fsgpu_index* ix, const float* embeddings, uint32_t n_wal, uint64_t virtual_base) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (n_wal && !embeddings) return fail(FSGPU_ERR_INVALID_CONFIG, "embeddings is NULL");
    if (virtual_base + n_wal > 0xFFFFFFFFull)  // resolve_wal_hit (search.rs:1590-1594)
        return fail(FSGPU_ERR_INVALID_CONFIG, "WAL entry index exceeds u32 range");
    if (n_wal && virtual_base < ix->row_base + ix->n_rows)
        return fail(FSGPU_ERR_INVALID_CONFIG, "WAL rows must be numbered after the slab's rows (virtual_base %llu < %llu)",
                    (unsigned long long)virtual_base, (unsigned long long)(ix->row_base + ix->n_rows));
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    if (n_wal) {
        CUDA_TRY(ix->d_wal.reserve((size_t)n_wal * ix->dim * 4));
        CUDA_TRY(h2d_complete(ix->d_wal.p, embeddings, (size_t)n_wal * ix->dim * 4));
    }
    ix->n_wal = n_wal;
    ix->wal_base = virtual_base;
    return FSGPU_OK;
}

extern "C" uint32_t fsgpu_index_wal_rows(const fsgpu_index* ix) { return ix ? ix->n_wal : 0; }

extern "C" int fsgpu_index_profile_enable(fsgpu_index* ix, int on) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    std::lock_guard<std::mutex> lock(ix->mu);
    ix->profiling = on != 0;
    return FSGPU_OK;
}

extern "C" int fsgpu_index_profile_read(fsgpu_index* ix, fsgpu_profile* out, int reset) {
    if (!ix || !out) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    consume_flags_locked(ix, true);
    for (auto& ev : ix->ev_pending) {
        CUDA_TRY(cudaEventSynchronize(ev.second));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
        ix->prof.scan_ms += ms;
        ix->ev_free.push_back(ev);
    }
    ix->ev_pending.clear();
    *out = ix->prof;
    if (reset) ix->prof = fsgpu_profile{};
    return FSGPU_OK;
}

// ─── search ─────────────────────────────────────────────────────────────────────────────────
extern "C" int fsgpu_search_top_k_device(const fsgpu_index* ix, const float* d_queries, uint32_t batch,
                                         uint32_t k, uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                                         uint32_t* d_out_counts, void* stream) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (batch && !d_queries) return fail(FSGPU_ERR_INVALID_CONFIG, "queries is NULL");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ix->stream;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int rc = begin_call_locked(ix, s, stream == nullptr);
        if (rc) return rc;
        if (k && d_out_keys) CUDA_TRY(cudaMemsetAsync(d_out_keys, 0, (size_t)batch * k * 8, s));
        ix->force_f16 = attempt == 1;
        rc = search_device_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, s);
        ix->force_f16 = false;
        if (!rc) rc = end_call_locked(ix, s);
        if (rc) return rc;
        if (stream) return FSGPU_OK;  // asynchronous: fsgpu_index_last_status reports once the stream is done
        bool retry = false;
        rc = finish_sync_call_locked(ix, s, &retry);
        if (rc || !retry) return rc;
    }
    return FSGPU_OK;
}

extern "C" int fsgpu_index_last_status(const fsgpu_index* ix, uint32_t* out_flags) {
    if (!ix || !out_flags) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    if (ix->flags_pending) CUDA_TRY(cudaEventSynchronize(ix->ev_flags));
    memcpy(out_flags, ix->h_flags, 16);
    consume_flags_locked(ix, true);
    if (out_flags[0]) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: top-k candidate buffer contract violated");
    return FSGPU_OK;
}

extern "C" int fsgpu_search_top_k_filtered(const fsgpu_index* ix, const float* queries, uint32_t batch, uint32_t k,
                                           uint32_t dim, const uint8_t* allow_bitmap, fsgpu_hit* out,
                                           uint32_t* out_counts);

extern "C" int fsgpu_search_top_k(const fsgpu_index* ix, const float* queries, uint32_t batch, uint32_t k,
                                  uint32_t dim, fsgpu_hit* out, uint32_t* out_counts) {
    return fsgpu_search_top_k_filtered(ix, queries, batch, k, dim, nullptr, out, out_counts);
}

// ─── the reference's quantised two-pass searches ────────────────────────────────────────────
// largest |element| of the slab as f16 magnitude bits, NaN ignored (f32::max ignores NaN: simd.rs:1842-1846)
__global__ void slab_max_abs_kernel(const uint16_t* __restrict__ slab, uint64_t n, uint32_t* __restrict__ out) {
    uint32_t m = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = slab[i] & 0x7FFFu;
        if (b <= 0x7C00u) m = max(m, b);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

static float f16_magnitude_to_f32(uint32_t hb) {
    const uint32_t exp = (hb >> 10) & 0x1F, man = hb & 0x3FF;
    if (exp == 31) return INFINITY;
    return exp == 0 ? std::ldexp((float)man, -24) : std::ldexp((float)(man | 0x400), (int)exp - 25);
}

extern "C" int fsgpu_search_top_k_two_pass(const fsgpu_index* ix, const float* query, uint32_t k, uint32_t candidate_multiplier,
                                           int bits, uint32_t dim, fsgpu_hit* out, uint32_t* out_count) {
    using u64 = unsigned long long;
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (bits != 8 && bits != 4) return fail(FSGPU_ERR_INVALID_CONFIG, "bits must be 8 or 4, got %d", bits);
    if (!out_count) return fail(FSGPU_ERR_INVALID_CONFIG, "out_count is NULL");
    // the reference's gate (search.rs:578-586, :882-890): k == 0, an empty index, resident WAL rows or a slab that is
    // not f16 take the exact search (which also reports the dimension mismatch); so does a k the select cannot hold
    if (k == 0 || ix->n_rows == 0 || ix->n_wal > 0 || ix->d_slab_f32 || k > kSelMaxK || ix->n_rows >= 0xFFFFFFFFull)
        return fsgpu_search_top_k(ix, query, 1, k, dim, out, out_count);
    if (dim != ix->dim) return fail(FSGPU_ERR_DIMENSION_MISMATCH, "expected %u, found %u", ix->dim, dim);
    if (!query || !out) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = ix->stream;
    int rc = begin_call_locked(ix, s, true);
    if (rc) return rc;
    const uint64_t n = ix->n_rows;
    // corpus-wide scale (once per index)
    if (ix->tp_max_abs < 0.0f) {
        CUDA_TRY(ix->ws_sel_state.reserve(2 * sizeof(SelState)));
        CUDA_TRY(cudaMemsetAsync(ix->ws_sel_state.p, 0, 4, s));
        slab_max_abs_kernel<<<ix->num_sms * 8, 256, 0, s>>>(ix->d_slab, n * dim, ix->ws_sel_state.as<uint32_t>());
        CUDA_TRY(cudaGetLastError());
        uint32_t hb = 0;
        CUDA_TRY(cudaMemcpyAsync(&hb, ix->ws_sel_state.p, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
        ix->tp_max_abs = f16_magnitude_to_f32(hb);
    }
    const float max_abs = ix->tp_max_abs;
    const uint32_t row_bytes = bits == 8 ? dim : (dim + 1u) / 2u;
    const uint8_t* codes = nullptr;
    if (bits == 8) {
        if (ix->i8_ok) {
            codes = ix->d_slab_i8.as<uint8_t>();  // the index's resident codes are the same quantiser's (simd.rs:1842-1859)
        } else {
            if (!ix->tp8_ready) {
                CUDA_TRY(ix->d_tp_codes8.reserve(n * dim));
                if (max_abs <= 0.0f) {
                    CUDA_TRY(cudaMemsetAsync(ix->d_tp_codes8.p, 0, n * dim, s));
                } else {
                    quantize_slab_i8_any_kernel<<<ix->num_sms * 8, 256, 0, s>>>(ix->d_slab, n * dim, 127.0f / max_abs,
                                                                               ix->d_tp_codes8.as<int8_t>());
                    CUDA_TRY(cudaGetLastError());
                }
                ix->tp8_ready = true;
            }
            codes = ix->d_tp_codes8.as<uint8_t>();
        }
    } else {
        if (!ix->tp4_ready) {
            CUDA_TRY(ix->d_tp_codes4.reserve(n * row_bytes));
            const float scale = max_abs > 1e-9f ? 7.0f / max_abs : 0.0f;  // simd.rs:2211
            pack_slab_4bit_kernel<<<ix->num_sms * 8, 256, 0, s>>>(ix->d_slab, n, dim, scale, ix->d_tp_codes4.as<uint8_t>());
            CUDA_TRY(cudaGetLastError());
            ix->tp4_ready = true;
        }
        codes = ix->d_tp_codes4.as<uint8_t>();
    }
    // the query's own codes (search.rs:1610-1655), on the host: `dim` values
    const uint32_t pad = (row_bytes + 15u) & ~15u;
    std::vector<int8_t> qc(2 * (size_t)pad, 0);
    float q_max = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) q_max = std::fmax(q_max, std::fabs(query[i]));  // f32::max: NaN is ignored
    if (bits == 8) {
        if (q_max > 0.0f) {
            const float scale = 127.0f / q_max;
            for (uint32_t i = 0; i < dim; ++i) {
                const float v = std::fmin(std::fmax(std::round(query[i] * scale), -127.0f), 127.0f);
                qc[i] = (int8_t)(std::isnan(v) ? 0 : (int)v);  // `NaN as i8` is 0 in Rust
            }
        }
    } else {
        const float scale = q_max > 1e-9f ? 7.0f / q_max : 0.0f;
        for (uint32_t i = 0; i < dim; ++i) {
            const float v = std::fmin(std::fmax(std::round(query[i] * scale), -7.0f), 7.0f);
            const int8_t c = (int8_t)(std::isnan(v) ? 0 : (int)v);
            (i % 2 == 0 ? qc[i / 2] : qc[pad + i / 2]) = c;
        }
    }
    CUDA_TRY(ix->ws_queries.reserve((size_t)dim * 4));
    CUDA_TRY(ix->ws_qhat.reserve(2 * (size_t)pad));
    CUDA_TRY(ix->ws_hits.reserve((size_t)k * sizeof(fsgpu_hit)));
    CUDA_TRY(ix->ws_counts.reserve(4));
    CUDA_TRY(ix->ws_sel_state.reserve(2 * sizeof(SelState)));
    CUDA_TRY(ix->ws_sort_a.reserve(n * 8));
    CUDA_TRY(ix->ws_sel_pos.reserve(n * 4));
    CUDA_TRY(ix->ws_sel_pos2.reserve((size_t)kSelMaxK * 4));
    CUDA_TRY(cudaMemcpyAsync(ix->ws_queries.p, query, (size_t)dim * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(ix->ws_qhat.p, qc.data(), qc.size(), cudaMemcpyHostToDevice, s));
    SelState* st0 = ix->ws_sel_state.as<SelState>();
    SelState* st1 = st0 + 1;
    uint32_t* n1 = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(st0) + offsetof(SelState, n_out));
    uint32_t* n2 = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(st1) + offsetof(SelState, n_out));
    CUDA_TRY(cudaMemsetAsync(st0, 0, 2 * sizeof(SelState), s));
    u64* keys = ix->ws_sort_a.as<u64>();
    const float* q = ix->ws_queries.as<float>();
    // candidate_count (search.rs:596-599): k * max(multiplier, 1), at most the record count, at least min(k, count)
    const uint64_t want = (uint64_t)k * std::max<uint32_t>(candidate_multiplier, 1u);
    const uint32_t cand = (uint32_t)std::max<uint64_t>(std::min<uint64_t>(want, n), std::min<uint64_t>(k, n));
    // pass 1: integer score of every live row as an order key, then the best `cand` of them
    const int scan_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 31) / 32, (uint64_t)ix->num_sms * 8));
    const int8_t* qa = ix->ws_qhat.as<int8_t>();
    if (bits == 8)
        two_pass_scan_kernel<8><<<scan_grid, 256, 2 * pad, s>>>(codes, row_bytes, ix->d_tomb, qa, qa + pad, n, ix->row_base, keys);
    else
        two_pass_scan_kernel<4><<<scan_grid, 256, 2 * pad, s>>>(codes, row_bytes, ix->d_tomb, qa, qa + pad, n, ix->row_base, keys);
    CUDA_TRY(cudaGetLastError());
    ix->prof.scan_launches += 1;
    ix->prof.scan_bytes += n * row_bytes;
    const int wide_grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 2047) / 2048, (uint64_t)ix->num_sms * 8));
    for (int p = 0; p < SelTraits<u64>::kPasses; ++p) sel_hist_kernel<u64><<<wide_grid, 256, 0, s>>>(keys, n, nullptr, st0, p, cand);
    sel_compact_kernel<u64><<<wide_grid, 256, 0, s>>>(keys, n, nullptr, st0, nullptr, nullptr, ix->ws_sel_pos.as<uint32_t>(),
                                                      (uint32_t)n, n1);
    // pass 2: exact f16 re-score of exactly those rows (search.rs:629-640), top k by the reference's total order
    gather_list_keys_kernel<<<ix->num_sms * 4, 256, query_smem_bytes(dim), s>>>(ix->d_slab, ix->row_base, dim, q, ix->ws_sel_pos.as<uint32_t>(),
                                                                         n1, ix->reduce_order, ix->tail_fma, keys);
    CUDA_TRY(cudaGetLastError());
    for (int p = 0; p < SelTraits<u64>::kPasses; ++p) sel_hist_kernel<u64><<<ix->num_sms, 256, 0, s>>>(keys, n, n1, st1, p, k);
    sel_compact_kernel<u64><<<ix->num_sms, 256, 0, s>>>(keys, n, n1, st1, nullptr, nullptr, ix->ws_sel_pos2.as<uint32_t>(), kSelMaxK, n2);
    CUDA_TRY(cudaFuncSetAttribute(sel_emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSelMaxK * 8)));
    sel_emit_kernel<<<1, 1024, kSelMaxK * 8, s>>>(keys, ix->ws_sel_pos2.as<uint32_t>(), n2, k, ix->slab_any(), ix->is_f32(), q, n, ix->row_base,
                                                  dim, ix->reduce_order, ix->tail_fma, nullptr, ix->ws_hits.as<fsgpu_hit>(),
                                                  ix->ws_counts.as<uint32_t>(), ix->d_error);
    CUDA_TRY(cudaGetLastError());
    ix->prof.other_launches += 17;
    CUDA_TRY(cudaMemcpyAsync(out, ix->ws_hits.p, (size_t)k * sizeof(fsgpu_hit), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(out_count, ix->ws_counts.p, 4, cudaMemcpyDeviceToHost, s));
    rc = end_call_locked(ix, s);
    if (rc) return rc;
    return finish_sync_call_locked(ix, s, nullptr);
}

// code slabs of the two-pass searches, for parity tests against the reference quantisers (host buffer of
// n_rows * dim bytes for bits = 8, n_rows * ceil(dim / 2) for bits = 4); builds them if needed
extern "C" int fsgpu_index_read_two_pass_codes(const fsgpu_index* ix, int bits, uint8_t* out) {
    if (!ix || !out || (bits != 8 && bits != 4)) return fail(FSGPU_ERR_INVALID_CONFIG, "bad argument");
    if (ix->n_rows == 0) return FSGPU_OK;
    if (ix->d_slab_f32) return fail(FSGPU_ERR_INVALID_CONFIG, "f32-quantised slab: no two-pass codes");
    std::vector<float> q(ix->dim, 0.0f);
    fsgpu_hit h;
    uint32_t c = 0;
    if (ix->n_wal == 0) {
        int rc = fsgpu_search_top_k_two_pass(ix, q.data(), 1, 1, bits, ix->dim, &h, &c);  // builds the slab as a side effect
        if (rc) return rc;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    const size_t bytes = ix->n_rows * (bits == 8 ? (size_t)ix->dim : ((size_t)ix->dim + 1) / 2);
    const void* src = bits == 8 ? (ix->i8_ok ? ix->d_slab_i8.p : (ix->tp8_ready ? ix->d_tp_codes8.p : nullptr))
                                : (ix->tp4_ready ? ix->d_tp_codes4.p : nullptr);
    if (!src) return fail(FSGPU_ERR_INVALID_CONFIG, "two-pass codes are not built (resident WAL rows?)");
    CUDA_TRY(cudaMemcpy(out, src, bytes, cudaMemcpyDeviceToHost));
    return FSGPU_OK;
}

// ─── filtered search ────────────────────────────────────────────────────────────────────────
__global__ void combine_exclusion_kernel(const uint8_t* __restrict__ tomb, const uint8_t* __restrict__ allow,
                                         size_t n_bytes, uint8_t* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_bytes; i += (size_t)gridDim.x * blockDim.x)
        out[i] = (uint8_t)((tomb ? tomb[i] : 0u) | (uint8_t)~allow[i]);
}

// Caller holds ix->mu.  `d_allow`: device bitmap, bit r%8 of byte r/8 set = local row r may be returned.
static int search_filtered_locked(const fsgpu_index* ix, const float* d_queries, uint32_t batch, uint32_t k,
                                  const uint8_t* d_allow, uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                                  uint32_t* d_out_counts, cudaStream_t s) {
    if (!d_allow || (ix->n_rows == 0 && ix->n_wal == 0))
        return search_device_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, s);
    if (ix->n_rows > 0) {
        const size_t n_bytes = (ix->n_rows + 7) / 8;
        CUDA_TRY(ix->ws_excl.reserve(n_bytes));
        combine_exclusion_kernel<<<(unsigned)std::min<size_t>((n_bytes + 255) / 256, 4096), 256, 0, s>>>(
            ix->d_tomb, d_allow, n_bytes, ix->ws_excl.as<uint8_t>());
        CUDA_TRY(cudaGetLastError());
        ix->prof.other_launches += 1;
        ix->d_excl = ix->ws_excl.as<uint8_t>();
    }
    ix->d_wal_allow = d_allow;  // WAL row w is bit n_rows + w of the same bitmap
    const int rc = search_device_locked(ix, d_queries, batch, k, d_out_keys, d_out_hits, d_out_counts, s);
    ix->d_excl = nullptr;
    ix->d_wal_allow = nullptr;
    return rc;
}

extern "C" int fsgpu_search_top_k_filtered_device(const fsgpu_index* ix, const float* d_queries, uint32_t batch,
                                                  uint32_t k, const uint8_t* d_allow_bitmap, uint64_t* d_out_keys,
                                                  fsgpu_hit* d_out_hits, uint32_t* d_out_counts, void* stream) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (batch && !d_queries) return fail(FSGPU_ERR_INVALID_CONFIG, "queries is NULL");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ix->stream;
    for (int attempt = 0; attempt < 2; ++attempt) {
        int rc = begin_call_locked(ix, s, stream == nullptr);
        if (rc) return rc;
        if (k && d_out_keys) CUDA_TRY(cudaMemsetAsync(d_out_keys, 0, (size_t)batch * k * 8, s));
        ix->force_f16 = attempt == 1;
        rc = search_filtered_locked(ix, d_queries, batch, k, d_allow_bitmap, d_out_keys, d_out_hits, d_out_counts, s);
        ix->force_f16 = false;
        if (!rc) rc = end_call_locked(ix, s);
        if (rc) return rc;
        if (stream) return FSGPU_OK;
        bool retry = false;
        rc = finish_sync_call_locked(ix, s, &retry);
        if (rc || !retry) return rc;
    }
    return FSGPU_OK;
}

extern "C" int fsgpu_search_top_k_filtered(const fsgpu_index* ix, const float* queries, uint32_t batch, uint32_t k,
                                           uint32_t dim, const uint8_t* allow_bitmap, fsgpu_hit* out,
                                           uint32_t* out_counts) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (dim != ix->dim) return fail(FSGPU_ERR_DIMENSION_MISMATCH, "expected %u, found %u", ix->dim, dim);
    if (batch == 0) return FSGPU_OK;
    if (!queries || !out_counts || (k && !out)) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    if (k == 0 || (ix->n_rows == 0 && ix->n_wal == 0)) {
        memset(out_counts, 0, (size_t)batch * 4);
        return FSGPU_OK;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = ix->stream;
    CUDA_TRY(ix->ws_queries.reserve((size_t)batch * dim * 4));
    CUDA_TRY(ix->ws_hits.reserve((size_t)batch * k * sizeof(fsgpu_hit)));
    CUDA_TRY(ix->ws_counts.reserve((size_t)batch * 4));
    const uint8_t* d_allow = nullptr;
    if (allow_bitmap) {
        const size_t n_bytes = (ix->n_rows + ix->n_wal + 7) / 8;
        CUDA_TRY(ix->ws_allow.reserve(n_bytes));
        d_allow = ix->ws_allow.as<uint8_t>();
    }
    // One or two finite queries on an index that holds int8 codes: int8 pass 1 + exact re-score
    // (half the bytes of the f16 scan); a position list that overflows re-runs the call on the f16 scan.
    const int min_batch = env_int("FSGPU_MMA_MIN_BATCH", 3);
    const uint32_t i8_max_batch = (uint32_t)std::max(0, env_int("FSGPU_I8_MAX_BATCH", 2));
    bool i8_single = ix->i8_ok && !ix->d_slab_f32 && env_int("FSGPU_MMA_I8", 1) != 0 && k <= 128 &&
                     (batch <= i8_max_batch || !(min_batch > 0 && batch >= (uint32_t)min_batch)) && batch <= 64 &&
                     k <= kFusedMaxK && ix->n_rows > 0;
    for (size_t i = 0; i8_single && i < (size_t)batch * dim; ++i) i8_single = std::isfinite(queries[i]);
    for (int attempt = 0; attempt < 2; ++attempt) {
        const bool i8_now = i8_single && attempt == 0;
        int rc = begin_call_locked(ix, s, true);
        if (rc) return rc;
        if (attempt == 0) {  // (the uploads above were enqueued before the wait: repeat them behind it)
            CUDA_TRY(cudaMemcpyAsync(ix->ws_queries.p, queries, (size_t)batch * dim * 4, cudaMemcpyHostToDevice, s));
            if (allow_bitmap)
                CUDA_TRY(cudaMemcpyAsync(ix->ws_allow.p, allow_bitmap, (ix->n_rows + ix->n_wal + 7) / 8, cudaMemcpyHostToDevice, s));
        }
        ix->use_i8_single = i8_now;
        ix->force_f16 = attempt == 1;
        rc = search_filtered_locked(ix, ix->ws_queries.as<float>(), batch, k, d_allow, nullptr,
                                    ix->ws_hits.as<fsgpu_hit>(), ix->ws_counts.as<uint32_t>(), s);
        ix->use_i8_single = false;
        ix->force_f16 = false;
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(out, ix->ws_hits.p, (size_t)batch * k * sizeof(fsgpu_hit), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(out_counts, ix->ws_counts.p, (size_t)batch * 4, cudaMemcpyDeviceToHost, s));
        rc = end_call_locked(ix, s);
        if (rc) return rc;
        const uint32_t* hf = ix->h_flags;
        bool retry = false;
        rc = finish_sync_call_locked(ix, s, &retry);
        if (rc) return rc;
        // int8 pass 1: word 1 = its position list overflowed; batched int8 form: bails -> f16 forms
        if (!((i8_now && hf[1]) || retry)) break;
    }
    return FSGPU_OK;
}

// ─── doc-id-hash filters ────────────────────────────────────────────────────────────────────
extern "C" int fsgpu_index_set_doc_hashes(fsgpu_index* ix, const uint64_t* hashes) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    CUDA_TRY(cudaStreamSynchronize(ix->stream));
    ix->has_hashes = false;
    if (!hashes || ix->n_rows == 0) return FSGPU_OK;
    CUDA_TRY(ix->d_hashes.reserve(ix->n_rows * 8));
    CUDA_TRY(h2d_complete(ix->d_hashes.p, hashes, ix->n_rows * 8));
    ix->has_hashes = true;
    return FSGPU_OK;
}

extern "C" int fsgpu_search_top_k_hashes(const fsgpu_index* ix, const float* queries, uint32_t batch, uint32_t k,
                                         uint32_t dim, const uint64_t* allowed_sorted, uint32_t n_allowed,
                                         const uint8_t* wal_allow_bitmap, fsgpu_hit* out, uint32_t* out_counts,
                                         int* out_used_gather) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (dim != ix->dim) return fail(FSGPU_ERR_DIMENSION_MISMATCH, "expected %u, found %u", ix->dim, dim);
    if (out_used_gather) *out_used_gather = 0;
    if (batch == 0) return FSGPU_OK;
    if (!queries || !out_counts || (k && !out)) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    if (n_allowed && !allowed_sorted) return fail(FSGPU_ERR_INVALID_CONFIG, "allowed_sorted is NULL");
    for (uint32_t i = 1; i < n_allowed; ++i)
        if (allowed_sorted[i - 1] >= allowed_sorted[i])
            return fail(FSGPU_ERR_INVALID_CONFIG, "allowed hashes must be strictly ascending");
    if (ix->n_rows && !ix->has_hashes)
        return fail(FSGPU_ERR_INVALID_CONFIG, "index has no doc-id hashes (fsgpu_index_set_doc_hashes)");
    if (k == 0 || (ix->n_rows == 0 && ix->n_wal == 0)) {
        memset(out_counts, 0, (size_t)batch * 4);
        return FSGPU_OK;
    }
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = ix->stream;
    CUDA_TRY(ix->ws_queries.reserve((size_t)batch * dim * 4));
    CUDA_TRY(ix->ws_hits.reserve((size_t)batch * k * sizeof(fsgpu_hit)));
    CUDA_TRY(ix->ws_counts.reserve((size_t)batch * 4));
    const size_t wal_bytes = (ix->n_wal + 7) / 8;
    if (ix->n_wal) CUDA_TRY(ix->ws_allow.reserve(wal_bytes));
    // try_gather_filtered (search.rs:1114-1131): the selective arm when allowed * 50 < record_count
    constexpr uint64_t kGatherSelectivityDivisor = 50;  // search.rs:33
    bool gather = ix->n_rows > 0 && (uint64_t)n_allowed * kGatherSelectivityDivisor < ix->n_rows && !ix->d_slab_f32;
    bool force_f16 = false;
    const uint32_t cap = std::max(64u, 2 * n_allowed);  // rows sharing a hash (collisions, duplicate doc ids)
    for (int attempt = 0; attempt < 3; ++attempt) {
        int rc0 = begin_call_locked(ix, s, true);
        if (rc0) return rc0;
        CUDA_TRY(cudaMemcpyAsync(ix->ws_queries.p, queries, (size_t)batch * dim * 4, cudaMemcpyHostToDevice, s));
        if (ix->n_wal) {
            // WAL rows: the host evaluated the filter on their doc ids (search.rs:1457-1465); bit w = WAL row w
            if (wal_allow_bitmap)
                CUDA_TRY(cudaMemcpyAsync(ix->ws_allow.p, wal_allow_bitmap, wal_bytes, cudaMemcpyHostToDevice, s));
            else
                CUDA_TRY(cudaMemsetAsync(ix->ws_allow.p, 0xFF, wal_bytes, s));
        }
        if (ix->n_rows > 0) {
            const size_t n_bytes = (ix->n_rows + 7) / 8;
            CUDA_TRY(ix->ws_excl.reserve(n_bytes));
            CUDA_TRY(ix->ws_allowed.reserve(std::max<size_t>(8, (size_t)n_allowed * 8)));
            if (n_allowed)
                CUDA_TRY(cudaMemcpyAsync(ix->ws_allowed.p, allowed_sorted, (size_t)n_allowed * 8, cudaMemcpyHostToDevice, s));
            if (gather) {
                CUDA_TRY(ix->ws_gather_pos.reserve((size_t)cap * 4));
                CUDA_TRY(ix->ws_gather_count.reserve(4));
                CUDA_TRY(cudaMemsetAsync(ix->ws_gather_count.p, 0, 4, s));
            }
            hash_filter_kernel<<<(unsigned)std::min<size_t>((n_bytes + 255) / 256, (size_t)ix->num_sms * 8), 256, 0, s>>>(
                ix->d_hashes.as<uint64_t>(), ix->n_rows, ix->d_tomb, ix->ws_allowed.as<uint64_t>(), n_allowed,
                ix->ws_excl.as<uint8_t>(), gather ? ix->ws_gather_pos.as<uint32_t>() : nullptr, cap,
                gather ? ix->ws_gather_count.as<uint32_t>() : nullptr);
            CUDA_TRY(cudaGetLastError());
            ix->prof.other_launches += 1;
            ix->d_excl = ix->ws_excl.as<uint8_t>();
            if (gather) {
                ix->d_gather_pos = ix->ws_gather_pos.as<uint32_t>();
                ix->d_gather_count = ix->ws_gather_count.as<uint32_t>();
                ix->gather_cap = cap;
            }
        }
        // WAL allow bits live at bit n_rows + w of d_wal_allow: point it so that bit 0 of ws_allow lands there
        ix->d_wal_allow = ix->n_wal ? ix->ws_allow.as<uint8_t>() : nullptr;
        ix->wal_allow_bit0_is_zero = true;
        ix->force_f16 = force_f16;
        int rc = search_device_locked(ix, ix->ws_queries.as<float>(), batch, k, nullptr,
                                      ix->ws_hits.as<fsgpu_hit>(), ix->ws_counts.as<uint32_t>(), s);
        ix->force_f16 = false;
        ix->d_excl = nullptr;
        ix->d_wal_allow = nullptr;
        ix->wal_allow_bit0_is_zero = false;
        ix->d_gather_pos = nullptr;
        ix->d_gather_count = nullptr;
        ix->gather_cap = 0;
        if (rc) return rc;
        uint32_t listed = 0;
        if (gather)
            CUDA_TRY(cudaMemcpyAsync(&listed, ix->ws_gather_count.p, 4, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(out, ix->ws_hits.p, (size_t)batch * k * sizeof(fsgpu_hit), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaMemcpyAsync(out_counts, ix->ws_counts.p, (size_t)batch * 4, cudaMemcpyDeviceToHost, s));
        rc = end_call_locked(ix, s);
        if (rc) return rc;
        bool retry = false;
        rc = finish_sync_call_locked(ix, s, &retry);
        if (rc) return rc;
        if (retry && !force_f16) {  // the batched int8 form bailed: same call in the f16 form
            force_f16 = true;
            continue;
        }
        if (!gather || listed <= cap) {
            if (out_used_gather) *out_used_gather = gather ? 1 : 0;
            return FSGPU_OK;
        }
        gather = false;  // more rows share the allowed hashes than the list holds: take the scan (same result)
    }
    return FSGPU_OK;
}

static int merge_top_k_impl(int device, const uint64_t* d_keys, const float* d_scores, const fsgpu_hit* d_hits_in,
                            uint32_t batch, uint32_t n_lists, uint32_t k_in, uint64_t list_stride,
                            uint64_t query_stride, uint32_t k_out, uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                            uint32_t* d_out_counts, void* stream);

extern "C" int fsgpu_merge_top_k_device(int device, const uint64_t* d_keys, const float* d_scores,
                                        uint32_t batch, uint32_t n_lists, uint32_t k_in,
                                        uint64_t list_stride, uint64_t query_stride, uint32_t k_out,
                                        uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                                        uint32_t* d_out_counts, void* stream) {
    return merge_top_k_impl(device, d_keys, d_scores, nullptr, batch, n_lists, k_in, list_stride, query_stride, k_out,
                            d_out_keys, d_out_hits, d_out_counts, stream);
}

extern "C" int fsgpu_merge_top_k_hits_device(int device, const uint64_t* d_keys, const fsgpu_hit* d_hits,
                                             uint32_t batch, uint32_t n_lists, uint32_t k_in,
                                             uint64_t list_stride, uint64_t query_stride, uint32_t k_out,
                                             uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                                             uint32_t* d_out_counts, void* stream) {
    return merge_top_k_impl(device, d_keys, nullptr, d_hits, batch, n_lists, k_in, list_stride, query_stride, k_out,
                            d_out_keys, d_out_hits, d_out_counts, stream);
}

static int merge_top_k_impl(int device, const uint64_t* d_keys, const float* d_scores, const fsgpu_hit* d_hits_in,
                            uint32_t batch, uint32_t n_lists, uint32_t k_in, uint64_t list_stride,
                            uint64_t query_stride, uint32_t k_out, uint64_t* d_out_keys, fsgpu_hit* d_out_hits,
                            uint32_t* d_out_counts, void* stream) {
    if (batch == 0) return FSGPU_OK;
    if (!d_keys) return fail(FSGPU_ERR_INVALID_CONFIG, "keys is NULL");
    if (k_out == 0 || k_in == 0 || n_lists == 0) {
        DeviceGuard g(device);
        if (d_out_counts) CUDA_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)batch * 4, (cudaStream_t)stream));
        return FSGPU_OK;
    }
    DeviceGuard g(device);
    static thread_local uint32_t* d_errs[64] = {nullptr};  // per-thread, per-device scratch flag
    if (device < 0 || device >= 64) return fail(FSGPU_ERR_INVALID_CONFIG, "device %d out of range", device);
    if (!d_errs[device]) CUDA_TRY(cudaMalloc(&d_errs[device], 4));
    uint32_t* d_err = d_errs[device];
    MergeArgs m{};
    m.keys = d_keys;
    m.scores = d_scores;
    m.hits_in = d_hits_in;
    m.list_stride = list_stride;
    m.query_stride = query_stride;
    m.n_lists = n_lists;
    m.k_in = k_in;
    m.k_out = k_out;
    m.cap = cand_capacity(k_out);
    if ((size_t)m.cap * 8 + 16 > 200 * 1024)
        return fail(FSGPU_ERR_INVALID_CONFIG, "k_out=%u exceeds the device merge window", k_out);
    m.out_keys = d_out_keys;
    m.out_hits = d_out_hits;
    m.out_counts = d_out_counts;
    m.error_flag = d_err;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemsetAsync(d_err, 0, 4, s));
    int rc = launch_merge(m, batch, s);
    if (rc) return rc;
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

// ─── gather-dot ─────────────────────────────────────────────────────────────────────────────
extern "C" int fsgpu_scores_for_rows_device(const fsgpu_index* ix, const float* d_queries, uint32_t batch,
                                            const uint32_t* d_rows, uint32_t n_per_query,
                                            float* d_out_scores, uint8_t* d_out_present, void* stream) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (batch == 0 || n_per_query == 0) return FSGPU_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ix->stream;
    dim3 grid((n_per_query + kScanWarps - 1) / kScanWarps, batch);
    scores_for_rows_kernel<<<grid, kScanThreads, 0, s>>>(ix->slab_any(), ix->is_f32(), ix->n_rows, ix->row_base, ix->dim,
                                                         d_queries, d_rows, 1u, n_per_query, ix->reduce_order,
                                                         ix->tail_fma, d_out_scores, d_out_present);
    CUDA_TRY(cudaGetLastError());
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_scores_for_hits_device(const fsgpu_index* ix, const float* d_queries, uint32_t batch,
                                            const fsgpu_hit* d_hits, uint32_t n_per_query, float* d_out_scores,
                                            uint8_t* d_out_present, void* stream) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (batch == 0 || n_per_query == 0) return FSGPU_OK;
    if (!d_queries || !d_hits || !d_out_scores) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : ix->stream;
    dim3 grid((n_per_query + kScanWarps - 1) / kScanWarps, batch);
    // the row of hit i is word 2i of the record array
    scores_for_rows_kernel<<<grid, kScanThreads, 0, s>>>(ix->slab_any(), ix->is_f32(), ix->n_rows, ix->row_base, ix->dim, d_queries,
                                                         reinterpret_cast<const uint32_t*>(d_hits), 2u, n_per_query,
                                                         ix->reduce_order, ix->tail_fma, d_out_scores, d_out_present);
    CUDA_TRY(cudaGetLastError());
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_merge_payload_device(int device, const uint64_t* d_keys, const float* d_payload, uint32_t batch,
                                          uint32_t n_lists, uint32_t k_in, uint64_t list_stride, uint64_t query_stride,
                                          uint64_t payload_list_stride, uint64_t payload_query_stride,
                                          const uint64_t* d_merged_keys, uint32_t k_out, float* d_out_payload,
                                          uint8_t* d_out_present, void* stream) {
    if (batch == 0 || k_out == 0) return FSGPU_OK;
    if (!d_keys || !d_payload || !d_merged_keys || !d_out_payload) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    dim3 grid((k_out + 255) / 256, batch);
    merge_payload_kernel<<<grid, 256, 0, s>>>(d_keys, list_stride, query_stride, n_lists, k_in, d_payload,
                                              payload_list_stride, payload_query_stride, d_merged_keys, k_out,
                                              d_out_payload, d_out_present);
    CUDA_TRY(cudaGetLastError());
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_scores_for_rows(const fsgpu_index* ix, const float* query, uint32_t dim,
                                     const uint32_t* rows, uint32_t n, float* out_scores,
                                     uint8_t* out_present) {
    if (!ix) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (dim != ix->dim) return fail(FSGPU_ERR_DIMENSION_MISMATCH, "expected %u, found %u", ix->dim, dim);
    if (n == 0) return FSGPU_OK;
    if (!query || !rows || !out_scores) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    {
        std::lock_guard<std::mutex> lock(ix->mu);
        DeviceGuard g(ix->device);
        cudaStream_t s = ix->stream;
        CUDA_TRY(ix->ws_queries.reserve((size_t)dim * 4));
        CUDA_TRY(ix->ws_rows.reserve((size_t)n * 4));
        CUDA_TRY(ix->ws_scores.reserve((size_t)n * 4));
        CUDA_TRY(ix->ws_present.reserve(n));
        int rc = begin_call_locked(ix, s, true);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(ix->ws_queries.p, query, (size_t)dim * 4, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(ix->ws_rows.p, rows, (size_t)n * 4, cudaMemcpyHostToDevice, s));
        dim3 grid((n + kScanWarps - 1) / kScanWarps, 1);
        scores_for_rows_kernel<<<grid, kScanThreads, 0, s>>>(
            ix->slab_any(), ix->is_f32(), ix->n_rows, ix->row_base, ix->dim, ix->ws_queries.as<float>(),
            ix->ws_rows.as<uint32_t>(), 1u, n, ix->reduce_order, ix->tail_fma, ix->ws_scores.as<float>(),
            ix->ws_present.as<uint8_t>());
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpyAsync(out_scores, ix->ws_scores.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        if (out_present)
            CUDA_TRY(cudaMemcpyAsync(out_present, ix->ws_present.p, n, cudaMemcpyDeviceToHost, s));
        rc = end_call_locked(ix, s);
        if (rc) return rc;
        rc = finish_sync_call_locked(ix, s, nullptr);
        if (rc) return rc;
    }
    return FSGPU_OK;
}

// ─── synthetic corpora ──────────────────────────────────────────────────────────────────────
extern "C" int fsgpu_synth_rows_device(int device, int kind, uint64_t seed_base, uint64_t row_start,
                                       uint64_t n_rows, uint32_t dim, uint32_t n_centroids, float noise,
                                       uint16_t* d_out_f16, void* stream) {
    if (n_rows == 0) return FSGPU_OK;
    if (!d_out_f16 || dim == 0) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL output or zero dim");
    if (kind != 0 && kind != 1) return fail(FSGPU_ERR_INVALID_CONFIG, "kind must be 0 (uniform) or 1 (clustered)");
    if (kind == 1 && n_centroids == 0) return fail(FSGPU_ERR_INVALID_CONFIG, "clustered corpus needs centroids");
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    float* d_cent = nullptr;
    if (kind == 1) {
        CUDA_TRY(cudaMalloc(&d_cent, (size_t)n_centroids * dim * 4));
        synth_centroids_kernel<<<(n_centroids + 63) / 64, 64, 0, s>>>(n_centroids, dim, d_cent);
    }
    synth_rows_kernel<<<(unsigned)((n_rows + 127) / 128), 128, 0, s>>>(kind, seed_base, row_start, n_rows, dim,
                                                                       n_centroids, noise, d_cent, d_out_f16);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (d_cent) cudaFree(d_cent);
    if (e != cudaSuccess) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: synth kernel failed: %s", cudaGetErrorString(e));
    return FSGPU_OK;
}

// ─── FSVI v1 reader (crates/frankensearch-index/src/lib.rs:6-43, :4049-4144, :6114) ─────────
static uint32_t crc32_ieee(const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static std::once_flag once;
    std::call_once(once, [] {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
    });
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

template <class T>
static bool rd(const std::vector<uint8_t>& d, size_t* cur, T* out) {
    if (*cur + sizeof(T) > d.size()) return false;
    memcpy(out, d.data() + *cur, sizeof(T));  // little-endian host
    *cur += sizeof(T);
    return true;
}

extern "C" int fsgpu_index_open_fsvi(const char* path, uint64_t row_start, uint64_t n_rows_or_0,
                                     const fsgpu_index_options* opts, fsgpu_index** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    if (!path) return fail(FSGPU_ERR_INVALID_CONFIG, "path is NULL");
    fsgpu_index_options o;
    if (opts) o = *opts; else fsgpu_index_options_default(&o);
    // Pending appends live in the `<path>.wal` sidecar (wal.rs:575-579); VectorIndex::open replays it
    // (lib.rs:1833-1878).  The doc ids of WAL rows are host state, so the replay is the host's job
    // (fsgpu_index_set_wal): a host that has not said it does so must not silently lose those rows.
    if (!(o.flags & FSGPU_OPEN_HOST_REPLAYS_WAL)) {
        const std::string wal = std::string(path) + ".wal";
        if (FILE* w = fopen(wal.c_str(), "rb")) {
            fseeko(w, 0, SEEK_END);
            const off_t wlen = ftello(w);
            fclose(w);
            if (wlen >= 20)  // WAL_HEADER_SIZE: anything shorter is ignored by the reference too (wal.rs:861-864)
                return fail(FSGPU_ERR_INVALID_CONFIG,
                            "%s has a WAL sidecar with pending appends: replay it on the host (fsgpu_index_set_wal) and pass "
                            "FSGPU_OPEN_HOST_REPLAYS_WAL, or compact the index first", path);
        }
    }
    FILE* f = fopen(path, "rb");
    if (!f) return fail(FSGPU_ERR_IO, "cannot open %s: %s", path, strerror(errno));
    fseeko(f, 0, SEEK_END);
    const uint64_t fsize = (uint64_t)ftello(f);  // 64-bit: a 50 M x 384 slab is 38 GB
    fseeko(f, 0, SEEK_SET);
    // header + record table + string table are read whole; the slab is streamed by row range
    std::vector<uint8_t> head((size_t)std::min<uint64_t>(fsize, 4 + 2 + 2 + 65535 + 2 + 65535 + 4 + 1 + 3 + 8 + 8 + 4));
    if (fread(head.data(), 1, head.size(), f) != head.size()) {
        fclose(f);
        return fail(FSGPU_ERR_IO, "short read on %s", path);
    }
    auto corrupt = [&](const char* why) {
        fclose(f);
        return fail(FSGPU_ERR_INDEX_CORRUPTED, "%s: %s", path, why);
    };
    size_t cur = 0;
    uint8_t magic[4];
    uint16_t version = 0, len16 = 0;
    if (head.size() < 4) return corrupt("file too small");
    memcpy(magic, head.data(), 4);
    cur = 4;
    if (memcmp(magic, "FSVI", 4) != 0) return corrupt("bad magic");
    if (!rd(head, &cur, &version)) return corrupt("truncated header");
    uint32_t dim = 0, crc_stored = 0;
    uint8_t quant = 0, reserved[3];
    uint64_t record_count = 0, vectors_offset = 0;
    size_t records_offset = 0;
    if (version == 1) {
        if (!rd(head, &cur, &len16) || cur + len16 > head.size()) return corrupt("truncated embedder_id");
        cur += len16;
        if (!rd(head, &cur, &len16) || cur + len16 > head.size()) return corrupt("truncated embedder_revision");
        cur += len16;
        if (!rd(head, &cur, &dim) || !rd(head, &cur, &quant) || !rd(head, &cur, &reserved) ||
            !rd(head, &cur, &record_count) || !rd(head, &cur, &vectors_offset))
            return corrupt("truncated header");
        const size_t crc_end = cur;
        if (!rd(head, &cur, &crc_stored)) return corrupt("truncated header crc");
        if (crc32_ieee(head.data(), crc_end) != crc_stored) return corrupt("header CRC mismatch");
        records_offset = cur;
    } else if (version == 2) {
        // FSVI v2 (identity-complete immutable artifact, lib.rs:4229-4520): a 332-byte fixed prefix —
        // header_size u32, binding schema u16 (= 1), quantization u8, flags u8 (= 0), publication nonce
        // u16, dimension u32, record_count u64, vectors_offset u64, generation {schema u16, reserved u16
        // (= 0), sequence u64, nonce [16]}, three canonical-identity lengths u32, eight SHA-256
        // fingerprints — then the three canonical identity documents and a CRC-32 of everything before it;
        // the record table starts at header_size.  The LAYOUT is read and checked here; admitting the
        // identity (fingerprint / bundle validation, lib.rs:4448-4500) is the host's decision before it
        // hands the file to the GPU.
        uint32_t header_size = 0, bundle_len = 0, space_len = 0, storage_len = 0;
        uint16_t schema = 0, nonce = 0, gen_schema = 0, gen_reserved = 0;
        uint8_t flags8 = 0, gen_nonce[16], fp[32];
        uint64_t gen_seq = 0;
        if (!rd(head, &cur, &header_size) || !rd(head, &cur, &schema) || !rd(head, &cur, &quant) || !rd(head, &cur, &flags8) ||
            !rd(head, &cur, &nonce) || !rd(head, &cur, &dim) || !rd(head, &cur, &record_count) ||
            !rd(head, &cur, &vectors_offset) || !rd(head, &cur, &gen_schema) || !rd(head, &cur, &gen_reserved) ||
            !rd(head, &cur, &gen_seq) || !rd(head, &cur, &gen_nonce) || !rd(head, &cur, &bundle_len) ||
            !rd(head, &cur, &space_len) || !rd(head, &cur, &storage_len))
            return corrupt("truncated v2 header");
        if (header_size < 336 || header_size > fsize) return corrupt("v2 header_size out of range");
        if (schema != 1) return corrupt("unsupported v2 identity binding schema");
        if (flags8 != 0) return corrupt("v2 header flags must be zero");
        if (gen_reserved != 0) return corrupt("v2 generation reserved field must be zero");
        for (uint32_t len : {bundle_len, space_len, storage_len})
            if (len == 0 || len > (1u << 20)) return corrupt("v2 canonical identity length must be non-zero and at most 1 MiB");
        for (int i = 0; i < 8; ++i) {
            if (!rd(head, &cur, &fp)) return corrupt("truncated v2 fingerprints");
            bool zero = true;
            for (uint8_t b : fp) zero = zero && b == 0;
            if (zero) return corrupt("v2 fingerprint must not be all zero");
        }
        if (cur != 332) return corrupt("v2 fixed header layout disagreement");
        if ((uint64_t)332 + bundle_len + space_len + storage_len + 4 != header_size)
            return corrupt("v2 canonical identity lengths do not end at the header CRC");
        if (head.size() < header_size) {  // identity documents can push the header past the v1-sized first read
            head.resize(header_size);
            fseeko(f, 0, SEEK_SET);
            if (fread(head.data(), 1, head.size(), f) != head.size()) return corrupt("short read (v2 header)");
        }
        memcpy(&crc_stored, head.data() + header_size - 4, 4);
        if (crc32_ieee(head.data(), header_size - 4) != crc_stored) return corrupt("v2 header CRC mismatch");
        records_offset = header_size;
        memset(reserved, 0, sizeof reserved);
    } else {
        return corrupt("unsupported FSVI version (v1 and v2 are read)");
    }
    if (quant > 1) return corrupt("unknown quantization");
    if (dim == 0) return corrupt("zero dimension");
    const uint64_t elem = quant == 1 ? 2 : 4;
    if (vectors_offset % 64 != 0) return corrupt("vector slab is not 64-byte aligned");
    if (records_offset + record_count * 16 > vectors_offset ||
        vectors_offset + record_count * dim * elem > fsize)
        return corrupt("section offsets exceed file size");
    if (row_start > record_count) {
        fclose(f);
        return fail(FSGPU_ERR_INVALID_CONFIG, "row_start beyond record_count");
    }
    const uint64_t n = n_rows_or_0 ? std::min(n_rows_or_0, record_count - row_start) : record_count - row_start;

    std::vector<uint8_t> meta(vectors_offset - records_offset);
    fseeko(f, (off_t)records_offset, SEEK_SET);
    if (!meta.empty() && fread(meta.data(), 1, meta.size(), f) != meta.size()) return corrupt("short read (records)");
    const uint8_t* strings = meta.data() + record_count * 16;
    const size_t strings_len = meta.size() - record_count * 16;
    std::vector<uint8_t> tomb((n + 7) / 8, 0);
    bool any_tomb = false;
    std::vector<uint8_t> doc_bytes;
    std::vector<uint64_t> doc_off(n + 1, 0);
    std::vector<uint64_t> hashes(n);
    for (uint64_t r = 0; r < n; ++r) {
        const uint8_t* rec = meta.data() + (row_start + r) * 16;
        uint32_t off;
        uint16_t len, flags;
        memcpy(&hashes[r], rec, 8);
        memcpy(&off, rec + 8, 4);
        memcpy(&len, rec + 12, 2);
        memcpy(&flags, rec + 14, 2);
        if ((size_t)off + len > strings_len) return corrupt("doc_id outside string table");
        doc_bytes.insert(doc_bytes.end(), strings + off, strings + off + len);
        doc_off[r + 1] = doc_bytes.size();
        if (flags & 1) {
            tomb[r >> 3] |= (uint8_t)(1u << (r & 7));
            any_tomb = true;
        }
    }
    meta.clear();
    meta.shrink_to_fit();

    o.slab_is_device = 0;
    o.tail_fma = 1;  // file-backed index scores with the bytes kernels (search.rs:1283, :1310)
    if (o.row_base == 0) o.row_base = row_start;
    DeviceGuard g(o.device);
    fsgpu_index* ix = new fsgpu_index();
    int rc = index_alloc_common(ix, &o, n, dim);
    if (rc) {
        fclose(f);
        fsgpu_index_destroy(ix);
        return rc;
    }
    // the slab goes file -> pinned staging -> device in 32 MiB pieces (two buffers in flight): no host copy
    // of the whole slab, no pageable memcpy
    const uint64_t slab_bytes = n * dim * elem;
    cudaError_t e = cudaSuccess;
    void* d_dst = nullptr;
    if (slab_bytes) {
        e = cudaMalloc(&d_dst, slab_bytes);
        if (e == cudaSuccess) {
            if (quant == 1) {
                ix->d_slab = static_cast<uint16_t*>(d_dst);
                ix->owns_slab = true;
            } else {
                ix->d_slab_f32 = static_cast<float*>(d_dst);
            }
        }
    }
    constexpr size_t kPiece = (size_t)32 << 20;
    uint8_t* stage[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    for (int i = 0; i < 2 && e == cudaSuccess && slab_bytes; ++i) {
        e = cudaHostAlloc(reinterpret_cast<void**>(&stage[i]), (size_t)std::min<uint64_t>(kPiece, slab_bytes), cudaHostAllocDefault);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    bool short_read = false;
    if (e == cudaSuccess && slab_bytes) {
        fseeko(f, (off_t)(vectors_offset + row_start * dim * elem), SEEK_SET);
        int which = 0;
        for (uint64_t off = 0; off < slab_bytes && e == cudaSuccess; off += kPiece, which ^= 1) {
            const size_t len = (size_t)std::min<uint64_t>(kPiece, slab_bytes - off);
            if (off >= 2 * kPiece) e = cudaEventSynchronize(done[which]);  // the buffer's previous copy has landed
            if (e != cudaSuccess) break;
            if (fread(stage[which], 1, len, f) != len) {
                short_read = true;
                break;
            }
            e = cudaMemcpyAsync(static_cast<uint8_t*>(d_dst) + off, stage[which], len, cudaMemcpyHostToDevice, ix->stream);
            if (e == cudaSuccess) e = cudaEventRecord(done[which], ix->stream);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ix->stream);
    }
    fclose(f);
    for (int i = 0; i < 2; ++i) {
        if (stage[i]) cudaFreeHost(stage[i]);
        if (done[i]) cudaEventDestroy(done[i]);
    }
    if (short_read || e != cudaSuccess) {
        fsgpu_index_destroy(ix);
        if (short_read) return fail(FSGPU_ERR_INDEX_CORRUPTED, "%s: short read (slab)", path);
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: slab upload failed: %s", cudaGetErrorString(e));
    }
    rc = upload_tombstones(ix, any_tomb ? tomb.data() : nullptr);
    if (!rc) rc = index_finish_setup(ix);
    if (rc) {
        fsgpu_index_destroy(ix);
        return rc;
    }
    ix->doc_bytes.swap(doc_bytes);
    ix->doc_off.swap(doc_off);
    if (n) {  // record-table hashes (FNV-1a of the doc id, lib.rs:6120-6127) for device-side hash filters
        rc = fsgpu_index_set_doc_hashes(ix, hashes.data());
        if (rc) {
            fsgpu_index_destroy(ix);
            return rc;
        }
    }
    *out = ix;
    return FSGPU_OK;
}

