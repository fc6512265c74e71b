// fsgpu_minilm.cu — C ABI of the MiniLM-L6-v2 encoder over minilm_kernels.cuh (weights upload,
// activation workspaces, the per-layer launch sequence).  No CPU compute path.
#include <cmath>
#include <cstring>
#include <map>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "fsgpu.h"
#include "fsgpu_common.cuh"
#include "fsgpu_host.cuh"
#include "minilm_fast_kernels.cuh"

using namespace fsgpu;

// ─── MiniLM-L6 encoder ──────────────────────────────────────────────────────────────────────
namespace {
struct SplitMat {  // an f32 matrix carried as hi + lo f16 halves, with its TMA descriptors
    __half *hi = nullptr, *lo = nullptr;
    CUtensorMap tm_hi, tm_lo;      // [128 rows x 64] boxes
    CUtensorMap tm64_hi, tm64_lo;  // [64 rows x 64] boxes (weights: the 128-wide tail tile of the pair GEMM)
    uint64_t rows = 0;
    uint32_t cols = 0;
};
struct MiniLmLayer {
    SplitMat qkv, attn_out, ffn_in, ffn_out;
    float *qkv_b = nullptr, *attn_out_b = nullptr, *attn_ln_g = nullptr, *attn_ln_b = nullptr, *ffn_in_b = nullptr,
          *ffn_out_b = nullptr, *ffn_ln_g = nullptr, *ffn_ln_b = nullptr;
};
}  // namespace

struct fsgpu_minilm {
    int device = 0, num_sms = 0;
    uint32_t vocab = 0, max_pos = 0, n_layers = 0, hidden = 0, inter = 0;
    float eps = 1e-12f;
    float *word = nullptr, *pos = nullptr, *type0 = nullptr, *emb_g = nullptr, *emb_b = nullptr;
    std::vector<MiniLmLayer> layers;
    std::vector<void*> owned;  // every device allocation of the weights
    cudaStream_t stream = nullptr;
    mutable std::mutex mu;
    // activations (grow-only, guarded by mu)
    mutable DevBuf ws_h32, ws_pre32, ws_qkv32, ws_ids, ws_lens, ws_out;
    mutable SplitMat act_h, act_ctx, act_ffn;
    mutable uint64_t act_rows = 0;
    // f16 form (minilm_fast_kernels.cuh): one f16 copy of every activation + the descriptors of its GEMMs
    mutable DevBuf f_h, f_qkv, f_ctx, f_ffn;
    mutable CUtensorMap f_tm_h, f_tm_ctx, f_tm_ffn;             // A operands: [128 rows x 64] boxes
    mutable CUtensorMap f_tm_qkv_out, f_tm_ffn_out, f_tm_pre;  // stores: f16 [64 x 32] boxes, f32 [32 x 32] boxes
    mutable CUtensorMap f_tm_h_out;                             // store: f16 [64 x 32] boxes over h (fused FFN + LayerNorm)
    mutable DevBuf f_offs;                                   // packed rows: [batch + 1] prefix sums of the lengths
    mutable const uint32_t* f_m_ptr = nullptr;                  // != nullptr while a packed forward is being enqueued: &offs[batch]
    mutable uint64_t f_rows = 0;
    mutable float* f_pre_ptr = nullptr;
    // the activation buffers are shared by every call: a call on another stream waits for the previous call's work
    mutable cudaEvent_t ev_last = nullptr;
    mutable cudaStream_t last_stream = nullptr;
    mutable bool have_last = false;
    // small batches replay a captured CUDA graph of the forward (44 launches of a few microseconds each are bound by
    // the host's launch calls): one graph per (batch, max_len, variant), inputs / outputs staged in fixed buffers
    struct FastGraph {
        cudaGraphExec_t exec = nullptr;
        const void* bufs[10] = {};
    };
    mutable std::map<uint64_t, FastGraph> f_graphs;
    mutable DevBuf g_ids, g_lens, g_out;
    mutable bool profiling = false;
    mutable fsgpu_minilm_profile prof{};
    mutable std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_pending, ev_free;
};

static int minilm_upload(fsgpu_minilm* e, const float* host, size_t count, float** dev) {
    if (!host) return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: a weight pointer is NULL");
    CUDA_TRY(cudaMalloc(dev, count * 4));
    e->owned.push_back(*dev);
    CUDA_TRY(h2d_complete(*dev, host, count * 4));
    return FSGPU_OK;
}

static int minilm_upload_split(fsgpu_minilm* e, const float* host, uint64_t rows, uint32_t cols, SplitMat* m) {
    float* tmp = nullptr;
    if (!host) return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: a weight pointer is NULL");
    const size_t n = (size_t)rows * cols;
    CUDA_TRY(cudaMalloc(&tmp, n * 4));
    cudaError_t err = h2d_complete(tmp, host, n * 4);
    if (err == cudaSuccess) err = cudaMalloc(&m->hi, n * 2);
    if (err == cudaSuccess) e->owned.push_back(m->hi);
    if (err == cudaSuccess) err = cudaMalloc(&m->lo, n * 2);
    if (err == cudaSuccess) e->owned.push_back(m->lo);
    if (err == cudaSuccess) {
        split_f16_kernel<<<e->num_sms * 4, 256, 0, e->stream>>>(tmp, n, m->hi, m->lo);
        err = cudaGetLastError();
        if (err == cudaSuccess) err = cudaStreamSynchronize(e->stream);
    }
    cudaFree(tmp);
    if (err != cudaSuccess) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: minilm weight upload failed: %s", cudaGetErrorString(err));
    m->rows = rows;
    m->cols = cols;
    if (!make_f16_tile_map(&m->tm_hi, m->hi, rows, cols) || !make_f16_tile_map(&m->tm_lo, m->lo, rows, cols) ||
        !make_f16_tile_map(&m->tm64_hi, m->hi, rows, cols, 64) ||
        !make_f16_tile_map(&m->tm64_lo, m->lo, rows, cols, 64))
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: cuTensorMapEncodeTiled failed for a minilm weight");
    return FSGPU_OK;
}

extern "C" void fsgpu_minilm_destroy(fsgpu_minilm* e) {
    if (!e) return;
    {
        DeviceGuard g(e->device);
        if (e->stream) cudaStreamSynchronize(e->stream);
        for (void* p : e->owned) cudaFree(p);
        for (SplitMat* m : {&e->act_h, &e->act_ctx, &e->act_ffn}) {
            if (m->hi) cudaFree(m->hi);
            if (m->lo) cudaFree(m->lo);
        }
        for (DevBuf* b : {&e->ws_h32, &e->ws_pre32, &e->ws_qkv32, &e->ws_ids, &e->ws_lens, &e->ws_out, &e->f_h, &e->f_qkv, &e->f_ctx,
                          &e->f_ffn, &e->g_ids, &e->g_lens, &e->g_out, &e->f_offs})
            b->release();
        if (e->ev_last) cudaEventDestroy(e->ev_last);
        for (auto& kv : e->f_graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        for (auto* v : {&e->ev_pending, &e->ev_free})
            for (auto& ev : *v) {
                cudaEventDestroy(ev.first);
                cudaEventDestroy(ev.second);
            }
        if (e->stream) cudaStreamDestroy(e->stream);
    }
    delete e;
}

extern "C" int fsgpu_minilm_create(const fsgpu_minilm_weights* w, int device, fsgpu_minilm** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    if (!w || !w->layers) return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: weights is NULL");
    if (w->hidden != kHidden || w->heads != kHeads)
        return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: only hidden=384 / heads=12 (all-MiniLM-L6-v2 geometry) is built, got %u / %u",
                    w->hidden, w->heads);
    if (w->intermediate == 0 || w->intermediate % 128 != 0 || w->n_layers == 0 || w->vocab_size == 0 || w->max_positions == 0)
        return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: bad geometry (intermediate=%u layers=%u vocab=%u positions=%u)",
                    w->intermediate, w->n_layers, w->vocab_size, w->max_positions);
    int ndev = 0;
    int rc = fsgpu_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(FSGPU_ERR_INVALID_CONFIG, "device %d not present", device);
    if (!fsgpu_tma_available()) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: cuTensorMapEncodeTiled is unavailable");
    DeviceGuard g(device);
    fsgpu_minilm* e = new fsgpu_minilm();
    e->device = device;
    cudaDeviceProp prop;
    cudaError_t err = cudaGetDeviceProperties(&prop, device);
    if (err == cudaSuccess && prop.major < 10) {
        delete e;
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
        delete e;
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: %s", cudaGetErrorString(err));
    }
    e->num_sms = prop.multiProcessorCount;
    e->vocab = w->vocab_size;
    e->max_pos = w->max_positions;
    e->n_layers = w->n_layers;
    e->hidden = w->hidden;
    e->inter = w->intermediate;
    e->eps = w->ln_eps;
    const uint32_t H = kHidden, I = w->intermediate;
    rc = minilm_upload(e, w->word_emb, (size_t)w->vocab_size * H, &e->word);
    if (!rc) rc = minilm_upload(e, w->pos_emb, (size_t)w->max_positions * H, &e->pos);
    if (!rc) rc = minilm_upload(e, w->type_emb, H, &e->type0);
    if (!rc) rc = minilm_upload(e, w->emb_ln_g, H, &e->emb_g);
    if (!rc) rc = minilm_upload(e, w->emb_ln_b, H, &e->emb_b);
    e->layers.resize(w->n_layers);
    for (uint32_t i = 0; i < w->n_layers && !rc; ++i) {
        const fsgpu_minilm_layer_weights& s = w->layers[i];
        MiniLmLayer& d = e->layers[i];
        rc = minilm_upload_split(e, s.qkv_w, 3 * H, H, &d.qkv);
        if (!rc) rc = minilm_upload_split(e, s.attn_out_w, H, H, &d.attn_out);
        if (!rc) rc = minilm_upload_split(e, s.ffn_in_w, I, H, &d.ffn_in);
        if (!rc) rc = minilm_upload_split(e, s.ffn_out_w, H, I, &d.ffn_out);
        if (!rc) rc = minilm_upload(e, s.qkv_b, 3 * H, &d.qkv_b);
        if (!rc) rc = minilm_upload(e, s.attn_out_b, H, &d.attn_out_b);
        if (!rc) rc = minilm_upload(e, s.attn_ln_g, H, &d.attn_ln_g);
        if (!rc) rc = minilm_upload(e, s.attn_ln_b, H, &d.attn_ln_b);
        if (!rc) rc = minilm_upload(e, s.ffn_in_b, I, &d.ffn_in_b);
        if (!rc) rc = minilm_upload(e, s.ffn_out_b, H, &d.ffn_out_b);
        if (!rc) rc = minilm_upload(e, s.ffn_ln_g, H, &d.ffn_ln_g);
        if (!rc) rc = minilm_upload(e, s.ffn_ln_b, H, &d.ffn_ln_b);
    }
    if (rc) {
        fsgpu_minilm_destroy(e);
        return rc;
    }
    *out = e;
    return FSGPU_OK;
}

static int minilm_reserve_act(const fsgpu_minilm* e, SplitMat* m, uint64_t rows, uint32_t cols) {
    if (m->hi) cudaFree(m->hi);
    if (m->lo) cudaFree(m->lo);
    m->hi = m->lo = nullptr;
    CUDA_TRY(cudaMalloc(&m->hi, (size_t)rows * cols * 2));
    CUDA_TRY(cudaMalloc(&m->lo, (size_t)rows * cols * 2));
    m->rows = rows;
    m->cols = cols;
    if (!make_f16_tile_map(&m->tm_hi, m->hi, rows, cols) || !make_f16_tile_map(&m->tm_lo, m->lo, rows, cols))
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: cuTensorMapEncodeTiled failed for a minilm activation");
    (void)e;
    return FSGPU_OK;
}

// C[m, n] = A * W^T (+ bias, residual, GELU); caller holds e->mu.
static int minilm_gemm(const fsgpu_minilm* e, const SplitMat& a, const SplitMat& w, uint32_t m, const float* bias,
                       const float* residual, float* out_f32, __half* out_hi, __half* out_lo, int gelu,
                       uint32_t products, cudaStream_t stream) {
    GemmArgs ga{};
    ga.m = m;
    ga.n = (uint32_t)w.rows;
    ga.k = w.cols;
    ga.products = products;
    ga.n_stages = products == 3 ? 3 : 6;
    ga.bias = bias;
    ga.residual = residual;
    ga.out_f32 = out_f32;
    ga.out_hi = out_hi;
    ga.out_lo = out_lo;
    ga.gelu = gelu;
    const size_t smem = gemm_smem_bytes(ga.n_stages, products);
    CUDA_TRY(cudaFuncSetAttribute(gemm_f16split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)gemm_smem_bytes(6, 1)));
    // CTA-pair tiles (256 x 256) are opt-in (FSGPU_MINILM_PAIR=1): measured 9 % slower than 128 x 128
    // tiles at 1024 x 32 tokens — the epilogue's global loads/stores, not tile traffic, bound these
    // K = 384 GEMMs (profiles/r01_minilm_gemm_attn_out_ncu.json)
    const bool pair = m >= 2 * kGemmTileM && e->num_sms >= 2 && env_int("FSGPU_MINILM_PAIR", 0) != 0;
    uint32_t tiles, grid;
    if (pair) {
        CUDA_TRY(cudaFuncSetAttribute(gemm_f16split_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)gemm_smem_bytes(6, 1)));
        tiles = ((m + 2 * kGemmTileM - 1) / (2 * kGemmTileM)) * ((ga.n + kGemmPairN - 1) / kGemmPairN);
        grid = 2 * std::min<uint32_t>(tiles, (uint32_t)e->num_sms / 2);
    } else {
        tiles = ((m + kGemmTileM - 1) / kGemmTileM) * (ga.n / kGemmTileN);
        grid = std::min<uint32_t>(tiles, (uint32_t)e->num_sms);
    }
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (e->profiling) {
        if (!e->ev_free.empty()) {
            ev = e->ev_free.back();
            e->ev_free.pop_back();
        } else {
            CUDA_TRY(cudaEventCreate(&ev.first));
            CUDA_TRY(cudaEventCreate(&ev.second));
        }
        CUDA_TRY(cudaEventRecord(ev.first, stream));
    }
    if (pair)
        gemm_f16split_pair_kernel<<<grid, kGemmThreads, smem, stream>>>(a.tm_hi, a.tm_lo, w.tm_hi, w.tm_lo, w.tm64_hi,
                                                                        w.tm64_lo, ga);
    else
        gemm_f16split_kernel<<<grid, kGemmThreads, smem, stream>>>(a.tm_hi, a.tm_lo, w.tm_hi, w.tm_lo, ga);
    CUDA_TRY(cudaGetLastError());
    if (e->profiling) {
        CUDA_TRY(cudaEventRecord(ev.second, stream));
        e->ev_pending.push_back(ev);
    }
    e->prof.gemm_launches += 1;
    e->prof.gemm_flops += 2.0 * (double)m * ga.n * ga.k * products;
    return FSGPU_OK;
}

// One linear of the f16 form: out = A W^T + bias (mode 0: f16, 1: f16 after erf-GELU, 2: f32).
static int minilm_fast_gemm(const fsgpu_minilm* e, const CUtensorMap& tm_a, const SplitMat& w, const CUtensorMap& tm_out,
                            uint32_t m, const float* bias, int mode, cudaStream_t stream) {
    FastGemmArgs ga{};
    ga.m = m;
    ga.n = (uint32_t)w.rows;
    ga.k = w.cols;
    ga.bias = bias;
    ga.mode = mode;
    ga.m_ptr = e->f_m_ptr;
    const uint32_t tiles = ((m + 127u) / 128u) * (ga.n / 128u);
    const uint32_t grid = std::min<uint32_t>(tiles, (uint32_t)e->num_sms);
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (e->profiling) {
        if (!e->ev_free.empty()) {
            ev = e->ev_free.back();
            e->ev_free.pop_back();
        } else {
            CUDA_TRY(cudaEventCreate(&ev.first));
            CUDA_TRY(cudaEventCreate(&ev.second));
        }
        CUDA_TRY(cudaEventRecord(ev.first, stream));
    }
    gemm_f16_fast_kernel<<<grid, kFastThreads, fast_gemm_smem_bytes(), stream>>>(tm_a, w.tm_hi, tm_out, ga);
    CUDA_TRY(cudaGetLastError());
    if (e->profiling) {
        CUDA_TRY(cudaEventRecord(ev.second, stream));
        e->ev_pending.push_back(ev);
    }
    e->prof.gemm_launches += 1;
    e->prof.gemm_flops += 2.0 * (double)m * ga.n * ga.k;
    return FSGPU_OK;
}

// The same linear for K <= 384 on CTA pairs with the activation tile resident (gemm_f16_ares_pair_kernel).
static int minilm_ares_gemm(const fsgpu_minilm* e, const CUtensorMap& tm_a, const SplitMat& w, const CUtensorMap& tm_out,
                            uint32_t m, const float* bias, int mode, cudaStream_t stream) {
    AresGemmArgs ga{};
    ga.m = m;
    ga.n = (uint32_t)w.rows;
    ga.k = w.cols;
    ga.bias = bias;
    ga.m_ptr = e->f_m_ptr;
    ga.mode = mode | (env_int("FSGPU_MINILM_DBG", 0) << 4);
    ga.k_chunks = ga.k > kAresMaxKb * kMmaKBlock ? ga.k / (kAresMaxKb * kMmaKBlock) : 1;
    if (ga.k % ga.k_chunks != 0 || (ga.k / ga.k_chunks) % kMmaKBlock != 0 || (ga.k_chunks > 1 && ga.n > 512))
        return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: linear %u x %u outside the pair GEMM's range", ga.n, ga.k);
    const uint32_t n_kb = ga.k / ga.k_chunks / kMmaKBlock;
    if (ga.n > kAresMaxN || n_kb > kAresMaxKb) return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: linear %u x %u outside the pair GEMM's range", ga.n, ga.k);
    ga.n_stages = (uint32_t)std::min<size_t>(8, (227 * 1024 - ares_gemm_smem_bytes(n_kb, 0)) / kMmaTileBytes);
    const size_t smem = ares_gemm_smem_bytes(n_kb, ga.n_stages);
    const uint32_t items = ((m + 255u) / 256u) * (ga.k_chunks > 1 ? 1u : (ga.n + 255u) / 256u);  // units a pair can own
    const uint32_t grid = 2 * std::min<uint32_t>(items, (uint32_t)e->num_sms / 2);
    std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
    if (e->profiling) {
        if (!e->ev_free.empty()) {
            ev = e->ev_free.back();
            e->ev_free.pop_back();
        } else {
            CUDA_TRY(cudaEventCreate(&ev.first));
            CUDA_TRY(cudaEventCreate(&ev.second));
        }
        CUDA_TRY(cudaEventRecord(ev.first, stream));
    }
    gemm_f16_ares_pair_kernel<<<grid, kAresThreads, smem, stream>>>(tm_a, w.tm_hi, w.tm64_hi, tm_out, ga);
    CUDA_TRY(cudaGetLastError());
    if (e->profiling) {
        CUDA_TRY(cudaEventRecord(ev.second, stream));
        e->ev_pending.push_back(ev);
    }
    e->prof.gemm_launches += 1;
    e->prof.gemm_flops += 2.0 * (double)m * ga.n * ga.k;
    return FSGPU_OK;
}

// The launch sequence of the f16 form (buffers reserved and descriptors built by the caller): also what a graph captures.
static int minilm_fast_enqueue(const fsgpu_minilm* e, const int32_t* d_ids, const int32_t* d_lens, uint32_t batch,
                               uint32_t max_len, float* d_out, cudaStream_t s, bool ares, bool ffn_out_pair, bool packed) {
    const uint64_t rows = (uint64_t)batch * max_len;
    const uint32_t m = (uint32_t)rows;
    const unsigned row_blocks = (unsigned)((rows + 7) / 8);
    const bool ffn_fused = ares && e->inter == 1536 && env_int("FSGPU_MINILM_FFN_FUSED", 1) != 0;
    __half* h16 = e->f_h.as<__half>();
    __half* qkv16 = e->f_qkv.as<__half>();
    __half* ctx16 = e->f_ctx.as<__half>();
    float* pre32 = e->ws_pre32.as<float>();
    float* h32 = e->ws_h32.as<float>();
    // packed rows (FSGPU_MINILM_PACKED, with the pair GEMMs): sequence b owns rows [offs[b], offs[b] + len_b); every kernel
    // below reads the row count offs[batch] on the device — the grids stay sized for batch * max_len
    uint32_t* offs = packed ? e->f_offs.as<uint32_t>() : nullptr;
    struct MPtrScope {  // the launchers read e->f_m_ptr
        const fsgpu_minilm* e;
        ~MPtrScope() { e->f_m_ptr = nullptr; }
    } m_scope{e};
    e->f_m_ptr = packed ? offs + batch : nullptr;
    if (packed) {
        minilm_offsets_kernel<<<1, 1024, 0, s>>>(d_lens, batch, max_len, offs);
        CUDA_TRY(cudaGetLastError());
    }
    minilm_fast_embed_kernel<<<row_blocks, 256, 0, s>>>(d_ids, batch, max_len, e->vocab, e->word, e->pos, e->type0, e->emb_g,
                                                        e->emb_b, e->eps, h16, offs);
    CUDA_TRY(cudaGetLastError());
    for (uint32_t li = 0; li < e->n_layers; ++li) {
        const MiniLmLayer& L = e->layers[li];
        const bool last = li + 1 == e->n_layers;
        // K = 384 linears: CTA pairs with the activation tile resident (FSGPU_MINILM_ARES=0: 128 x 128 tiles)
        auto lin384 = ares ? minilm_ares_gemm : minilm_fast_gemm;
        int rc = lin384(e, e->f_tm_h, L.qkv, e->f_tm_qkv_out, m, L.qkv_b, 0, s);
        if (rc) return rc;
        minilm_fast_attention_kernel<<<(batch * kHeads + 3) / 4, 128, 0, s>>>(qkv16, d_lens, batch, max_len, ctx16, offs);
        CUDA_TRY(cudaGetLastError());
        rc = lin384(e, e->f_tm_ctx, L.attn_out, e->f_tm_pre, m, L.attn_out_b, 2, s);
        if (rc) return rc;
        minilm_fast_ln_kernel<<<row_blocks, 256, 0, s>>>(pre32, h16, rows, L.attn_ln_g, L.attn_ln_b, e->eps, nullptr, e->f_m_ptr);
        CUDA_TRY(cudaGetLastError());
        if (ffn_fused) {  // FFN-in -> GELU -> FFN-out in one kernel: the [rows x 1536] intermediate stays on the SM
            FfnArgs fa{};
            fa.m = m;
            fa.m_ptr = e->f_m_ptr;
            fa.bias1 = L.ffn_in_b;
            fa.bias2 = L.ffn_out_b;
            fa.dbg = (uint32_t)env_int("FSGPU_MINILM_FFN_DBG", 0);
            const bool ln_fused = env_int("FSGPU_MINILM_FFN_LN", 1) != 0;  // residual + LayerNorm in the kernel's final epilogue
            if (ln_fused) {
                fa.h16 = h16;
                fa.ln_g = L.ffn_ln_g;
                fa.ln_b = L.ffn_ln_b;
                fa.h32 = last ? h32 : nullptr;
                fa.eps = e->eps;
            }
            long long* d_ts = nullptr;
            if (li == 0 && env_int("FSGPU_MINILM_FFN_TS", 0) != 0) {  // debugging aid (never with a captured graph)
                CUDA_TRY(cudaMalloc(&d_ts, 24 * 8 * sizeof(long long)));
                CUDA_TRY(cudaMemsetAsync(d_ts, 0, 24 * 8 * sizeof(long long), s));
                fa.ts = d_ts;
            }
            const uint32_t m_tiles = (m + 255u) / 256u;
            const uint32_t grid = 2 * std::min<uint32_t>(m_tiles, (uint32_t)e->num_sms / 2);
            std::pair<cudaEvent_t, cudaEvent_t> ev{nullptr, nullptr};
            if (e->profiling) {
                if (!e->ev_free.empty()) {
                    ev = e->ev_free.back();
                    e->ev_free.pop_back();
                } else {
                    CUDA_TRY(cudaEventCreate(&ev.first));
                    CUDA_TRY(cudaEventCreate(&ev.second));
                }
                CUDA_TRY(cudaEventRecord(ev.first, s));
            }
            ffn_fused_pair_kernel<<<grid, kFfnThreads, ffn_fused_smem_bytes(), s>>>(e->f_tm_h, L.ffn_in.tm64_hi, L.ffn_out.tm64_hi,
                                                                                    e->f_tm_pre, e->f_tm_h_out, fa);
            CUDA_TRY(cudaGetLastError());
            if (e->profiling) {
                CUDA_TRY(cudaEventRecord(ev.second, s));
                e->ev_pending.push_back(ev);
            }
            if (d_ts) {
                long long t[24 * 8];
                CUDA_TRY(cudaStreamSynchronize(s));
                CUDA_TRY(cudaMemcpy(t, d_ts, sizeof(t), cudaMemcpyDeviceToHost));
                cudaFree(d_ts);
                const long long t0 = t[0];
                fprintf(stderr, "[ffn ts] chunk: issuer acc1_empty g1_issued g_full g2_issued | epilogue acc1_full ld_done g_empty g_written (cycles)\n");
                for (int c = 0; c < 24; ++c) {
                    fprintf(stderr, "[ffn ts] %2d:", c);
                    for (int i = 0; i < 8; ++i) fprintf(stderr, " %7lld%s", t[c * 8 + i] ? t[c * 8 + i] - t0 : -1, i == 3 ? " |" : "");
                    fprintf(stderr, "\n");
                }
            }
            e->prof.gemm_launches += 2;  // two linears
            e->prof.gemm_flops += 2.0 * 2.0 * (double)m * kHidden * e->inter;
            if (!ln_fused) {
                minilm_fast_ln_kernel<<<row_blocks, 256, 0, s>>>(pre32, h16, rows, L.ffn_ln_g, L.ffn_ln_b, e->eps, last ? h32 : nullptr, e->f_m_ptr);
                CUDA_TRY(cudaGetLastError());
            }
            e->prof.other_launches += ln_fused ? 2 : 3;
            continue;
        }
        rc = lin384(e, e->f_tm_h, L.ffn_in, e->f_tm_ffn_out, m, L.ffn_in_b, 1, s);
        if (rc) return rc;
        // FFN-out (K = 1536): 128 x 128 tiles with both operands streamed; the pair kernel's K-chunked mode
        // (FSGPU_MINILM_ARES_FFN_OUT=1) measures the same 54 us per layer — two waves of whole 256-row tiles
        rc = (ffn_out_pair ? minilm_ares_gemm : minilm_fast_gemm)(e, e->f_tm_ffn, L.ffn_out, e->f_tm_pre, m, L.ffn_out_b, 2, s);
        if (rc) return rc;
        minilm_fast_ln_kernel<<<row_blocks, 256, 0, s>>>(pre32, h16, rows, L.ffn_ln_g, L.ffn_ln_b, e->eps, last ? h32 : nullptr, e->f_m_ptr);
        CUDA_TRY(cudaGetLastError());
        e->prof.other_launches += 3;
    }
    minilm_pool_kernel<<<batch, 128, 0, s>>>(h32, d_lens, max_len, d_out, offs);
    CUDA_TRY(cudaGetLastError());
    e->prof.other_launches += 2;
    return FSGPU_OK;
}

// The f16 form of the forward (max_len <= 32).  Caller holds e->mu and has selected the device.
static int minilm_embed_fast_locked(const fsgpu_minilm* e, const int32_t* d_ids, const int32_t* d_lens, uint32_t batch,
                                    uint32_t max_len, float* d_out, cudaStream_t s, bool sync) {
    const uint64_t rows = (uint64_t)batch * max_len;
    const uint32_t m = (uint32_t)rows, H = kHidden, I = e->inter;
    CUDA_TRY(e->ws_h32.reserve(rows * H * 4));
    CUDA_TRY(e->ws_pre32.reserve(rows * H * 4));
    void* before[4] = {e->f_h.p, e->f_qkv.p, e->f_ctx.p, e->f_ffn.p};
    CUDA_TRY(e->f_h.reserve(rows * H * 2));
    CUDA_TRY(e->f_qkv.reserve(rows * 3 * H * 2));
    CUDA_TRY(e->f_ctx.reserve(rows * H * 2));
    CUDA_TRY(e->f_ffn.reserve(rows * I * 2));
    if (rows != e->f_rows || before[0] != e->f_h.p || before[1] != e->f_qkv.p || before[2] != e->f_ctx.p ||
        before[3] != e->f_ffn.p || e->ws_pre32.as<float>() != e->f_pre_ptr) {  // descriptors carry base and row count
        const bool ok = make_f16_tile_map(&e->f_tm_h, e->f_h.p, rows, H) && make_f16_tile_map(&e->f_tm_ctx, e->f_ctx.p, rows, H) &&
                        make_f16_tile_map(&e->f_tm_ffn, e->f_ffn.p, rows, I) &&
                        make_tile_map_2d(&e->f_tm_qkv_out, e->f_qkv.p, rows, 3 * H, 2, 64, 32) &&
                        make_tile_map_2d(&e->f_tm_ffn_out, e->f_ffn.p, rows, I, 2, 64, 32) &&
                        make_tile_map_2d(&e->f_tm_pre, e->ws_pre32.p, rows, H, 4, 32, 32) &&
                        make_tile_map_2d(&e->f_tm_h_out, e->f_h.p, rows, H, 2, 64, 32);
        if (!ok) return fail(FSGPU_ERR_SUBSYSTEM, "gpu: cuTensorMapEncodeTiled failed for a minilm activation");
        e->f_rows = rows;
        e->f_pre_ptr = e->ws_pre32.as<float>();
    }
    CUDA_TRY(cudaFuncSetAttribute(gemm_f16_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_gemm_smem_bytes()));
    CUDA_TRY(cudaFuncSetAttribute(gemm_f16_ares_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CUDA_TRY(cudaFuncSetAttribute(ffn_fused_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ffn_fused_smem_bytes()));
    const bool ares = env_int("FSGPU_MINILM_ARES", 1) != 0 && e->num_sms >= 2 && m >= 256;
    const bool ffn_out_pair = ares && I % (kAresMaxKb * kMmaKBlock) == 0 && env_int("FSGPU_MINILM_ARES_FFN_OUT", 0) != 0;
    const bool packed = ares && env_int("FSGPU_MINILM_PACKED", 1) != 0;  // no padding rows (see minilm_offsets_kernel)
    CUDA_TRY(e->f_offs.reserve(((size_t)batch + 1) * 4));

    // Small batches (a query or a handful: <= 4096 token rows): the forward is bound by the host's 44 launch calls, so a
    // captured graph of it is replayed; ids / lens / output go through fixed staging buffers.  FSGPU_MINILM_GRAPH=0: off.
    bool use_graph = !e->profiling && rows <= 4096 && env_int("FSGPU_MINILM_GRAPH", 1) != 0;
    if (use_graph) {  // a caller that is itself capturing this stream gets the plain launches in its graph
        cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(s, &st) != cudaSuccess || st != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            use_graph = false;
        }
    }
    if (!use_graph) {
        int rc = minilm_fast_enqueue(e, d_ids, d_lens, batch, max_len, d_out, s, ares, ffn_out_pair, packed);
        if (rc) return rc;
        if (sync) CUDA_TRY(cudaStreamSynchronize(s));
        return FSGPU_OK;
    }
    CUDA_TRY(e->g_ids.reserve(4096 * 4));
    CUDA_TRY(e->g_lens.reserve(4096 * 4));
    CUDA_TRY(e->g_out.reserve((size_t)4096 * H * 4));
    const uint64_t key = ((uint64_t)batch << 32) | ((uint64_t)max_len << 8) | (ares ? 1u : 0u) | (ffn_out_pair ? 2u : 0u) |
                         (env_int("FSGPU_MINILM_FFN_FUSED", 1) != 0 ? 4u : 0u) | (env_int("FSGPU_MINILM_FFN_LN", 1) != 0 ? 8u : 0u) |
                         (packed ? 16u : 0u);
    const void* bufs[10] = {e->f_h.p, e->f_qkv.p, e->f_ctx.p, e->f_ffn.p, e->ws_pre32.p, e->ws_h32.p, e->g_ids.p, e->g_lens.p, e->g_out.p, e->f_offs.p};
    if (e->f_graphs.size() >= 64 && !e->f_graphs.count(key)) {  // bounded cache: shapes are few in practice
        for (auto& kv : e->f_graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        e->f_graphs.clear();
    }
    fsgpu_minilm::FastGraph& fg = e->f_graphs[key];
    if (fg.exec && memcmp(fg.bufs, bufs, sizeof(bufs)) != 0) {  // a buffer moved (grew): the captured pointers are stale
        cudaGraphExecDestroy(fg.exec);
        fg.exec = nullptr;
    }
    bool captured_now = false;
    if (!fg.exec) {
        captured_now = true;  // (the capture ran the launch sequence once: its counters already cover this call)
        cudaGraph_t graph = nullptr;
        CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        int rc = minilm_fast_enqueue(e, e->g_ids.as<int32_t>(), e->g_lens.as<int32_t>(), batch, max_len, e->g_out.as<float>(), s,
                                     ares, ffn_out_pair, packed);
        cudaError_t ce = cudaStreamEndCapture(s, &graph);
        if (rc || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            return rc ? rc : fail(FSGPU_ERR_SUBSYSTEM, "gpu: minilm graph capture failed: %s", cudaGetErrorString(ce));
        }
        ce = cudaGraphInstantiate(&fg.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) {
            fg.exec = nullptr;
            return fail(FSGPU_ERR_SUBSYSTEM, "gpu: minilm graph instantiation failed: %s", cudaGetErrorString(ce));
        }
        memcpy(fg.bufs, bufs, sizeof(bufs));
    }
    CUDA_TRY(cudaMemcpyAsync(e->g_ids.p, d_ids, rows * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(e->g_lens.p, d_lens, (size_t)batch * 4, cudaMemcpyDeviceToDevice, s));
    CUDA_TRY(cudaGraphLaunch(fg.exec, s));
    CUDA_TRY(cudaMemcpyAsync(d_out, e->g_out.p, (size_t)batch * H * 4, cudaMemcpyDeviceToDevice, s));
    if (!captured_now) {
        e->prof.gemm_launches += 4 * e->n_layers;
        e->prof.other_launches += 3 * e->n_layers + 2;
    }
    if (sync) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

// Caller holds e->mu and has selected the device.
static int minilm_embed_body_locked(const fsgpu_minilm* e, const int32_t* d_ids, const int32_t* d_lens, uint32_t batch,
                                    uint32_t max_len, float* d_out, cudaStream_t s, bool sync);

// Every forward goes through here (e->mu held).  The encoder owns ONE set of activation buffers, so a call on stream B
// must not start while the kernels of the previous call are still running on stream A: B waits on the event recorded
// at the end of that call (a no-op for back-to-back calls on one stream).  A caller that is CAPTURING its stream gets no
// such ordering (the replays happen at times the library does not see): replays of one encoder must be ordered by the
// caller — include/fsgpu.h.
static int minilm_embed_locked(const fsgpu_minilm* e, const int32_t* d_ids, const int32_t* d_lens, uint32_t batch,
                               uint32_t max_len, float* d_out, cudaStream_t s, bool sync) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    bool capturing = false;
    if (cudaStreamIsCapturing(s, &st) != cudaSuccess) cudaGetLastError();
    else capturing = st != cudaStreamCaptureStatusNone;
    if (!capturing && e->have_last && e->last_stream != s) CUDA_TRY(cudaStreamWaitEvent(s, e->ev_last, 0));
    int rc = minilm_embed_body_locked(e, d_ids, d_lens, batch, max_len, d_out, s, sync);
    if (rc || capturing) return rc;
    if (!e->ev_last) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_last, cudaEventDisableTiming));
    CUDA_TRY(cudaEventRecord(e->ev_last, s));
    e->last_stream = s;
    e->have_last = true;
    return FSGPU_OK;
}

static int minilm_embed_body_locked(const fsgpu_minilm* e, const int32_t* d_ids, const int32_t* d_lens, uint32_t batch,
                                    uint32_t max_len, float* d_out, cudaStream_t s, bool sync) {
    if (max_len == 0 || max_len > e->max_pos)
        return fail(FSGPU_ERR_EMBEDDING_FAILED, "minilm: max_len %u outside 1..%u", max_len, e->max_pos);
    const uint64_t rows = (uint64_t)batch * max_len;
    if (rows > 0x7FFFFF00ull) return fail(FSGPU_ERR_INVALID_CONFIG, "minilm: batch * max_len too large");
    // FSGPU_MINILM_PRODUCTS: 0 (default) = the f16 form for query lengths <= 32 (minilm_fast_kernels.cuh),
    // 3 = split-f16 operands, three products (an f32 forward to ~2e-6), 1 = the split form's hi halves only
    const int products_env = env_int("FSGPU_MINILM_PRODUCTS", 0);
    if (products_env == 0 && max_len <= 32 && e->hidden == kHidden && e->inter % 128 == 0)
        return minilm_embed_fast_locked(e, d_ids, d_lens, batch, max_len, d_out, s, sync);
    const uint32_t m = (uint32_t)rows, H = kHidden, I = e->inter;
    if (rows != e->act_rows) {  // descriptors carry the row count: (re)build on a shape change
        CUDA_TRY(cudaStreamSynchronize(s));
        int rc = minilm_reserve_act(e, &e->act_h, rows, H);
        if (!rc) rc = minilm_reserve_act(e, &e->act_ctx, rows, H);
        if (!rc) rc = minilm_reserve_act(e, &e->act_ffn, rows, I);
        if (rc) return rc;
        e->act_rows = rows;
    }
    CUDA_TRY(e->ws_h32.reserve(rows * H * 4));
    CUDA_TRY(e->ws_pre32.reserve(rows * H * 4));
    CUDA_TRY(e->ws_qkv32.reserve(rows * 3 * H * 4));
    const uint32_t products = products_env == 1 ? 1 : 3;
    const unsigned row_blocks = (unsigned)((rows + 7) / 8);
    float* h32 = e->ws_h32.as<float>();
    float* pre32 = e->ws_pre32.as<float>();
    float* qkv32 = e->ws_qkv32.as<float>();

    minilm_embed_kernel<<<row_blocks, 256, 0, s>>>(d_ids, batch, max_len, e->vocab, e->word, e->pos, e->type0, e->emb_g,
                                                   e->emb_b, e->eps, h32, e->act_h.hi, e->act_h.lo);
    CUDA_TRY(cudaGetLastError());
    const size_t att_smem = ((size_t)2 * max_len * 33 + (size_t)4 * max_len) * 4;
    CUDA_TRY(cudaFuncSetAttribute(minilm_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
    for (uint32_t li = 0; li < e->n_layers; ++li) {
        const MiniLmLayer& L = e->layers[li];
        int rc = minilm_gemm(e, e->act_h, L.qkv, m, L.qkv_b, nullptr, qkv32, nullptr, nullptr, 0, products, s);
        if (rc) return rc;
        if (max_len <= 32)
            minilm_attention_short_kernel<<<(batch * kHeads + 3) / 4, 128, 0, s>>>(qkv32, d_lens, batch, max_len,
                                                                                   e->act_ctx.hi, e->act_ctx.lo);
        else
            minilm_attention_kernel<<<batch * kHeads, 128, att_smem, s>>>(qkv32, d_lens, max_len, e->act_ctx.hi, e->act_ctx.lo);
        CUDA_TRY(cudaGetLastError());
        rc = minilm_gemm(e, e->act_ctx, L.attn_out, m, L.attn_out_b, h32, pre32, nullptr, nullptr, 0, products, s);
        if (rc) return rc;
        minilm_layernorm_kernel<<<row_blocks, 256, 0, s>>>(pre32, rows, L.attn_ln_g, L.attn_ln_b, e->eps, h32, e->act_h.hi,
                                                           e->act_h.lo);
        CUDA_TRY(cudaGetLastError());
        rc = minilm_gemm(e, e->act_h, L.ffn_in, m, L.ffn_in_b, nullptr, nullptr, e->act_ffn.hi, e->act_ffn.lo, 1, products, s);
        if (rc) return rc;
        rc = minilm_gemm(e, e->act_ffn, L.ffn_out, m, L.ffn_out_b, h32, pre32, nullptr, nullptr, 0, products, s);
        if (rc) return rc;
        minilm_layernorm_kernel<<<row_blocks, 256, 0, s>>>(pre32, rows, L.ffn_ln_g, L.ffn_ln_b, e->eps, h32, e->act_h.hi,
                                                           e->act_h.lo);
        CUDA_TRY(cudaGetLastError());
        e->prof.other_launches += 3;
    }
    minilm_pool_kernel<<<batch, 128, 0, s>>>(h32, d_lens, max_len, d_out);
    CUDA_TRY(cudaGetLastError());
    e->prof.other_launches += 2;
    if (sync) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_minilm_embed_device(const fsgpu_minilm* e, const int32_t* d_ids, const int32_t* d_lens,
                                         uint32_t batch, uint32_t max_len, float* d_out, void* stream) {
    if (!e) return fail(FSGPU_ERR_INVALID_CONFIG, "encoder is NULL");
    if (batch == 0) return FSGPU_OK;
    if (!d_ids || !d_lens || !d_out) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(e->mu);
    DeviceGuard g(e->device);
    return minilm_embed_locked(e, d_ids, d_lens, batch, max_len, d_out, stream ? (cudaStream_t)stream : e->stream,
                               stream == nullptr);
}

extern "C" int fsgpu_minilm_embed(const fsgpu_minilm* e, const int32_t* ids, const int32_t* lens, uint32_t batch,
                                  uint32_t max_len, float* out) {
    if (!e) return fail(FSGPU_ERR_INVALID_CONFIG, "encoder is NULL");
    if (batch == 0) return FSGPU_OK;
    if (!ids || !lens || !out) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(e->mu);
    DeviceGuard g(e->device);
    CUDA_TRY(e->ws_ids.reserve((size_t)batch * max_len * 4));
    CUDA_TRY(e->ws_lens.reserve((size_t)batch * 4));
    CUDA_TRY(e->ws_out.reserve((size_t)batch * kHidden * 4));
    CUDA_TRY(cudaMemcpyAsync(e->ws_ids.p, ids, (size_t)batch * max_len * 4, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY(cudaMemcpyAsync(e->ws_lens.p, lens, (size_t)batch * 4, cudaMemcpyHostToDevice, e->stream));
    int rc = minilm_embed_locked(e, e->ws_ids.as<int32_t>(), e->ws_lens.as<int32_t>(), batch, max_len,
                                 e->ws_out.as<float>(), e->stream, false);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, e->ws_out.p, (size_t)batch * kHidden * 4, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(cudaStreamSynchronize(e->stream));
    return FSGPU_OK;
}

extern "C" int fsgpu_minilm_profile_enable(fsgpu_minilm* e, int on) {
    if (!e) return fail(FSGPU_ERR_INVALID_CONFIG, "encoder is NULL");
    std::lock_guard<std::mutex> lock(e->mu);
    e->profiling = on != 0;
    return FSGPU_OK;
}

extern "C" int fsgpu_minilm_profile_read(fsgpu_minilm* e, fsgpu_minilm_profile* out, int reset) {
    if (!e || !out) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    std::lock_guard<std::mutex> lock(e->mu);
    DeviceGuard g(e->device);
    for (auto& ev : e->ev_pending) {
        CUDA_TRY(cudaEventSynchronize(ev.second));
        float ms = 0.0f;
        CUDA_TRY(cudaEventElapsedTime(&ms, ev.first, ev.second));
        e->prof.gemm_ms += ms;
        e->ev_free.push_back(ev);
    }
    e->ev_pending.clear();
    *out = e->prof;
    if (reset) e->prof = fsgpu_minilm_profile{};
    return FSGPU_OK;
}

// ─── safetensors loader (fsgpu_minilm_load) ───────────────────────────────────────
// `model.safetensors` of sentence-transformers/all-MiniLM-L6-v2 (sha256 53aa5117...d9db, 90 868 376 B,
// crates/frankensearch-embed/src/model_manifest.rs:343-349): an 8-byte little-endian header length, a
// JSON object {tensor name: {"dtype", "shape", "data_offsets": [begin, end]}, "__metadata__": {...}},
// then the tensor bytes.  Names are those of a Hugging Face BertModel, optionally prefixed with "bert."
// or "0.auto_model."; F32, F16 and BF16 tensors are accepted (widened on the host).
namespace {
struct StTensor {
    std::string dtype;
    std::vector<uint64_t> shape;
    uint64_t begin = 0, end = 0;
};

// A minimal parser for the flat two-level JSON of a safetensors header.
struct StParser {
    const char* p;
    const char* end;
    bool ok = true;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p;
    }
    bool eat(char c) {
        ws();
        if (p < end && *p == c) {
            ++p;
            return true;
        }
        return false;
    }
    std::string str() {
        ws();
        std::string s;
        if (p >= end || *p != '"') {
            ok = false;
            return s;
        }
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) ++p;  // escapes do not occur in tensor names; keep the next byte
            s.push_back(*p++);
        }
        if (p >= end) ok = false;
        ++p;
        return s;
    }
    uint64_t num() {
        ws();
        uint64_t v = 0;
        bool any = false;
        while (p < end && *p >= '0' && *p <= '9') {
            v = v * 10 + (uint64_t)(*p++ - '0');
            any = true;
        }
        if (!any) ok = false;
        return v;
    }
    void skip_value() {  // strings, numbers, nested objects / arrays (the __metadata__ entry)
        ws();
        if (p >= end) {
            ok = false;
            return;
        }
        if (*p == '"') {
            str();
        } else if (*p == '{' || *p == '[') {
            const char open = *p, close = open == '{' ? '}' : ']';
            int depth = 0;
            bool in_str = false;
            for (; p < end; ++p) {
                if (in_str) {
                    if (*p == '\\') ++p;
                    else if (*p == '"') in_str = false;
                } else if (*p == '"') in_str = true;
                else if (*p == open) ++depth;
                else if (*p == close && --depth == 0) {
                    ++p;
                    return;
                }
            }
            ok = false;
        } else {
            while (p < end && *p != ',' && *p != '}' && *p != ']') ++p;
        }
    }
};

static bool st_parse_header(const char* json, size_t len, std::vector<std::pair<std::string, StTensor>>* out) {
    StParser ps{json, json + len};
    if (!ps.eat('{')) return false;
    if (ps.eat('}')) return true;
    do {
        const std::string name = ps.str();
        if (!ps.ok || !ps.eat(':')) return false;
        if (name == "__metadata__") {
            ps.skip_value();
        } else {
            StTensor t;
            if (!ps.eat('{')) return false;
            do {
                const std::string key = ps.str();
                if (!ps.ok || !ps.eat(':')) return false;
                if (key == "dtype") {
                    t.dtype = ps.str();
                } else if (key == "shape") {
                    if (!ps.eat('[')) return false;
                    if (!ps.eat(']')) {
                        do t.shape.push_back(ps.num()); while (ps.eat(','));
                        if (!ps.eat(']')) return false;
                    }
                } else if (key == "data_offsets") {
                    if (!ps.eat('[')) return false;
                    t.begin = ps.num();
                    if (!ps.eat(',')) return false;
                    t.end = ps.num();
                    if (!ps.eat(']')) return false;
                } else {
                    ps.skip_value();
                }
            } while (ps.ok && ps.eat(','));
            if (!ps.ok || !ps.eat('}')) return false;
            out->emplace_back(name, t);
        }
    } while (ps.ok && ps.eat(','));
    return ps.ok && ps.eat('}');
}

static float st_widen16(uint16_t bits, bool bf16) {
    uint32_t u;
    if (bf16) {
        u = (uint32_t)bits << 16;
    } else {  // IEEE f16 -> f32
        const uint32_t sign = (uint32_t)(bits & 0x8000u) << 16, exp = (bits >> 10) & 0x1Fu, man = bits & 0x3FFu;
        if (exp == 0) {
            if (man == 0) u = sign;
            else {
                int e = -1;
                uint32_t m = man;
                do { ++e; m <<= 1; } while (!(m & 0x400u));
                u = sign | ((uint32_t)(127 - 15 - e) << 23) | ((m & 0x3FFu) << 13);
            }
        } else if (exp == 31) {
            u = sign | 0x7F800000u | (man << 13);
        } else {
            u = sign | ((exp + 112u) << 23) | (man << 13);
        }
    }
    float f;
    memcpy(&f, &u, 4);
    return f;
}
}  // namespace

extern "C" int fsgpu_minilm_load(const char* safetensors_path, int device, fsgpu_minilm** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    if (!safetensors_path) return fail(FSGPU_ERR_INVALID_CONFIG, "path is NULL");
    FILE* f = fopen(safetensors_path, "rb");
    if (!f) return fail(FSGPU_ERR_IO, "cannot open %s", safetensors_path);
    fseeko(f, 0, SEEK_END);
    const uint64_t fsize = (uint64_t)ftello(f);
    fseeko(f, 0, SEEK_SET);
    uint64_t hlen = 0;
    if (fsize < 8 || fread(&hlen, 1, 8, f) != 8 || hlen == 0 || hlen > fsize - 8 || hlen > (100u << 20)) {
        fclose(f);
        return fail(FSGPU_ERR_EMBEDDING_FAILED, "%s: not a safetensors file (bad header length)", safetensors_path);
    }
    std::vector<char> header(hlen);
    std::vector<uint8_t> data(fsize - 8 - hlen);
    const bool read_ok = fread(header.data(), 1, hlen, f) == hlen && (data.empty() || fread(data.data(), 1, data.size(), f) == data.size());
    fclose(f);
    if (!read_ok) return fail(FSGPU_ERR_IO, "%s: short read", safetensors_path);
    std::vector<std::pair<std::string, StTensor>> tensors;
    if (!st_parse_header(header.data(), hlen, &tensors))
        return fail(FSGPU_ERR_EMBEDDING_FAILED, "%s: malformed safetensors header", safetensors_path);

    std::vector<std::vector<float>> keep;  // widened / concatenated tensors the weight struct points into
    std::string missing;
    auto find = [&](const std::string& key, std::vector<uint64_t>* shape) -> const float* {
        for (const char* prefix : {"", "bert.", "0.auto_model."}) {
            const std::string full = std::string(prefix) + key;
            for (auto& kv : tensors) {
                if (kv.first != full) continue;
                const StTensor& t = kv.second;
                uint64_t count = 1;
                for (uint64_t d : t.shape) count *= d;
                const uint64_t esz = t.dtype == "F32" ? 4 : (t.dtype == "F16" || t.dtype == "BF16") ? 2 : 0;
                if (!esz || t.end < t.begin || t.end > data.size() || t.end - t.begin != count * esz) {
                    missing = full + " (unsupported dtype or bad offsets)";
                    return nullptr;
                }
                if (shape) *shape = t.shape;
                if (esz == 4 && (reinterpret_cast<uintptr_t>(data.data() + t.begin) & 3u) == 0)
                    return reinterpret_cast<const float*>(data.data() + t.begin);
                keep.emplace_back(count);
                if (esz == 4) {
                    memcpy(keep.back().data(), data.data() + t.begin, count * 4);
                } else {
                    const bool bf = t.dtype == "BF16";
                    for (uint64_t i = 0; i < count; ++i) {
                        uint16_t b;
                        memcpy(&b, data.data() + t.begin + 2 * i, 2);
                        keep.back()[i] = st_widen16(b, bf);
                    }
                }
                return keep.back().data();
            }
        }
        if (missing.empty()) missing = key;
        return nullptr;
    };
    std::vector<uint64_t> shp;
    fsgpu_minilm_weights w{};
    w.word_emb = find("embeddings.word_embeddings.weight", &shp);
    if (!w.word_emb || shp.size() != 2) return fail(FSGPU_ERR_EMBEDDING_FAILED, "%s: tensor %s is missing", safetensors_path, missing.c_str());
    w.vocab_size = (uint32_t)shp[0];
    w.hidden = (uint32_t)shp[1];
    w.pos_emb = find("embeddings.position_embeddings.weight", &shp);
    if (w.pos_emb && shp.size() == 2) w.max_positions = (uint32_t)shp[0];
    w.type_emb = find("embeddings.token_type_embeddings.weight", nullptr);
    w.emb_ln_g = find("embeddings.LayerNorm.weight", nullptr);
    w.emb_ln_b = find("embeddings.LayerNorm.bias", nullptr);
    uint32_t n_layers = 0;
    for (;; ++n_layers) {
        const std::string key = "encoder.layer." + std::to_string(n_layers) + ".attention.self.query.weight";
        bool have = false;
        for (auto& kv : tensors)
            for (const char* prefix : {"", "bert.", "0.auto_model."}) have = have || kv.first == std::string(prefix) + key;
        if (!have) break;
    }
    std::vector<fsgpu_minilm_layer_weights> layers(n_layers);
    const uint64_t H = w.hidden;
    for (uint32_t i = 0; i < n_layers && missing.empty(); ++i) {
        const std::string p = "encoder.layer." + std::to_string(i) + ".";
        fsgpu_minilm_layer_weights& L = layers[i];
        const float* qkv_w[3];
        const float* qkv_b[3];
        const char* names[3] = {"query", "key", "value"};
        for (int j = 0; j < 3; ++j) {
            qkv_w[j] = find(p + "attention.self." + names[j] + ".weight", nullptr);
            qkv_b[j] = find(p + "attention.self." + names[j] + ".bias", nullptr);
        }
        if (!missing.empty()) break;
        keep.emplace_back(3 * H * H);  // query | key | value rows (fsgpu_minilm_layer_weights.qkv_w)
        float* cw = keep.back().data();
        keep.emplace_back(3 * H);
        float* cb = keep.back().data();
        for (int j = 0; j < 3; ++j) {
            memcpy(cw + j * H * H, qkv_w[j], H * H * 4);
            memcpy(cb + j * H, qkv_b[j], H * 4);
        }
        L.qkv_w = cw;
        L.qkv_b = cb;
        L.attn_out_w = find(p + "attention.output.dense.weight", nullptr);
        L.attn_out_b = find(p + "attention.output.dense.bias", nullptr);
        L.attn_ln_g = find(p + "attention.output.LayerNorm.weight", nullptr);
        L.attn_ln_b = find(p + "attention.output.LayerNorm.bias", nullptr);
        L.ffn_in_w = find(p + "intermediate.dense.weight", &shp);
        if (L.ffn_in_w && shp.size() == 2) w.intermediate = (uint32_t)shp[0];
        L.ffn_in_b = find(p + "intermediate.dense.bias", nullptr);
        L.ffn_out_w = find(p + "output.dense.weight", nullptr);
        L.ffn_out_b = find(p + "output.dense.bias", nullptr);
        L.ffn_ln_g = find(p + "output.LayerNorm.weight", nullptr);
        L.ffn_ln_b = find(p + "output.LayerNorm.bias", nullptr);
    }
    if (!missing.empty() || n_layers == 0)
        return fail(FSGPU_ERR_EMBEDDING_FAILED, "%s: tensor %s is missing", safetensors_path,
                    missing.empty() ? "encoder.layer.0.attention.self.query.weight" : missing.c_str());
    w.n_layers = n_layers;
    w.heads = kHeads;       // config.json num_attention_heads of all-MiniLM-L6-v2 (native.rs:36-45)
    w.ln_eps = 1e-12f;      // layer_norm_eps
    w.layers = layers.data();
    return fsgpu_minilm_create(&w, device, out);
}
