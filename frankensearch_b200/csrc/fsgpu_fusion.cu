// fsgpu_fusion.cu — C ABI of the fusion stage (RRF, two-tier blend) and the potion static embedder
// over fusion_kernels.cuh.  No CPU compute path: every entry point launches kernels or fails.
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include "fsgpu.h"
#include "fsgpu_common.cuh"
#include "fsgpu_host.cuh"
#include "fusion_kernels.cuh"

using namespace fsgpu;

// ─── fusion ─────────────────────────────────────────────────────────────────────────────────
static void sanitize_rrf(const fsgpu_rrf_config* c, RrfArgs* a) {
    double k = c ? c->k : 60.0, wl = c ? c->lexical_weight : 1.0, ws = c ? c->semantic_weight : 1.0;
    if (!(std::isfinite(k) && k >= 0.0)) k = 60.0;        // rrf.rs:124-130
    if (!(std::isfinite(wl) && wl > 0.0)) wl = 1.0;       // rrf.rs:92-98
    if (!(std::isfinite(ws) && ws > 0.0)) ws = 1.0;
    a->k = k;
    a->w_lex = wl;
    a->w_sem = ws;
    a->tiebreak = c ? c->tiebreak : 0;
}

static int launch_rrf(RrfArgs& a, uint32_t batch, cudaStream_t s) {
    const uint32_t m = a.n_lex_max + a.n_sem_max;
    if (m > kFusionMaxEntries)
        return fail(FSGPU_ERR_INVALID_CONFIG, "rrf: %u candidates exceed the device window of %u", m,
                    kFusionMaxEntries);
    const size_t smem = (size_t)host_next_pow2(std::max(m, 1u)) * 20 + 16;
    if (m > kFusionWideFrom) {
        CUDA_TRY(cudaFuncSetAttribute(rrf_fuse_kernel<kFusionWideThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rrf_fuse_kernel<kFusionWideThreads><<<batch, kFusionWideThreads, smem, s>>>(a);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(rrf_fuse_kernel<kFusionThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        rrf_fuse_kernel<kFusionThreads><<<batch, kFusionThreads, smem, s>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return FSGPU_OK;
}

extern "C" int fsgpu_rrf_fuse_device(int device, const fsgpu_rrf_config* config, uint32_t batch,
                                     const uint64_t* d_lex_ids, const float* d_lex_scores,
                                     const uint32_t* d_lex_tie, const uint32_t* d_lex_counts,
                                     uint32_t n_lex_max, const fsgpu_hit* d_sem_hits,
                                     const uint32_t* d_sem_tie, const uint32_t* d_sem_counts,
                                     uint32_t n_sem_max, uint32_t limit, uint32_t offset,
                                     fsgpu_fused_hit* d_out, uint32_t* d_out_counts, void* stream) {
    if (batch == 0) return FSGPU_OK;
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    if (limit == 0) {
        if (d_out_counts) CUDA_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)batch * 4, s));
        return FSGPU_OK;
    }
    RrfArgs a{};
    sanitize_rrf(config, &a);
    a.lex_ids = d_lex_ids;
    a.lex_scores = d_lex_scores;
    a.lex_tie = d_lex_tie;
    a.lex_counts = d_lex_counts;
    a.n_lex_max = n_lex_max;
    a.sem_hits = d_sem_hits;
    a.sem_tie = d_sem_tie;
    a.sem_counts = d_sem_counts;
    a.n_sem_max = n_sem_max;
    a.limit = limit;
    a.offset = offset;
    a.out = d_out;
    a.out_counts = d_out_counts;
    int rc = launch_rrf(a, batch, s);
    if (rc) return rc;
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_rrf_fuse(int device, const fsgpu_rrf_config* config, uint32_t batch,
                              const uint64_t* lex_ids, const float* lex_scores, const uint32_t* lex_tie,
                              const uint32_t* lex_counts, uint32_t n_lex_max, const uint32_t* sem_rows,
                              const float* sem_scores, const uint32_t* sem_tie, const uint32_t* sem_counts,
                              uint32_t n_sem_max, uint32_t limit, uint32_t offset, fsgpu_fused_hit* out,
                              uint32_t* out_counts) {
    if (batch == 0) return FSGPU_OK;
    if (!out_counts) return fail(FSGPU_ERR_INVALID_CONFIG, "out_counts is NULL");
    int ndev = 0;
    int rc = fsgpu_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(FSGPU_ERR_INVALID_CONFIG, "device %d not present", device);
    if (limit == 0) {  // rrf.rs:1172-1183 window == 0
        memset(out_counts, 0, (size_t)batch * 4);
        return FSGPU_OK;
    }
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    const size_t nl = (size_t)batch * n_lex_max, ns = (size_t)batch * n_sem_max;
    for (size_t i = 0; i < nl; ++i)
        if (lex_ids[i] >> 40) return fail(FSGPU_ERR_INVALID_CONFIG, "rrf: lexical id exceeds 40 bits");
    for (const uint32_t* t : {lex_tie, sem_tie})
        if (t)
            for (size_t i = 0; i < (t == lex_tie ? nl : ns); ++i)
                if (t[i] >> kTieBits)
                    return fail(FSGPU_ERR_INVALID_CONFIG, "rrf: tie rank exceeds %u bits", kTieBits);
    DeviceGuard g(device);
    Staging st;
    RrfArgs a{};
    sanitize_rrf(config, &a);
    uint64_t* d_lex_ids; float* d_lex_scores; uint32_t *d_lex_tie, *d_lex_counts;
    uint32_t *d_sem_rows, *d_sem_tie, *d_sem_counts; float* d_sem_scores;
    fsgpu_fused_hit* d_out; uint32_t* d_out_counts;
    CUDA_TRY(st.up(lex_ids, nl, &d_lex_ids));
    CUDA_TRY(st.up(lex_scores, nl, &d_lex_scores));
    CUDA_TRY(st.up(lex_tie, nl, &d_lex_tie));
    CUDA_TRY(st.up(lex_counts, batch, &d_lex_counts));
    CUDA_TRY(st.up(sem_rows, ns, &d_sem_rows));
    CUDA_TRY(st.up(sem_scores, ns, &d_sem_scores));
    CUDA_TRY(st.up(sem_tie, ns, &d_sem_tie));
    CUDA_TRY(st.up(sem_counts, batch, &d_sem_counts));
    CUDA_TRY(st.alloc((size_t)batch * limit, &d_out));
    CUDA_TRY(st.alloc(batch, &d_out_counts));
    a.lex_ids = d_lex_ids; a.lex_scores = d_lex_scores; a.lex_tie = d_lex_tie; a.lex_counts = d_lex_counts;
    a.n_lex_max = n_lex_max;
    a.sem_rows = d_sem_rows; a.sem_scores = d_sem_scores; a.sem_tie = d_sem_tie; a.sem_counts = d_sem_counts;
    a.n_sem_max = n_sem_max;
    a.limit = limit; a.offset = offset; a.out = d_out; a.out_counts = d_out_counts;
    rc = launch_rrf(a, batch, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(out, d_out, (size_t)batch * limit * sizeof(fsgpu_fused_hit), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(out_counts, d_out_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost));
    return FSGPU_OK;
}

static float sanitize_alpha(float blend_factor) {  // blend.rs:518-524
    float alpha = blend_factor;
    if (!std::isfinite(alpha)) alpha = 0.7f;
    return std::min(1.0f, std::max(0.0f, alpha));
}

static int launch_blend(const BlendArgs& a, uint32_t batch, cudaStream_t s) {
    const uint32_t m = a.n_fast_max + (a.union_form ? a.n_quality_max : 0);
    if (m > kFusionMaxEntries)
        return fail(FSGPU_ERR_INVALID_CONFIG, "blend: %u hits exceed the device window of %u", m, kFusionMaxEntries);
    const size_t smem = (size_t)host_next_pow2(std::max(m, 1u)) * 20 + 16;
    if (m > kFusionWideFrom) {
        CUDA_TRY(cudaFuncSetAttribute(blend_two_tier_kernel<kFusionWideThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        blend_two_tier_kernel<kFusionWideThreads><<<batch, kFusionWideThreads, smem, s>>>(a);
    } else {
        CUDA_TRY(cudaFuncSetAttribute(blend_two_tier_kernel<kFusionThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        blend_two_tier_kernel<kFusionThreads><<<batch, kFusionThreads, smem, s>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    return FSGPU_OK;
}

extern "C" int fsgpu_blend_two_tier_device(int device, float blend_factor, uint32_t batch,
                                           const fsgpu_hit* d_fast_hits, const uint32_t* d_fast_tie,
                                           const uint32_t* d_fast_counts, uint32_t n_fast_max,
                                           const fsgpu_hit* d_quality_hits, const float* d_quality_scores,
                                           const uint8_t* d_quality_present, const uint32_t* d_quality_tie,
                                           const uint32_t* d_quality_counts, uint32_t n_quality_max,
                                           fsgpu_hit* d_out, uint32_t* d_out_counts, void* stream) {
    if (batch == 0) return FSGPU_OK;
    if (!d_out || !d_out_counts) return fail(FSGPU_ERR_INVALID_CONFIG, "blend: output is NULL");
    if (n_fast_max && !d_fast_hits) return fail(FSGPU_ERR_INVALID_CONFIG, "blend: fast hits is NULL");
    const bool union_form = d_quality_hits != nullptr;
    if (!union_form && n_fast_max && !d_quality_scores)
        return fail(FSGPU_ERR_INVALID_CONFIG, "blend: aligned form needs quality scores (or a retrieved quality list)");
    DeviceGuard g(device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n_fast_max + (union_form ? n_quality_max : 0) == 0) {
        CUDA_TRY(cudaMemsetAsync(d_out_counts, 0, (size_t)batch * 4, s));
        return FSGPU_OK;
    }
    BlendArgs a{};
    a.alpha = sanitize_alpha(blend_factor);
    a.fast_hits = d_fast_hits;
    a.fast_tie = d_fast_tie;
    a.fast_counts = d_fast_counts;
    a.n_fast_max = n_fast_max;
    a.union_form = union_form ? 1u : 0u;
    a.quality_hits = d_quality_hits;
    a.quality_scores = d_quality_scores;
    a.quality_present = d_quality_present;
    a.quality_tie = d_quality_tie;
    a.quality_counts = d_quality_counts;
    a.n_quality_max = union_form ? n_quality_max : n_fast_max;
    a.out = d_out;
    a.out_stride = n_fast_max + (union_form ? n_quality_max : 0);
    a.out_counts = d_out_counts;
    int rc = launch_blend(a, batch, s);
    if (rc) return rc;
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_blend_two_tier(int device, float blend_factor, const uint32_t* fast_rows,
                                    const float* fast_scores, const uint32_t* fast_tie, uint32_t n_fast,
                                    const uint32_t* quality_rows, const float* quality_scores,
                                    const uint8_t* quality_present, const uint32_t* quality_tie,
                                    uint32_t n_quality, fsgpu_hit* out, uint32_t* out_count) {
    if (!out_count) return fail(FSGPU_ERR_INVALID_CONFIG, "out_count is NULL");
    int ndev = 0;
    int rc = fsgpu_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(FSGPU_ERR_INVALID_CONFIG, "device %d not present", device);
    const bool union_form = quality_rows != nullptr;
    if (!union_form && n_quality != n_fast && n_quality != 0)
        return fail(FSGPU_ERR_INVALID_CONFIG, "aligned blend needs one quality slot per fast hit");
    const uint32_t m = n_fast + (union_form ? n_quality : 0);
    if (m == 0) {
        *out_count = 0;
        return FSGPU_OK;
    }
    if (m > kFusionMaxEntries)
        return fail(FSGPU_ERR_INVALID_CONFIG, "blend: %u hits exceed the device window of %u", m, kFusionMaxEntries);
    DeviceGuard g(device);
    Staging st;
    BlendArgs a{};
    a.alpha = sanitize_alpha(blend_factor);
    uint32_t *d_fr, *d_ft, *d_qr, *d_qt; float *d_fs, *d_qs; uint8_t* d_qp; fsgpu_hit* d_out; uint32_t* d_cnt;
    std::vector<uint8_t> none;
    const uint32_t nq_eff = union_form ? n_quality : n_fast;
    if (!union_form && n_quality == 0) {  // aligned form with no quality scores at all
        none.assign(n_fast, 0);
        quality_present = none.data();
    }
    std::vector<float> zero_q;
    if (!quality_scores) {
        zero_q.assign(nq_eff, 0.0f);
        quality_scores = zero_q.data();
    }
    CUDA_TRY(st.up(fast_rows, n_fast, &d_fr));
    CUDA_TRY(st.up(fast_scores, n_fast, &d_fs));
    CUDA_TRY(st.up(fast_tie, n_fast, &d_ft));
    CUDA_TRY(st.up(quality_rows, union_form ? n_quality : 0, &d_qr));
    CUDA_TRY(st.up(quality_scores, nq_eff, &d_qs));
    CUDA_TRY(st.up(quality_present, union_form ? 0 : n_fast, &d_qp));
    CUDA_TRY(st.up(quality_tie, union_form ? n_quality : 0, &d_qt));
    CUDA_TRY(st.alloc(m, &d_out));
    CUDA_TRY(st.alloc(1, &d_cnt));
    a.fast_rows = d_fr; a.fast_scores = d_fs; a.fast_tie = d_ft; a.n_fast_max = n_fast;
    a.union_form = union_form ? 1u : 0u;
    a.quality_rows = d_qr; a.quality_scores = d_qs; a.quality_present = d_qp; a.quality_tie = d_qt;
    a.n_quality_max = nq_eff;
    a.out = d_out; a.out_stride = m; a.out_counts = d_cnt;
    rc = launch_blend(a, 1, nullptr);
    if (rc) return rc;
    uint32_t cnt = 0;
    CUDA_TRY(cudaMemcpy(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(out, d_out, (size_t)cnt * sizeof(fsgpu_hit), cudaMemcpyDeviceToHost));
    *out_count = cnt;
    return FSGPU_OK;
}

// ─── potion ─────────────────────────────────────────────────────────────────────────────────
struct fsgpu_potion {
    int device = 0;
    uint64_t vocab = 0;
    uint32_t dim = 0;
    float* d_table = nullptr;
    cudaStream_t stream = nullptr;
    mutable std::mutex mu;
};

extern "C" void fsgpu_potion_destroy(fsgpu_potion* e) {
    if (!e) return;
    {
        DeviceGuard g(e->device);
        if (e->stream) {
            cudaStreamSynchronize(e->stream);
            cudaStreamDestroy(e->stream);
        }
        if (e->d_table) cudaFree(e->d_table);
    }
    delete e;
}

extern "C" int fsgpu_potion_create(const float* table, uint64_t vocab, uint32_t dim, int device,
                                   fsgpu_potion** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    if (!table || vocab == 0 || dim == 0)  // validate_model2vec_accumulation_shape (embed/src/simd.rs:118+)
        return fail(FSGPU_ERR_INVALID_CONFIG, "potion table must be a non-empty [vocab, dim] matrix");
    int ndev = 0;
    int rc = fsgpu_device_count(&ndev);
    if (rc) return rc;
    if (device < 0 || device >= ndev) return fail(FSGPU_ERR_INVALID_CONFIG, "device %d not present", device);
    fsgpu_potion* e = new fsgpu_potion();
    e->device = device;
    e->vocab = vocab;
    e->dim = dim;
    DeviceGuard g(device);
    cudaError_t err = cudaMalloc(&e->d_table, vocab * dim * 4);
    if (err == cudaSuccess) err = h2d_complete(e->d_table, table, vocab * dim * 4);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
    if (err != cudaSuccess) {
        fsgpu_potion_destroy(e);
        return fail(FSGPU_ERR_SUBSYSTEM, "gpu: potion table upload failed: %s", cudaGetErrorString(err));
    }
    *out = e;
    return FSGPU_OK;
}

extern "C" int fsgpu_potion_embed_device(const fsgpu_potion* e, const uint32_t* d_ids, const uint64_t* d_offsets,
                                         uint32_t batch, float* d_out, void* stream) {
    if (!e) return fail(FSGPU_ERR_INVALID_CONFIG, "encoder is NULL");
    if (batch == 0) return FSGPU_OK;
    std::lock_guard<std::mutex> lock(e->mu);
    DeviceGuard g(e->device);
    cudaStream_t s = stream ? (cudaStream_t)stream : e->stream;
    potion_embed_kernel<<<batch, 256, (size_t)e->dim * 4, s>>>(e->d_table, e->vocab, e->dim, d_ids, d_offsets, d_out);
    CUDA_TRY(cudaGetLastError());
    if (!stream) CUDA_TRY(cudaStreamSynchronize(s));
    return FSGPU_OK;
}

extern "C" int fsgpu_potion_embed(const fsgpu_potion* e, const uint32_t* ids, const uint64_t* offsets,
                                  uint32_t batch, float* out) {
    if (!e) return fail(FSGPU_ERR_INVALID_CONFIG, "encoder is NULL");
    if (batch == 0) return FSGPU_OK;
    if (!offsets || !out) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    const uint64_t total = offsets[batch];
    DeviceGuard g(e->device);
    Staging st;
    uint32_t* d_ids; uint64_t* d_off; float* d_out;
    std::vector<uint32_t> pad(1, 0);
    CUDA_TRY(st.up(total ? ids : pad.data(), std::max<uint64_t>(1, total), &d_ids));
    CUDA_TRY(st.up(offsets, (size_t)batch + 1, &d_off));
    CUDA_TRY(st.alloc((size_t)batch * e->dim, &d_out));
    int rc = fsgpu_potion_embed_device(e, d_ids, d_off, batch, d_out, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(out, d_out, (size_t)batch * e->dim * 4, cudaMemcpyDeviceToHost));
    return FSGPU_OK;
}

