// fsgpu_sharded.cu — a row-sharded index over several GPUs of one box behind the C ABI, in ONE process
// (the Rust host has no torch.distributed and no NCCL): contiguous row shards, one shard
// per device, the same exact search on every shard, and the k-way merge of merge_partial_heaps
// (crates/frankensearch-index/src/search.rs:1704-1720) lifted across devices.
//
// There is no separate exchange step.  With peer access the search kernels of shard s store their
// top-k keys and hits STRAIGHT into slot s of the merge device's buffer over NVLink/NVSwitch (the last
// kernel of each search — the refine / merge kernel — has the remote buffer as its output pointer), so
// the "all-gather" is the kernels' own epilogue stores: 16 bytes x batch x k per shard.  One event per
// shard orders the merge kernel behind them.  Without peer access the shard writes locally and a
// cudaMemcpyPeerAsync carries the block.  Built entirely on the public entry points of fsgpu.h.
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "fsgpu.h"
#include "fsgpu_host.cuh"

namespace {

struct Shard {
    int device = 0;
    fsgpu_index* index = nullptr;
    uint64_t row_base = 0, n_rows = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf d_queries, d_local;  // d_local: results when the merge device's memory is not reachable
    bool direct = false;        // kernels store into the merge device's buffer themselves
    // worker
    std::thread thread;
    std::mutex mu;
    std::condition_variable cv;
    bool has_job = false, job_done = true, quit = false;
    int rc = 0;
    std::string error;
};

}  // namespace

struct fsgpu_sharded {
    std::vector<Shard*> shards;
    int merge_device = 0;
    uint32_t dim = 0;
    uint64_t n_rows = 0;
    bool owns_shards = false;
    cudaStream_t merge_stream = nullptr;
    DevBuf d_gather, d_out_keys, d_out_hits, d_out_counts;
    std::mutex mu;  // one search at a time
    // the job every worker runs
    const float* job_queries = nullptr;
    uint32_t job_batch = 0, job_k = 0;
};

static void worker_run(fsgpu_sharded* sh, Shard* s, size_t slot) {
    cudaSetDevice(s->device);
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(s->mu);
            s->cv.wait(lk, [&] { return s->has_job || s->quit; });
            if (s->quit) return;
            s->has_job = false;
        }
        const uint32_t batch = sh->job_batch, k = sh->job_k;
        const size_t block = (size_t)batch * k;  // entries per shard: keys [batch,k] u64 | hits [batch,k]
        int rc = 0;
        std::string err;
        auto cuda_ok = [&](cudaError_t e, const char* what) {
            if (e != cudaSuccess && rc == 0) {
                rc = FSGPU_ERR_SUBSYSTEM;
                err = std::string("gpu: ") + what + " failed on shard device: " + cudaGetErrorString(e);
            }
            return e == cudaSuccess;
        };
        if (cuda_ok(s->d_queries.reserve((size_t)batch * sh->dim * 4), "query buffer") &&
            cuda_ok(cudaMemcpyAsync(s->d_queries.p, sh->job_queries, (size_t)batch * sh->dim * 4, cudaMemcpyHostToDevice, s->stream),
                    "query upload")) {
            uint64_t* keys = sh->d_gather.as<uint64_t>() + slot * 2 * block;
            if (!s->direct) {
                if (cuda_ok(s->d_local.reserve(2 * block * 8), "result buffer")) keys = s->d_local.as<uint64_t>();
            }
            if (rc == 0) {
                fsgpu_hit* hits = reinterpret_cast<fsgpu_hit*>(keys + block);
                rc = fsgpu_search_top_k_device(s->index, s->d_queries.as<float>(), batch, k, keys, hits, nullptr, s->stream);
                if (rc) err = fsgpu_last_error();
            }
            if (rc == 0 && !s->direct)
                cuda_ok(cudaMemcpyPeerAsync(sh->d_gather.as<uint64_t>() + slot * 2 * block, sh->merge_device, s->d_local.p,
                                            s->device, 2 * block * 8, s->stream), "peer copy");
            if (rc == 0) cuda_ok(cudaEventRecord(s->done, s->stream), "event record");
        }
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->rc = rc;
            s->error = err;
            s->job_done = true;
        }
        s->cv.notify_all();
    }
}

extern "C" void fsgpu_sharded_destroy(fsgpu_sharded* sh) {
    if (!sh) return;
    for (Shard* s : sh->shards) {
        if (s->thread.joinable()) {
            {
                std::lock_guard<std::mutex> lk(s->mu);
                s->quit = true;
            }
            s->cv.notify_all();
            s->thread.join();
        }
        DeviceGuard g(s->device);
        if (s->stream) {
            cudaStreamSynchronize(s->stream);
            cudaStreamDestroy(s->stream);
        }
        if (s->done) cudaEventDestroy(s->done);
        s->d_queries.release();
        s->d_local.release();
        if (sh->owns_shards && s->index) fsgpu_index_destroy(s->index);
        delete s;
    }
    {
        DeviceGuard g(sh->merge_device);
        if (sh->merge_stream) {
            cudaStreamSynchronize(sh->merge_stream);
            cudaStreamDestroy(sh->merge_stream);
        }
        for (DevBuf* b : {&sh->d_gather, &sh->d_out_keys, &sh->d_out_hits, &sh->d_out_counts}) b->release();
    }
    delete sh;
}

static int sharded_finish(fsgpu_sharded* sh) {
    sh->merge_device = sh->shards[0]->device;
    {
        DeviceGuard g(sh->merge_device);
        CUDA_TRY(cudaStreamCreateWithFlags(&sh->merge_stream, cudaStreamNonBlocking));
    }
    for (size_t i = 0; i < sh->shards.size(); ++i) {
        Shard* s = sh->shards[i];
        DeviceGuard g(s->device);
        CUDA_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming));
        s->direct = s->device == sh->merge_device;
        if (!s->direct) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, s->device, sh->merge_device) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(sh->merge_device, 0);
                if (e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled) s->direct = true;
                cudaGetLastError();
            }
        }
        s->thread = std::thread(worker_run, sh, s, i);
    }
    return FSGPU_OK;
}

extern "C" int fsgpu_sharded_from_shards(fsgpu_index* const* shards, int n_shards, int take_ownership,
                                         fsgpu_sharded** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    if (!shards || n_shards <= 0) return fail(FSGPU_ERR_INVALID_CONFIG, "no shards");
    uint64_t next = fsgpu_index_row_base(shards[0]);
    const uint32_t dim = fsgpu_index_dim(shards[0]);
    for (int i = 0; i < n_shards; ++i) {
        if (!shards[i]) return fail(FSGPU_ERR_INVALID_CONFIG, "shard %d is NULL", i);
        if (fsgpu_index_dim(shards[i]) != dim) return fail(FSGPU_ERR_DIMENSION_MISMATCH, "expected %u, found %u", dim, fsgpu_index_dim(shards[i]));
        if (fsgpu_index_row_base(shards[i]) != next)  // contiguity keeps global_row = base + local and the lower-row tie-break
            return fail(FSGPU_ERR_INVALID_CONFIG, "shard %d starts at row %llu, expected %llu (shards must be contiguous row ranges in order)",
                        i, (unsigned long long)fsgpu_index_row_base(shards[i]), (unsigned long long)next);
        next += fsgpu_index_rows(shards[i]);
    }
    fsgpu_sharded* sh = new fsgpu_sharded();
    sh->dim = dim;
    sh->n_rows = next - fsgpu_index_row_base(shards[0]);
    sh->owns_shards = take_ownership != 0;
    for (int i = 0; i < n_shards; ++i) {
        Shard* s = new Shard();
        s->index = shards[i];
        s->device = fsgpu_index_device(shards[i]);
        s->row_base = fsgpu_index_row_base(shards[i]);
        s->n_rows = fsgpu_index_rows(shards[i]);
        sh->shards.push_back(s);
    }
    const int rc = sharded_finish(sh);
    if (rc) {
        sh->owns_shards = false;
        fsgpu_sharded_destroy(sh);
        return rc;
    }
    *out = sh;
    return FSGPU_OK;
}

extern "C" int fsgpu_sharded_create_f16(const uint16_t* slab, uint64_t n_rows, uint32_t dim, const uint8_t* tombstones,
                                        const int* devices, int n_devices, const fsgpu_index_options* opts,
                                        fsgpu_sharded** out) {
    if (!out) return fail(FSGPU_ERR_INVALID_CONFIG, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices <= 0) return fail(FSGPU_ERR_INVALID_CONFIG, "no devices");
    if (n_rows && !slab) return fail(FSGPU_ERR_INVALID_CONFIG, "slab is NULL");
    fsgpu_index_options o;
    if (opts) o = *opts; else fsgpu_index_options_default(&o);
    if (o.slab_is_device) return fail(FSGPU_ERR_INVALID_CONFIG, "fsgpu_sharded_create_f16 takes a host slab");
    std::vector<fsgpu_index*> shards;
    const uint64_t base0 = o.row_base;
    for (int i = 0; i < n_devices; ++i) {
        // contiguous ranges [i*N/G, (i+1)*N/G); a tombstone bitmap is re-packed per shard
        const uint64_t lo = (uint64_t)i * n_rows / n_devices, hi = (uint64_t)(i + 1) * n_rows / n_devices;
        std::vector<uint8_t> tomb;
        if (tombstones) {
            tomb.assign((hi - lo + 7) / 8, 0);
            for (uint64_t r = lo; r < hi; ++r)
                if ((tombstones[r >> 3] >> (r & 7)) & 1) tomb[(r - lo) >> 3] |= (uint8_t)(1u << ((r - lo) & 7));
        }
        fsgpu_index_options oi = o;
        oi.device = devices[i];
        oi.row_base = base0 + lo;
        fsgpu_index* ix = nullptr;
        const int rc = fsgpu_index_create_f16(slab + lo * dim, hi - lo, dim, tombstones ? tomb.data() : nullptr, &oi, &ix);
        if (rc) {
            for (fsgpu_index* p : shards) fsgpu_index_destroy(p);
            return rc;
        }
        shards.push_back(ix);
    }
    const int rc = fsgpu_sharded_from_shards(shards.data(), n_devices, 1, out);
    if (rc)
        for (fsgpu_index* p : shards) fsgpu_index_destroy(p);
    return rc;
}

extern "C" int fsgpu_sharded_shard_count(const fsgpu_sharded* sh) { return sh ? (int)sh->shards.size() : 0; }
extern "C" uint64_t fsgpu_sharded_rows(const fsgpu_sharded* sh) { return sh ? sh->n_rows : 0; }
extern "C" fsgpu_index* fsgpu_sharded_shard(const fsgpu_sharded* sh, int i) {
    return (sh && i >= 0 && (size_t)i < sh->shards.size()) ? sh->shards[i]->index : nullptr;
}
extern "C" int fsgpu_sharded_is_direct(const fsgpu_sharded* sh, int i) {
    return (sh && i >= 0 && (size_t)i < sh->shards.size() && sh->shards[i]->direct) ? 1 : 0;
}

extern "C" int fsgpu_sharded_search_top_k(fsgpu_sharded* sh, const float* queries, uint32_t batch, uint32_t k, uint32_t dim,
                                          fsgpu_hit* out, uint32_t* out_counts) {
    if (!sh) return fail(FSGPU_ERR_INVALID_CONFIG, "index is NULL");
    if (dim != sh->dim) return fail(FSGPU_ERR_DIMENSION_MISMATCH, "expected %u, found %u", sh->dim, dim);
    if (batch == 0) return FSGPU_OK;
    if (!queries || !out_counts || (k && !out)) return fail(FSGPU_ERR_INVALID_CONFIG, "NULL argument");
    if (k == 0 || sh->n_rows == 0) {  // search.rs:438-440
        memset(out_counts, 0, (size_t)batch * 4);
        return FSGPU_OK;
    }
    std::lock_guard<std::mutex> lock(sh->mu);
    const size_t g = sh->shards.size(), block = (size_t)batch * k;
    {
        DeviceGuard guard(sh->merge_device);
        CUDA_TRY(sh->d_gather.reserve(g * 2 * block * 8));
        CUDA_TRY(sh->d_out_keys.reserve(block * 8));
        CUDA_TRY(sh->d_out_hits.reserve(block * sizeof(fsgpu_hit)));
        CUDA_TRY(sh->d_out_counts.reserve((size_t)batch * 4));
        // a grown buffer must exist before any shard stores into it
        CUDA_TRY(cudaStreamSynchronize(sh->merge_stream));
    }
    sh->job_queries = queries;
    sh->job_batch = batch;
    sh->job_k = k;
    for (Shard* s : sh->shards) {  // every shard enqueues its search from its own host thread, concurrently
        {
            std::lock_guard<std::mutex> lk(s->mu);
            s->has_job = true;
            s->job_done = false;
        }
        s->cv.notify_all();
    }
    int rc = 0;
    std::string err;
    for (Shard* s : sh->shards) {
        std::unique_lock<std::mutex> lk(s->mu);
        s->cv.wait(lk, [&] { return s->job_done; });
        if (s->rc && !rc) {
            rc = s->rc;
            err = s->error;
        }
    }
    DeviceGuard guard(sh->merge_device);
    if (rc) {
        for (Shard* s : sh->shards) {
            DeviceGuard gs(s->device);
            cudaStreamSynchronize(s->stream);
        }
        return fail(rc, "%s", err.c_str());
    }
    for (Shard* s : sh->shards) CUDA_TRY(cudaStreamWaitEvent(sh->merge_stream, s->done, 0));
    const uint64_t* keys = sh->d_gather.as<uint64_t>();
    rc = fsgpu_merge_top_k_hits_device(sh->merge_device, keys, reinterpret_cast<const fsgpu_hit*>(keys + block), batch, (uint32_t)g,
                                       k, 2 * block, k, k, sh->d_out_keys.as<uint64_t>(), sh->d_out_hits.as<fsgpu_hit>(),
                                       sh->d_out_counts.as<uint32_t>(), sh->merge_stream);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(out, sh->d_out_hits.p, block * sizeof(fsgpu_hit), cudaMemcpyDeviceToHost, sh->merge_stream));
    CUDA_TRY(cudaMemcpyAsync(out_counts, sh->d_out_counts.p, (size_t)batch * 4, cudaMemcpyDeviceToHost, sh->merge_stream));
    CUDA_TRY(cudaStreamSynchronize(sh->merge_stream));
    // the shards' own flag words (contract violations) are reported by fsgpu_index_last_status
    for (Shard* s : sh->shards) {
        uint32_t flags[4];
        const int rs = fsgpu_index_last_status(s->index, flags);
        if (rs) return rs;
    }
    return FSGPU_OK;
}
