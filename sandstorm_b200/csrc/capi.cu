// Context, memory helpers and error plumbing of the C ABI (include/sandstorm_b200.h).
#include "ctx.h"
#include <cstring>

using namespace ss;

namespace ss {

ss_status scratch_reserve(ss_ctx *ctx, size_t bytes, void **out) {
    if (bytes > ctx->scratch_bytes) {
        if (ctx->scratch) {
            SS_CUDA_CHECK(ctx, cudaDeviceSynchronize());
            SS_CUDA_CHECK(ctx, cudaFree(ctx->scratch));
            ctx->scratch = nullptr;
            ctx->scratch_bytes = 0;
        }
        SS_CUDA_CHECK(ctx, cudaMalloc(&ctx->scratch, bytes));
        ctx->scratch_bytes = bytes;
    }
    *out = ctx->scratch;
    return SS_OK;
}

cudaError_t dev_alloc(ss_ctx *ctx, void **out, size_t bytes) {
    const size_t want = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);     // 1 MiB granules: sizes repeat exactly
    auto it = ctx->pool_free.lower_bound(want);
    if (it != ctx->pool_free.end() && it->first <= want + want / 8) {             // reuse a block at most 12.5 % larger
        *out = it->second;
        ctx->pool_free.erase(it);
        return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, want);
    if (e == cudaErrorMemoryAllocation) {                                            // give the cached blocks back and retry
        cudaGetLastError();
        dev_trim(ctx);
        e = cudaMalloc(out, want);
    }
    if (e == cudaSuccess) ctx->pool_size[*out] = want;
    else *out = nullptr;
    return e;
}

void dev_free(ss_ctx *ctx, void *ptr) {
    if (!ptr) return;
    auto it = ctx->pool_size.find(ptr);
    if (it == ctx->pool_size.end()) { cudaFree(ptr); return; }
    ctx->pool_free.emplace(it->second, ptr);
}

ss_status stage_upload(ss_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st) {
    if (ctx->stage_busy) {
        SS_CUDA_CHECK(ctx, cudaEventSynchronize(ctx->stage_done));
        ctx->stage_busy = false;
    }
    if (bytes > ctx->stage_bytes) {
        if (ctx->stage) SS_CUDA_CHECK(ctx, cudaFreeHost(ctx->stage));
        ctx->stage = nullptr;
        ctx->stage_bytes = 0;
        const size_t want = (bytes + (1u << 20) - 1) & ~(size_t)((1u << 20) - 1);
        SS_CUDA_CHECK(ctx, cudaHostAlloc(&ctx->stage, want, cudaHostAllocDefault));
        ctx->stage_bytes = want;
    }
    if (!ctx->stage_done) SS_CUDA_CHECK(ctx, cudaEventCreateWithFlags(&ctx->stage_done, cudaEventDisableTiming));
    memcpy(ctx->stage, h_src, bytes);
    SS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_dst, ctx->stage, bytes, cudaMemcpyHostToDevice, st));
    SS_CUDA_CHECK(ctx, cudaEventRecord(ctx->stage_done, st));
    ctx->stage_busy = true;
    return SS_OK;
}

void dev_trim(ss_ctx *ctx) {
    cudaDeviceSynchronize();
    for (auto &kv : ctx->pool_free) { cudaFree(kv.second); ctx->pool_size.erase(kv.second); }
    ctx->pool_free.clear();
}

ss_status cached_table(ss_ctx *ctx, std::tuple<int, int, int> key, size_t n_elems,
                       void (*fill)(Fp *dst, size_t n, int log_n, int variant), Fp **out) {
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end()) {
        *out = static_cast<Fp *>(it->second);
        return SS_OK;
    }
    std::vector<Fp> host(n_elems);
    fill(host.data(), n_elems, std::get<1>(key), std::get<2>(key));
    void *d = nullptr;
    SS_CUDA_CHECK(ctx, cudaMalloc(&d, n_elems * sizeof(Fp)));
    // synchronous copy: the pageable staging vector dies at scope exit
    SS_CUDA_CHECK(ctx, cudaMemcpy(d, host.data(), n_elems * sizeof(Fp), cudaMemcpyHostToDevice));
    ctx->tables[key] = d;
    *out = static_cast<Fp *>(d);
    return SS_OK;
}

}  // namespace ss

extern "C" {

int ss_version(void) { return 100; }

ss_status ss_create(int device, ss_ctx **out) {
    if (!out) return SS_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SS_ERR_CUDA;   // fail loudly: no CPU fallback
    if (device < 0 || device >= count) return SS_ERR_INVALID;
    if (cudaSetDevice(device) != cudaSuccess) return SS_ERR_CUDA;
    ss_ctx *ctx = new ss_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return SS_ERR_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    *out = ctx;
    return SS_OK;
}

void ss_destroy(ss_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (ss_tree *t : ctx->live_trees)
        if (ctx->detach_tree) ctx->detach_tree(t);            // their blocks are freed with the pool below
    ctx->live_trees.clear();
    dev_trim(ctx);
    for (auto &kv : ctx->pool_size) cudaFree(kv.first);       // blocks still held by (now detached) trees die with the context
    for (auto &kv : ctx->tables) cudaFree(kv.second);
    for (auto &kv : ctx->scale_tables) { cudaFree(kv.second.first); cudaFree(kv.second.second); }
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->stage) cudaFreeHost(ctx->stage);
    if (ctx->stage_done) cudaEventDestroy(ctx->stage_done);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *ss_last_error(const ss_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

ss_status ss_sync(ss_ctx *ctx) {
    if (!ctx) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaDeviceSynchronize());
    return SS_OK;
}

uint64_t ss_kernel_launches(const ss_ctx *ctx) { return ctx ? ctx->launches : 0; }

ss_status ss_set_option(ss_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return SS_ERR_INVALID;
    ctx->options[key] = value;
    return SS_OK;
}
int64_t ss_get_option(const ss_ctx *ctx, const char *key, int64_t dflt) { return (ctx && key) ? option(ctx, key, dflt) : dflt; }

ss_status ss_malloc(ss_ctx *ctx, size_t bytes, void **d_ptr) {
    if (!ctx || !d_ptr) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaMalloc(d_ptr, bytes));
    return SS_OK;
}

ss_status ss_free(ss_ctx *ctx, void *d_ptr) {
    if (!ctx) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaFree(d_ptr));
    return SS_OK;
}

ss_status ss_host_register(ss_ctx *ctx, void *h_ptr, size_t bytes) {
    if (!ctx || !h_ptr) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault));
    return SS_OK;
}

ss_status ss_host_unregister(ss_ctx *ctx, void *h_ptr) {
    if (!ctx || !h_ptr) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaHostUnregister(h_ptr));
    return SS_OK;
}

ss_status ss_memcpy_h2d(ss_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, pick_stream(ctx, stream)));
    return SS_OK;
}

ss_status ss_memcpy_d2h(ss_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, pick_stream(ctx, stream)));
    return SS_OK;
}

}  // extern "C"
