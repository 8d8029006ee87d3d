// Internal context shared by the translation units of libsandstorm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <vector>
#include "../../include/sandstorm_b200.h"
#include "fp252.cuh"

struct ss_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;            // ctx-owned default stream
    std::string err;
    // cached device tables, keyed by (kind, log_n, variant)
    std::map<std::tuple<int, int, int>, void *> tables;
    // grow-only scratch (LDE coefficient buffer etc.)
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    int sm_count = 148;
    unsigned long long launches = 0;          // kernels launched by this ctx (ss_kernel_launches)
    std::map<std::string, long long> options; // tuning switches / read-back values (ss_set_option, ss_get_option)
    // device blocks released by dev_free, kept for reuse (size -> pointers) and the size of every live block
    std::multimap<size_t, void *> pool_free;
    std::map<void *, size_t> pool_size;
    // trees built through this context and not yet freed: ss_destroy releases their device memory and detaches them, so a
    // later ss_tree_free of such a tree only deletes the handle instead of touching a dead context
    std::set<struct ss_tree *> live_trees;
    void (*detach_tree)(struct ss_tree *) = nullptr;
    // geometric scale tables c * h^k of the sharded transforms, keyed by (log length, c, h) (ntt_host.cu custom_scale)
    std::map<std::vector<uint32_t>, std::pair<void *, void *>> scale_tables;
    // pinned staging area for small host -> device uploads that must not block the caller (program blobs)
    void *stage = nullptr;
    size_t stage_bytes = 0;
    cudaEvent_t stage_done = nullptr;
    bool stage_busy = false;
};

namespace ss {

inline long long option(const ss_ctx *ctx, const char *key, long long dflt) {
    auto it = ctx->options.find(key);
    return it == ctx->options.end() ? dflt : it->second;
}

inline ss_status fail(ss_ctx *ctx, ss_status code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define SS_CUDA_CHECK(ctx, expr)                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return ss::fail(ctx, e_ == cudaErrorMemoryAllocation ? SS_ERR_OOM : SS_ERR_CUDA,       \
                            "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// `stream` follows the CUDA convention: a cudaStream_t passed as void*, NULL = the default stream.
inline cudaStream_t pick_stream(ss_ctx *, void *stream) { return reinterpret_cast<cudaStream_t>(stream); }

ss_status scratch_reserve(ss_ctx *ctx, size_t bytes, void **out);
// Caching device allocator for the per-call work buffers (Merkle trees, OOD weights, partial sums): a prove step
// asks for the same sizes again and again, and cudaMalloc / cudaFree of multi-GB blocks cost tens to hundreds of
// milliseconds and synchronise the device.  Blocks are handed out again in stream order (one stream per ctx).
cudaError_t dev_alloc(ss_ctx *ctx, void **out, size_t bytes);
void dev_free(ss_ctx *ctx, void *ptr);
void dev_trim(ss_ctx *ctx);
// Stream-ordered upload of a pageable host buffer without synchronising the stream: the bytes are copied into a
// pinned staging area owned by the context (waiting only for the PREVIOUS staged upload, if it is still in flight).
ss_status stage_upload(ss_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, cudaStream_t st);
// Goldilocks transforms (ntt_goldilocks.cu): one u64 per element, same semantics as the Fp252 entry points
ss_status gl_ntt(ss_ctx *ctx, void *d_cols, uint64_t col_stride, int n_cols, int log_n, int inverse, int coset, ss_order in_order,
                 ss_order out_order, cudaStream_t st);
ss_status gl_lde(ss_ctx *ctx, const void *d_trace, uint64_t trace_stride, int n_cols, int log_n, int log_blowup, void *d_lde,
                 uint64_t lde_stride, void *d_coeffs, uint64_t coeff_stride, ss_order out_order, cudaStream_t st);
// uploads a host table once and caches it; `fill` computes n elements of 32 bytes
ss_status cached_table(ss_ctx *ctx, std::tuple<int, int, int> key, size_t n_elems,
                       void (*fill)(Fp *dst, size_t n, int log_n, int variant), Fp **out);

}  // namespace ss
