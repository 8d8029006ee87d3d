// Several GPUs behind the C ABI (SURVEY.md §8e plan B): the collectives of the row-sharded hot path issued
// from C++ over NCCL, so that a host without torch (the Rust prover of INTEGRATION.md) drives W GPUs with one context per
// process / device.  The algorithms are the ones sandstorm_b200/parallel.py specifies (and tests/test_parallel_gloo.py checks
// with big-int local operations); this file only replaces torch.distributed by NCCL calls around the same local kernels.
//
// NCCL is bound at run time (dlopen of libnccl.so.2 — the copy already loaded in the process if there is one), so the
// library has no link-time dependency on it and single-GPU users never touch it.
//
//   ss_dist_unique_id / ss_dist_init / ss_dist_finalize    communicator life cycle (the 128-byte id travels by the host's means)
//   ss_dist_lde          one column: block-cyclic evaluations on <w_n> (or 3<w_n>) -> block-cyclic evaluations on 3<w_N>
//   ss_dist_halo         after an LDE phase: the `halo` rows that follow every owned piece, from the next rank
//   ss_dist_commit       bit-reversed-order Merkle root of a block-cyclic matrix (digest all-to-all + sub-trees + combine)
//   ss_dist_allgather    in-place all-gather of a block-cyclic vector (the DEEP evaluations before FRI)
//   ss_dist_open / ss_dist_gather_rows    query phase: authentication paths through the sharded tree, opened rows (collective)
#include "ctx.h"
#include <dlfcn.h>
#include <cstring>

using namespace ss;

namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_UINT8 = 1 };

struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

NcclApi &nccl() {
    static NcclApi api;
    if (api.lib) return api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    auto sym = [&](const char *name) { return dlsym(h, name); };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    if (api.GetUniqueId && api.CommInitRank && api.Send && api.Recv && api.AllGather && api.AllReduce && api.GroupStart && api.GroupEnd) api.lib = h;
    return api;
}

struct Dist {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, log_w = 0;
};
std::map<ss_ctx *, Dist> g_dist;

#define SS_NCCL_CHECK(ctx, expr)                                                                                  \
    do {                                                                                                          \
        const int r_ = (expr);                                                                                    \
        if (r_ != 0) return ss::fail(ctx, SS_ERR_CUDA, "%s failed: %s", #expr, nccl().GetErrorString ? nccl().GetErrorString(r_) : "?"); \
    } while (0)

ss_status get(ss_ctx *ctx, Dist **out) {
    auto it = g_dist.find(ctx);
    if (it == g_dist.end() || !it->second.comm) return fail(ctx, SS_ERR_INVALID, "ss_dist_*: call ss_dist_init first");
    *out = &it->second;
    return SS_OK;
}

// out[src] <- in[dst] of rank src; pieces of `bytes` each
ss_status all_to_all(ss_ctx *ctx, Dist *d, const uint8_t *send, uint8_t *recv, size_t bytes, cudaStream_t st) {
    SS_NCCL_CHECK(ctx, nccl().GroupStart());
    for (int q = 0; q < d->world; ++q) {
        SS_NCCL_CHECK(ctx, nccl().Send(send + (size_t)q * bytes, bytes, NCCL_UINT8, q, d->comm, st));
        SS_NCCL_CHECK(ctx, nccl().Recv(recv + (size_t)q * bytes, bytes, NCCL_UINT8, q, d->comm, st));
    }
    SS_NCCL_CHECK(ctx, nccl().GroupEnd());
    return SS_OK;
}

int brev_small(int v, int bits) {
    int r = 0;
    for (int i = 0; i < bits; ++i) r |= ((v >> i) & 1) << (bits - 1 - i);
    return r;
}

// canonical-int arithmetic for the few host-side constants (3^r / n, 3^W, w_N^r): Montgomery form via the fp:: host paths
Fp fp_pow_small(Fp base, unsigned long long e) { return fp::pow_u64(base, e); }

// dst[q][k1][u] = src[k1][u][brev_w(q)]: the digest of row k1 m + r s + u W + c goes to rank brev_w(c)
__global__ void digest_route_kernel(const uint4 *src, uint4 *dst, unsigned long long s_over_w, int W, int log_w) {
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;   // over W * W * s_over_w outputs
    const unsigned long long total = (unsigned long long)W * W * s_over_w;
    if (t >= total) return;
    const unsigned long long u = t % s_over_w, k1 = (t / s_over_w) % W, q = t / (s_over_w * W);
    const unsigned int c = __brev((unsigned)q) >> (32 - log_w);
    const unsigned long long from = (k1 * s_over_w + u) * W + c;
    dst[2 * t] = src[2 * from];
    dst[2 * t + 1] = src[2 * from + 1];
}
// dst[k1][src][u] = recv[src][k1][u]
__global__ void digest_order_kernel(const uint4 *recv, uint4 *dst, unsigned long long s_over_w, int W) {
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long total = (unsigned long long)W * W * s_over_w;
    if (t >= total) return;
    const unsigned long long u = t % s_over_w, src = (t / s_over_w) % W, k1 = t / (s_over_w * W);
    const unsigned long long from = (src * W + k1) * s_over_w + u;
    dst[2 * t] = recv[2 * from];
    dst[2 * t + 1] = recv[2 * from + 1];
}

}  // namespace

extern "C" {

ss_status ss_dist_unique_id(ss_ctx *ctx, uint8_t id[128]) {
    if (!ctx || !id) return SS_ERR_INVALID;
    if (!nccl().lib) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_dist_unique_id: libnccl.so.2 not found");
    ncclUniqueId u;
    SS_NCCL_CHECK(ctx, nccl().GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return SS_OK;
}

ss_status ss_dist_init(ss_ctx *ctx, const uint8_t id[128], int rank, int world) {
    if (!ctx || !id || world < 2 || world > 8 || (world & (world - 1)) || rank < 0 || rank >= world)
        return fail(ctx, SS_ERR_INVALID, "ss_dist_init: world must be 2, 4 or 8 and 0 <= rank < world");
    if (!nccl().lib) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_dist_init: libnccl.so.2 not found");
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    Dist d;
    d.rank = rank; d.world = world;
    while ((1 << d.log_w) < world) ++d.log_w;
    SS_NCCL_CHECK(ctx, nccl().CommInitRank(&d.comm, world, u, rank));
    g_dist[ctx] = d;
    return SS_OK;
}

ss_status ss_dist_finalize(ss_ctx *ctx) {
    auto it = g_dist.find(ctx);
    if (it == g_dist.end()) return SS_OK;
    if (it->second.comm && nccl().CommDestroy) nccl().CommDestroy(it->second.comm);
    g_dist.erase(it);
    return SS_OK;
}

ss_status ss_dist_lde(ss_ctx *ctx, ss_field field, const void *d_src, int log_n, int log_blowup, int src_on_coset, void *d_dst,
                      void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    Dist *d;
    ss_status rc = get(ctx, &d);
    if (rc) return rc;
    if (field != SS_FIELD_FP252 || !d_src || !d_dst || log_blowup < 0 || log_n < 2 * d->log_w + 1)
        return fail(ctx, SS_ERR_INVALID, "ss_dist_lde: bad arguments");
    const int W = d->world, r = d->rank, log_w = d->log_w;
    const unsigned long long n = 1ull << log_n, m = n / W, s = m / W, mN = m << log_blowup, sN = mN / W;
    cudaStream_t st = pick_stream(ctx, stream);
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    Fp *send, *recv, *z, *recv2;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&send), (size_t)m * 32));
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&recv), (size_t)m * 32));
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&z), (size_t)mN * 32));
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&recv2), (size_t)mN * 32));
    auto done = [&](ss_status code) { dev_free(ctx, send); dev_free(ctx, recv); dev_free(ctx, z); dev_free(ctx, recv2); return code; };
    const Fp *src = static_cast<const Fp *>(d_src);
    Fp *dst = static_cast<Fp *>(d_dst);
    if ((rc = ss_shard_dft(ctx, field, src + (size_t)r * s, m, send, s, s, log_w, 1, log_n, (uint64_t)r * s, st))) return done(rc);
    if ((rc = all_to_all(ctx, d, reinterpret_cast<const uint8_t *>(send), reinterpret_cast<uint8_t *>(recv), (size_t)s * 32, st))) return done(rc);
    // c0 = 3^r / n (or 1 / n for a column already on the coset), h0 = 3^W (or 1), tw = w_N^r
    const Fp three = fp::from_u32(3);
    const Fp ninv = fp::inv(fp::pow_u64(fp::from_u32(2), (unsigned long long)log_n));
    Fp c0 = src_on_coset ? ninv : fp::mul(ninv, fp::pow_u64(three, (unsigned long long)r));
    Fp h0 = src_on_coset ? fp::one() : fp::pow_u64(three, (unsigned long long)W);
    c0 = fp::canon(c0); h0 = fp::canon(h0);
    Fp tw = fp::one();
    if (r) {
        // w_N = 3^((p-1) / N): (p - 1) >> log_N as limbs
        uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0x00000011u, 0x08000000u};
        for (int sft = 0; sft < log_n + log_blowup; ++sft)
            for (int i = 0; i < 8; ++i) { e[i] >>= 1; if (i < 7) e[i] |= e[i + 1] << 31; }
        tw = fp::canon(fp::pow_u64(fp::pow_limbs(three, e, 8), (unsigned long long)r));
    }
    int log_m = 0;
    while ((1ull << log_m) < m) ++log_m;
    if ((rc = ss_ntt_shard(ctx, field, recv, m, 1, log_m, 3, log_blowup, c0.l, h0.l, r ? tw.l : nullptr, z, mN, st))) return done(rc);
    if ((rc = all_to_all(ctx, d, reinterpret_cast<const uint8_t *>(z), reinterpret_cast<uint8_t *>(recv2), (size_t)sN * 32, st))) return done(rc);
    if ((rc = ss_shard_dft(ctx, field, recv2, sN, dst + (size_t)r * sN, mN, sN, log_w, 0, -1, 0, st))) return done(rc);
    return done(SS_OK);
}

ss_status ss_dist_halo(ss_ctx *ctx, void *d_cols, uint64_t col_stride, int n_cols, int log_rows, uint64_t halo, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    Dist *d;
    ss_status rc = get(ctx, &d);
    if (rc) return rc;
    const int W = d->world, r = d->rank;
    const unsigned long long N = 1ull << log_rows, m = N / W, s = m / W;
    if (!d_cols || n_cols < 1 || halo > s) return fail(ctx, SS_ERR_INVALID, "ss_dist_halo: halo larger than a piece (gather the columns instead)");
    if (halo == 0) return SS_OK;
    cudaStream_t st = pick_stream(ctx, stream);
    const int prv = (r + W - 1) % W, nxt = (r + 1) % W;
    uint8_t *base = static_cast<uint8_t *>(d_cols);
    SS_NCCL_CHECK(ctx, nccl().GroupStart());
    for (int j = 0; j < n_cols; ++j)
        for (int k1 = 0; k1 < W; ++k1) {
            const unsigned long long mine = (unsigned long long)k1 * m + (unsigned long long)r * s, theirs = (unsigned long long)k1 * m + (unsigned long long)nxt * s;
            SS_NCCL_CHECK(ctx, nccl().Send(base + 32ull * ((unsigned long long)j * col_stride + mine), halo * 32, NCCL_UINT8, prv, d->comm, st));
            SS_NCCL_CHECK(ctx, nccl().Recv(base + 32ull * ((unsigned long long)j * col_stride + theirs), halo * 32, NCCL_UINT8, nxt, d->comm, st));
        }
    SS_NCCL_CHECK(ctx, nccl().GroupEnd());
    return SS_OK;
}

ss_status ss_dist_allgather(ss_ctx *ctx, void *d_vec, int log_rows, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    Dist *d;
    ss_status rc = get(ctx, &d);
    if (rc) return rc;
    const int W = d->world, r = d->rank;
    const unsigned long long N = 1ull << log_rows, m = N / W, s = m / W;
    if (!d_vec || s == 0) return fail(ctx, SS_ERR_INVALID, "ss_dist_allgather: bad arguments");
    cudaStream_t st = pick_stream(ctx, stream);
    uint8_t *v = static_cast<uint8_t *>(d_vec);
    // piece k1 of rank q lives at rows k1 m + q s: W all-gathers of s rows each, straight into place
    SS_NCCL_CHECK(ctx, nccl().GroupStart());
    for (int k1 = 0; k1 < W; ++k1)
        SS_NCCL_CHECK(ctx, nccl().AllGather(v + 32ull * ((unsigned long long)k1 * m + (unsigned long long)r * s), v + 32ull * (unsigned long long)k1 * m, (size_t)s * 32,
                                            NCCL_UINT8, d->comm, st));
    SS_NCCL_CHECK(ctx, nccl().GroupEnd());
    return SS_OK;
}

ss_status ss_dist_commit(ss_ctx *ctx, ss_tree_kind kind, int n_friendly, const void *d_cols, uint64_t col_stride, int n_cols,
                         int log_rows, uint8_t root[32], ss_tree **out_subtree, uint8_t *out_subroots, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    Dist *d;
    ss_status rc = get(ctx, &d);
    if (rc) return rc;
    const int W = d->world, r = d->rank, log_w = d->log_w;
    if (!d_cols || !root || n_cols < 1 || log_rows < 2 * log_w + 1) return fail(ctx, SS_ERR_INVALID, "ss_dist_commit: bad arguments");
    const unsigned long long N = 1ull << log_rows, m = N / W, s = m / W, per = N / W;
    cudaStream_t st = pick_stream(ctx, stream);
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    uint8_t *D, *send, *recv;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&D), (size_t)per * 32));
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&send), (size_t)per * 32));
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&recv), (size_t)per * 32));
    auto done = [&](ss_status code) { dev_free(ctx, D); dev_free(ctx, send); dev_free(ctx, recv); return code; };
    const uint8_t *cols = static_cast<const uint8_t *>(d_cols);
    for (int k1 = 0; k1 < W; ++k1) {
        const unsigned long long lo = (unsigned long long)k1 * m + (unsigned long long)r * s;
        if (n_cols == 1) {
            SS_CUDA_CHECK(ctx, cudaMemcpyAsync(D + 32ull * k1 * s, cols + 32ull * lo, (size_t)s * 32, cudaMemcpyDeviceToDevice, st));
        } else if ((rc = ss_hash_rows(ctx, kind, d_cols, col_stride, n_cols, log_rows, SS_ORDER_NATURAL, lo, s, D + 32ull * k1 * s, st))) {
            return done(rc);
        }
    }
    const unsigned long long total = (unsigned long long)W * W * (s / W);
    digest_route_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(D), reinterpret_cast<uint4 *>(send), s / W, W, log_w);
    if ((rc = all_to_all(ctx, d, send, recv, (size_t)(per / W) * 32, st))) return done(rc);
    digest_order_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4 *>(recv), reinterpret_cast<uint4 *>(D), s / W, W);
    ctx->launches += 2;
    if ((rc = ss_bitrev_permute32(ctx, D, log_rows - log_w, st))) return done(rc);
    const int friendly = kind == SS_TREE_FRIENDLY ? (n_friendly > log_w ? n_friendly - log_w : 0) : 0;
    ss_tree *tree = nullptr;
    if (n_cols == 1) rc = ss_merkle_build(ctx, kind, friendly, D, per, 1, log_rows - log_w, SS_ORDER_NATURAL, &tree, st);
    else rc = ss_merkle_build_from_leaves(ctx, kind, friendly, D, log_rows - log_w, &tree, st);
    if (rc) return done(rc);
    uint8_t sub[32];
    rc = ss_merkle_root(ctx, tree, sub);
    if (rc || !out_subtree) ss_tree_free(tree);
    if (rc) return done(rc);
    if (out_subtree) *out_subtree = tree;                   // kept for ss_dist_open; the caller frees it
    // all-gather of the W sub-roots through a small device buffer
    uint8_t *d_roots = recv;
    SS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_roots + 32 * r, sub, 32, cudaMemcpyHostToDevice, st));
    SS_NCCL_CHECK(ctx, nccl().AllGather(d_roots + 32 * r, d_roots, 32, NCCL_UINT8, d->comm, st));
    uint8_t subs[8 * 32];
    SS_CUDA_CHECK(ctx, cudaMemcpyAsync(subs, d_roots, 32 * (size_t)W, cudaMemcpyDeviceToHost, st));
    SS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    if (out_subroots) memcpy(out_subroots, subs, 32 * (size_t)W);
    rc = ss_merkle_combine(ctx, kind, subs, log_w, root);
    return done(rc);
}

// every rank contributes the bytes of the slots it owns (zero elsewhere): sum = the whole buffer, on every rank
static ss_status share_bytes(ss_ctx *ctx, Dist *d, uint8_t *h_buf, size_t bytes, cudaStream_t st) {
    uint8_t *dev = nullptr;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&dev), bytes));
    cudaMemcpyAsync(dev, h_buf, bytes, cudaMemcpyHostToDevice, st);
    const int r = nccl().AllReduce(dev, dev, bytes, NCCL_UINT8, /*ncclSum*/ 0, d->comm, st);
    cudaMemcpyAsync(h_buf, dev, bytes, cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    dev_free(ctx, dev);
    if (r != 0) return fail(ctx, SS_ERR_CUDA, "ncclAllReduce failed: %s", nccl().GetErrorString ? nccl().GetErrorString(r) : "?");
    if (e != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_dist: %s", cudaGetErrorString(e));
    return SS_OK;
}

ss_status ss_dist_open(ss_ctx *ctx, ss_tree_kind kind, int subroots_algebraic, const ss_tree *subtree, const uint8_t *h_subroots,
                       const uint64_t *h_positions, size_t n, uint8_t *h_paths, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    Dist *d;
    ss_status rc = get(ctx, &d);
    if (rc) return rc;
    if (!subtree || !h_subroots || (n && (!h_positions || !h_paths))) return fail(ctx, SS_ERR_INVALID, "ss_dist_open: bad arguments");
    if (n == 0) return SS_OK;
    const int sub_log = ss_tree_log_rows(subtree), log_w = d->log_w, depth = sub_log + log_w;
    std::vector<uint64_t> local;
    std::vector<size_t> slot;
    for (size_t i = 0; i < n; ++i) {
        if (h_positions[i] >> depth) return fail(ctx, SS_ERR_INVALID, "ss_dist_open: position out of range");
        if ((int)(h_positions[i] >> sub_log) == d->rank) { local.push_back(h_positions[i] & ((1ull << sub_log) - 1)); slot.push_back(i); }
    }
    memset(h_paths, 0, n * (size_t)depth * 32);
    if (!local.empty()) {
        std::vector<uint8_t> low(local.size() * (size_t)sub_log * 32), top((size_t)log_w * 32);
        if ((rc = ss_merkle_open(ctx, subtree, local.data(), local.size(), low.data()))) return rc;
        if ((rc = ss_merkle_combine_open(ctx, kind, h_subroots, log_w, subroots_algebraic, (uint64_t)d->rank, top.data()))) return rc;
        for (size_t k = 0; k < local.size(); ++k) {
            uint8_t *out = h_paths + slot[k] * (size_t)depth * 32;
            memcpy(out, low.data() + k * (size_t)sub_log * 32, (size_t)sub_log * 32);
            memcpy(out + (size_t)sub_log * 32, top.data(), top.size());
        }
    }
    return share_bytes(ctx, d, h_paths, n * (size_t)depth * 32, pick_stream(ctx, stream));
}

ss_status ss_dist_gather_rows(ss_ctx *ctx, const void *d_cols, uint64_t col_stride, int n_cols, int log_rows, const uint64_t *h_indices,
                              size_t n, void *h_rows, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    Dist *d;
    ss_status rc = get(ctx, &d);
    if (rc) return rc;
    if (!d_cols || n_cols < 1 || log_rows < 2 * d->log_w || (n && (!h_indices || !h_rows))) return fail(ctx, SS_ERR_INVALID, "ss_dist_gather_rows: bad arguments");
    if (n == 0) return SS_OK;
    const unsigned long long m = (1ull << log_rows) / d->world, s = m / d->world;
    const size_t row_bytes = (size_t)n_cols * 32;
    std::vector<uint64_t> mine;
    std::vector<size_t> slot;
    for (size_t i = 0; i < n; ++i) {
        if (h_indices[i] >> log_rows) return fail(ctx, SS_ERR_INVALID, "ss_dist_gather_rows: row out of range");
        if ((int)((h_indices[i] % m) / s) == d->rank) { mine.push_back(h_indices[i]); slot.push_back(i); }
    }
    uint8_t *out = static_cast<uint8_t *>(h_rows);
    memset(out, 0, n * row_bytes);
    if (!mine.empty()) {
        std::vector<uint8_t> got(mine.size() * row_bytes);
        if ((rc = ss_rows_gather(ctx, d_cols, col_stride, n_cols, mine.data(), mine.size(), got.data()))) return rc;
        for (size_t k = 0; k < mine.size(); ++k) memcpy(out + slot[k] * row_bytes, got.data() + k * row_bytes, row_bytes);
    }
    return share_bytes(ctx, d, out, n * row_bytes, pick_stream(ctx, stream));
}

}  // extern "C"
