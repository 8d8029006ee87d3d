// Merkle commitment of a column-major Fp252 matrix on the GPU: the reference's
// MatrixMerkleTree::from_matrix variants (crypto/src/merkle/mod.rs:110-123, :289-304), row hashing
// (crypto/src/merkle/utils.rs:9-46), level / hash selection (crypto/src/merkle/mixed.rs:110-125,
// :148-155) and single-column first level (crypto/src/merkle/mod.rs:422-437).
//
// Node array layout follows ministark's MerkleTreeImpl ([RECALLED]): nodes[1] is the
// root, nodes[i] = H(nodes[2i], nodes[2i+1]) with depth = floor(log2 i), nodes[n/2 + i] built from
// leaves 2i, 2i+1 with depth = log2(n) - 1.  A node is 32 bytes: the byte digest, or the Montgomery
// limbs of the felt for algebraic (Pedersen) levels.
//
// HBM traffic (algorithmic): leaves N*32*C read + N*32 written; nodes (N-1)*(64 read + 32 written).
#include "ctx.h"
#include <cstring>
#include "hashes.cuh"
#include "pedersen.cuh"

using namespace ss;

struct ss_tree {
    ss_ctx *ctx;
    int kind, n_friendly, log_rows, n_cols;
    uint8_t *d_leaves;   // n * 32: row digests, or the raw column when n_cols == 1
    uint8_t *d_nodes;    // n * 32: slot 0 unused
};

namespace {

enum ByteHash : int { BH_KECCAK = 0, BH_BLAKE2S = 1, BH_SHA256 = 2 };
enum Mask : int { MASK_NONE = 0, MASK_KEEP_FIRST20 = 1, MASK_KEEP_LAST20 = 2 };

constexpr int T_PEDERSEN = 10;

__device__ __forceinline__ unsigned long long brev_bits(unsigned long long x, int bits) {
    return bits ? (__brevll(x) >> (64 - bits)) : 0ull;
}

// digest words (as stored in memory, little-endian u32 view of the 32 digest bytes) with the mask applied
__device__ __forceinline__ void store_digest(uint8_t *dst, const uint32_t (&w)[8], int mask) {
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = w[i];
    if (mask == MASK_KEEP_FIRST20) { o[5] = 0; o[6] = 0; o[7] = 0; }      // bytes 20..31
    if (mask == MASK_KEEP_LAST20) { o[0] = 0; o[1] = 0; o[2] = 0; }       // bytes 0..11
    uint4 *q = reinterpret_cast<uint4 *>(dst);
    q[0] = make_uint4(o[0], o[1], o[2], o[3]);
    q[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

// Hash of a message made of `n_elems` 32-byte big-endian felts: element e of this message lives at
// base + e * elem_stride (in Fp units).  Used for rows (elem_stride = column stride) and for the
// single-column first level (two consecutive leaves).
// col_bits != 0: message element e is read from column brev_{col_bits}(e) (FRI layer rows, see ss_order).
template <int BH>
__device__ __forceinline__ void hash_felts(const Fp *base, unsigned long long elem_stride, int n_elems, uint32_t (&out)[8], int col_bits = 0) {
    const uint32_t *w32 = reinterpret_cast<const uint32_t *>(base);
    auto col = [&](int e) -> unsigned long long { return col_bits ? (unsigned long long)(__brev((unsigned)e) >> (32 - col_bits)) : (unsigned long long)e; };
    if (BH == BH_KECCAK) {
        // message lane g (8 bytes) = big-endian bytes of u64 limb (3 - g%4) of element g/4
        const unsigned long long *w64 = reinterpret_cast<const unsigned long long *>(base);
        auto lane = [&](int g) -> uint64_t {
            const unsigned long long v = __ldg(w64 + col(g >> 2) * elem_stride * 4ull + (3 - (g & 3)));
            return hash::bswap64(v);
        };
        uint64_t d[4];
        hash::keccak256_lanes(lane, n_elems * 4, d);
#pragma unroll
        for (int i = 0; i < 4; ++i) { out[2 * i] = (uint32_t)d[i]; out[2 * i + 1] = (uint32_t)(d[i] >> 32); }
    } else if (BH == BH_BLAKE2S) {
        // message word g (LE u32 of 4 message bytes) = bswap32 of u32 limb (7 - g%8) of element g/8
        auto word = [&](int g) -> uint32_t {
            return hash::bswap32(__ldg(w32 + col(g >> 3) * elem_stride * 8ull + (7 - (g & 7))));
        };
        hash::blake2s256_words(word, n_elems * 8, out);
    } else {
        // SHA-256 reads big-endian words: exactly the u32 limb, digest words stored big-endian
        auto word = [&](int g) -> uint32_t {
            return __ldg(w32 + col(g >> 3) * elem_stride * 8ull + (7 - (g & 7)));
        };
        uint32_t h[8];
        hash::sha256_words(word, n_elems * 8, h);
#pragma unroll
        for (int i = 0; i < 8; ++i) out[i] = hash::bswap32(h[i]);
    }
}

// Leaf digests.  Tree leaf p commits matrix row (bitrev ? brev(p) : p).  One thread per ROW in natural order:
// consecutive threads read consecutive 32-byte elements of each column (coalesced column streams — the 32 * n_cols
// bytes per row are the kernel's HBM traffic), and the 32-byte digest goes to its leaf slot (a scattered 32-byte
// sector write when the order is bit-reversed).
template <int BH>
__global__ void __launch_bounds__(128) leaf_hash_kernel(const Fp *cols, unsigned long long col_stride, int n_cols,
                                                          int log_rows, int bitrev, int col_bits, int mask, uint8_t *out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >> log_rows) return;
    uint32_t d[8];
    hash_felts<BH>(cols + i, col_stride, n_cols, d, col_bits);
    store_digest(out + 32ull * (bitrev ? brev_bits(i, log_rows) : i), d, mask);
}
// The same for a RANGE of tree leaves [leaf_begin, leaf_begin + count), one thread per leaf (ss_hash_rows: the share of
// one GPU when a commitment is split over several); out[k] = digest of leaf leaf_begin + k.
template <int BH>
__global__ void __launch_bounds__(128) leaf_range_hash_kernel(const Fp *cols, unsigned long long col_stride, int n_cols, int log_rows,
                                                                int bitrev, int col_bits, int mask, unsigned long long leaf_begin,
                                                                unsigned long long count, uint8_t *out) {
    const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (k >= count) return;
    const unsigned long long leaf = leaf_begin + k;
    uint32_t d[8];
    hash_felts<BH>(cols + (bitrev ? brev_bits(leaf, log_rows) : leaf), col_stride, n_cols, d, col_bits);
    store_digest(out + 32ull * k, d, mask);
}

// first level of the single-column variant: node i = H(BE32(leaf 2i) || BE32(leaf 2i+1))
template <int BH>
__global__ void __launch_bounds__(128) leafpair_hash_kernel(const Fp *leaves, unsigned long long count, int mask, uint8_t *out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t d[8];
    hash_felts<BH>(leaves + 2ull * i, 1ull, 2, d);
    store_digest(out + 32ull * i, d, mask);
}

// digest of 64 bytes of children held in 16 registers (little-endian u32 words of the 64 bytes)
template <int BH>
__device__ __forceinline__ void node_digest(const uint32_t (&c)[16], uint32_t (&d)[8]) {
    if (BH == BH_KECCAK) {
        auto lane = [&](int g) -> uint64_t { return (uint64_t)c[2 * g] | ((uint64_t)c[2 * g + 1] << 32); };
        uint64_t h[4];
        hash::keccak256_lanes(lane, 8, h);
#pragma unroll
        for (int k = 0; k < 4; ++k) { d[2 * k] = (uint32_t)h[k]; d[2 * k + 1] = (uint32_t)(h[k] >> 32); }
    } else if (BH == BH_BLAKE2S) {
        auto word = [&](int g) -> uint32_t { return c[g]; };
        hash::blake2s256_words(word, 16, d);
    } else {
        auto word = [&](int g) -> uint32_t { return hash::bswap32(c[g]); };
        uint32_t h[8];
        hash::sha256_words(word, 16, h);
#pragma unroll
        for (int k = 0; k < 8; ++k) d[k] = hash::bswap32(h[k]);
    }
}

// node i = H(child 2i || child 2i+1) over raw digest bytes
template <int BH>
__global__ void __launch_bounds__(128) node_hash_kernel(const uint8_t *children, unsigned long long count, int mask, uint8_t *out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint4 *q = reinterpret_cast<const uint4 *>(children + 64ull * i);
    uint32_t c[16], d[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const uint4 v = __ldg(q + k); c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w; }
    node_digest<BH>(c, d);
    store_digest(out + 32ull * i, d, mask);
}

// FUSED_LEVELS tree levels in one launch: a CTA takes 2^FUSED_LEVELS consecutive children (digests at depth dc), keeps the
// sub-tree above them in shared memory and writes every node it computes to its slot of the node array
// (node j of depth d lives at nodes[2^d + j]).  One read of the children and one write per node instead of a launch and
// an HBM round trip per level.
constexpr int FUSED_LEVELS = 9;
template <int BH>
__global__ void __launch_bounds__(256) node_levels_fused_kernel(const uint8_t *children, int dc, int levels, int mask, uint8_t *nodes) {
    __shared__ uint4 sm[2 << FUSED_LEVELS];                      // 2^levels digests of 32 bytes
    const unsigned long long per = 1ull << levels;
    const uint4 *src = reinterpret_cast<const uint4 *>(children) + 2ull * per * blockIdx.x;
    for (unsigned int k = threadIdx.x; k < 2 * per; k += blockDim.x) sm[k] = __ldg(src + k);
    __syncthreads();
    for (int lvl = 1; lvl <= levels; ++lvl) {
        const unsigned int cnt = 1u << (levels - lvl);           // nodes of this CTA at depth dc - lvl
        uint32_t d[8];
        // (reads of level lvl-1 and writes of level lvl both live in sm[0 .. 2 cnt): separate them with barriers)
        for (unsigned int base = 0; base < cnt; base += blockDim.x) {
            const unsigned int j = base + threadIdx.x;
            const bool on = j < cnt;
            uint32_t c[16];
            if (on) {
#pragma unroll
                for (int k = 0; k < 4; ++k) { const uint4 v = sm[4 * j + k]; c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w; }
                node_digest<BH>(c, d);
                if (mask == MASK_KEEP_FIRST20) { d[5] = 0; d[6] = 0; d[7] = 0; }
                if (mask == MASK_KEEP_LAST20) { d[0] = 0; d[1] = 0; d[2] = 0; }
            }
            __syncthreads();                                     // everybody has read its children of this batch
            if (on) {
                sm[2 * j] = make_uint4(d[0], d[1], d[2], d[3]);
                sm[2 * j + 1] = make_uint4(d[4], d[5], d[6], d[7]);
                uint4 *dst = reinterpret_cast<uint4 *>(nodes + 32ull * ((1ull << (dc - lvl)) + (unsigned long long)cnt * blockIdx.x + j));
                dst[0] = sm[2 * j]; dst[1] = sm[2 * j + 1];
            }
            __syncthreads();
        }
    }
}

// Three levels per launch with every thread busy: a thread owns 8 consecutive children (256 contiguous bytes) and computes
// the 4 + 2 + 1 nodes above them one after the other, in registers.  Same hash count as level-by-level launches (the levels
// are compute-bound), a third of the launches and no re-read of the two intermediate levels.
template <int BH>
__global__ void __launch_bounds__(128) node_levels3_kernel(const uint8_t *children, int dc, unsigned long long count, int mask, uint8_t *nodes) {
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;     // node index at depth dc - 3
    if (t >= count) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(children) + 16ull * t;
    auto put = [&](int depth, unsigned long long j, const uint32_t (&d)[8]) {
        uint4 *dst = reinterpret_cast<uint4 *>(nodes + 32ull * ((1ull << depth) + j));
        dst[0] = make_uint4(d[0], d[1], d[2], d[3]);
        dst[1] = make_uint4(d[4], d[5], d[6], d[7]);
    };
    auto masked = [&](uint32_t (&d)[8]) {
        if (mask == MASK_KEEP_FIRST20) { d[5] = 0; d[6] = 0; d[7] = 0; }
        if (mask == MASK_KEEP_LAST20) { d[0] = 0; d[1] = 0; d[2] = 0; }
    };
    uint32_t top[16];
#pragma unroll
    for (int h = 0; h < 2; ++h) {                       // the two halves of 4 children each
        uint32_t mid[16];
#pragma unroll
        for (int q = 0; q < 2; ++q) {                   // a pair of children -> one node of depth dc - 1
            uint32_t c[16], d[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) { const uint4 v = __ldg(src + 8 * h + 4 * q + k); c[4 * k] = v.x; c[4 * k + 1] = v.y; c[4 * k + 2] = v.z; c[4 * k + 3] = v.w; }
            node_digest<BH>(c, d);
            masked(d);
            put(dc - 1, 4 * t + 2 * h + q, d);
#pragma unroll
            for (int k = 0; k < 8; ++k) mid[8 * q + k] = d[k];
        }
        uint32_t d[8];
        node_digest<BH>(mid, d);
        masked(d);
        put(dc - 2, 2 * t + h, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) top[8 * h + k] = d[k];
    }
    uint32_t d[8];
    node_digest<BH>(top, d);
    masked(d);
    put(dc - 3, t, d);
}

__device__ __forceinline__ Fp load_fp(const uint8_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}
__device__ __forceinline__ void store_fp(uint8_t *p, const Fp &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
// mixed.rs:148-155: digest bytes -> big-endian integer -> Fp (Montgomery form)
__device__ __forceinline__ Fp digest_to_felt(const uint8_t *p) {
    const Fp raw = load_fp(p);
    Fp v;
#pragma unroll
    for (int w = 0; w < 8; ++w) v.l[7 - w] = hash::bswap32(raw.l[w]);
    return fp::canon(fp::mul(v, fp::r2()));
}

struct PedersenTable {
    const AffinePt *pts;   // ec::PED_TABLE_POINTS entries followed by P0
};
__device__ __forceinline__ AffinePt load_pt(const AffinePt *p) {
    AffinePt r;
    r.x = load_fp(reinterpret_cast<const uint8_t *>(&p->x));
    r.y = load_fp(reinterpret_cast<const uint8_t *>(&p->y));
    return r;
}

// mode 0: children are felts (Montgomery);  1: children are byte digests (hash_boundary);
// mode 2: children are felts and the node is PedersenHashFn::hash_elements([l0, l1])
//         = H(H(H(0, l0), l1), 2)   (crypto/src/hash/pedersen.rs:67-76; single-column first level)
__global__ void __launch_bounds__(128) pedersen_node_kernel(const uint8_t *children, unsigned long long count, int mode,
                                                              PedersenTable tab, uint8_t *out) {
    __shared__ Fp inv_sm[256];
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const bool on = i < count;                                   // (every thread takes part in the block-wide inversions)
    Fp a = fp::zero(), b = fp::zero();
    if (on) {
        const uint8_t *c = children + 64ull * i;
        if (mode == 1) { a = digest_to_felt(c); b = digest_to_felt(c + 32); }
        else { a = load_fp(c); b = load_fp(c + 32); }
    }
    const AffinePt p0 = load_pt(tab.pts + ec::PED_TABLE_POINTS);
    auto ld = [&](int idx) { return load_pt(tab.pts + idx); };
    // x = X / Z^2 with the Z's of the block inverted together
    auto hash = [&](const Fp &u, const Fp &v) {
        JacPt s;
        s.x = fp::one(); s.z = fp::one();
        if (on) s = ec::pedersen_sum(u, v, p0, ld);
        if (ec::is_zero_mod_p(s.z)) s.z = fp::one();             // infinity (no valid input gives it): keep the block's product invertible
        const Fp zi = ec::block_inverse<128>(s.z, inv_sm);
        return fp::canon(fp::mul(s.x, fp::sqr(zi)));
    };
    Fp h;
    if (mode == 2) {
        h = hash(fp::zero(), a);
        h = hash(h, b);
        h = hash(h, fp::from_u32(2));
    } else {
        h = hash(a, b);
    }
    if (on) store_fp(out + 32ull * i, h);
}

__global__ void __launch_bounds__(128) pedersen_batch_kernel(const Fp *a, const Fp *b, Fp *out, unsigned long long count, PedersenTable tab) {
    __shared__ Fp inv_sm[256];
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const bool on = i < count;
    const AffinePt p0 = load_pt(tab.pts + ec::PED_TABLE_POINTS);
    auto ld = [&](int idx) { return load_pt(tab.pts + idx); };
    JacPt s;
    s.x = fp::one(); s.z = fp::one();
    if (on) s = ec::pedersen_sum(load_fp(reinterpret_cast<const uint8_t *>(a + i)), load_fp(reinterpret_cast<const uint8_t *>(b + i)), p0, ld);
    if (ec::is_zero_mod_p(s.z)) s.z = fp::one();
    const Fp zi = ec::block_inverse<128>(s.z, inv_sm);
    if (on) store_fp(reinterpret_cast<uint8_t *>(out + i), fp::canon(fp::mul(s.x, fp::sqr(zi))));
}

__global__ void copy_column_kernel(const Fp *col, int log_rows, int bitrev, uint8_t *out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >> log_rows) return;
    const unsigned long long src = bitrev ? brev_bits(i, log_rows) : i;
    store_fp(out + 32ull * i, load_fp(reinterpret_cast<const uint8_t *>(col + src)));
}

// out[k] = (v < n ? leaves[v] : nodes[v - n]) for virtual index v
__global__ void gather32_kernel(const uint8_t *leaves, const uint8_t *nodes, unsigned long long n,
                                const unsigned long long *vidx, unsigned long long count, uint8_t *out) {
    const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (k >= count) return;
    const unsigned long long v = vidx[k];
    const uint8_t *src = v < n ? leaves + 32ull * v : nodes + 32ull * (v - n);
    const uint4 *q = reinterpret_cast<const uint4 *>(src);
    uint4 *o = reinterpret_cast<uint4 *>(out + 32ull * k);
    o[0] = q[0]; o[1] = q[1];
}

__global__ void rows_gather_kernel(const Fp *cols, unsigned long long stride, int n_cols, const unsigned long long *idx,
                                   unsigned long long count, Fp *out) {
    const unsigned long long k = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (k >= count * (unsigned long long)n_cols) return;
    const unsigned long long i = k / n_cols, j = k % n_cols;
    store_fp(reinterpret_cast<uint8_t *>(out + k), load_fp(reinterpret_cast<const uint8_t *>(cols + j * stride + idx[i])));
}

// ---------------------------------------------------------------------------------- host side
ss_status pedersen_table(ss_ctx *ctx, PedersenTable *out) {
    Fp *d;
    ss_status rc = cached_table(ctx, {T_PEDERSEN, 0, 0}, 2 * (size_t)(ec::PED_TABLE_POINTS + 1), ec::fill_pedersen_table, &d);
    if (rc) return rc;
    out->pts = reinterpret_cast<const AffinePt *>(d);
    return SS_OK;
}

inline unsigned grid_for(unsigned long long count, int block) { return (unsigned)((count + block - 1) / block); }

template <typename F>
ss_status by_byte_hash(int bh, F f) {
    switch (bh) {
    case BH_KECCAK: return f(std::integral_constant<int, BH_KECCAK>());
    case BH_BLAKE2S: return f(std::integral_constant<int, BH_BLAKE2S>());
    default: return f(std::integral_constant<int, BH_SHA256>());
    }
}

void byte_hash_of(int kind, int &bh, int &mask) {
    switch (kind) {
    case SS_TREE_KECCAK: bh = BH_KECCAK; mask = MASK_NONE; break;
    case SS_TREE_KECCAK_M20: bh = BH_KECCAK; mask = MASK_KEEP_FIRST20; break;
    case SS_TREE_FRIENDLY: case SS_TREE_BLAKE2S_M20: bh = BH_BLAKE2S; mask = MASK_KEEP_LAST20; break;
    default: bh = BH_SHA256; mask = MASK_NONE; break;
    }
}

void register_tree(ss_ctx *ctx, ss_tree *t) {
    ctx->live_trees.insert(t);
    ctx->detach_tree = [](ss_tree *tree) { tree->ctx = nullptr; tree->d_leaves = nullptr; tree->d_nodes = nullptr; };
}

// node levels above row digests (n_cols >= 2 trees): byte-hash levels FUSED_LEVELS at a time, Pedersen levels one by one
void build_node_levels(ss_ctx *ctx, ss_tree *t, int bh, int mask, const PedersenTable &tab, cudaStream_t st) {
    const int height = t->log_rows;
    const int transition = t->kind == SS_TREE_FRIENDLY ? t->n_friendly : 0;
    int d = height - 1;                                   // depth of the next level to build
    while (d >= 0) {
        const uint8_t *children = (d == height - 1) ? t->d_leaves : t->d_nodes + 64ull * (1ull << d);
        if (d >= transition) {
            const int left = d - transition + 1;          // byte-hash levels left
            const bool fused = ss::option(ctx, "merkle_fused", 1) != 0;
            if (fused && d + 1 <= FUSED_LEVELS && left >= 2) {
                // the top of the tree: one CTA walks all remaining levels in shared memory instead of one tiny launch each
                const int levels = left;
                by_byte_hash(bh, [&](auto BH) {
                    node_levels_fused_kernel<decltype(BH)::value><<<1u << (d + 1 - levels), 256, 0, st>>>(children, d + 1, levels, mask, t->d_nodes);
                    ctx->launches++;
                    return SS_OK;
                });
                d -= levels;
            } else if (fused && left >= 3 && d >= 2) {
                const unsigned long long c = 1ull << (d - 2);                  // nodes at depth d - 2
                by_byte_hash(bh, [&](auto BH) {
                    node_levels3_kernel<decltype(BH)::value><<<grid_for(c, 128), 128, 0, st>>>(children, d + 1, c, mask, t->d_nodes);
                    ctx->launches++;
                    return SS_OK;
                });
                d -= 3;
            } else {
                const unsigned long long c = 1ull << d;
                by_byte_hash(bh, [&](auto BH) {
                    node_hash_kernel<decltype(BH)::value><<<grid_for(c, 128), 128, 0, st>>>(children, c, mask, t->d_nodes + 32ull * c);
                    ctx->launches++;
                    return SS_OK;
                });
                --d;
            }
        } else {
            const unsigned long long c = 1ull << d;
            const bool child_high = (d + 1 < transition) && (d != height - 1);
            pedersen_node_kernel<<<grid_for(c, 128), 128, 0, st>>>(children, c, child_high ? 0 : 1, tab, t->d_nodes + 32ull * c);
            ctx->launches++;
            --d;
        }
    }
}

}  // namespace

extern "C" {

ss_status ss_merkle_build(ss_ctx *ctx, ss_tree_kind kind, int n_friendly, const void *d_cols, uint64_t col_stride,
                          int n_cols, int log_rows, ss_order row_order, ss_tree **out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (!out || !d_cols || n_cols < 1 || log_rows < 1 || log_rows > 40 || col_stride < (1ull << log_rows) ||
        (int)kind < 0 || (int)kind > SS_TREE_SHA256 || n_friendly < 0)
        return fail(ctx, SS_ERR_INVALID, "ss_merkle_build: bad arguments (n_cols=%d log_rows=%d)", n_cols, log_rows);
    *out = nullptr;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    const unsigned long long n = 1ull << log_rows;
    const Fp *cols = static_cast<const Fp *>(d_cols);
    const int bitrev = ((int)row_order & 1) ? 1 : 0;
    int col_bits = 0;
    if ((int)row_order & 2) {                               // SS_ORDER_BITREV_RC: columns in bit-reversed order too
        while ((1 << col_bits) < n_cols) ++col_bits;
        if ((1 << col_bits) != n_cols) return fail(ctx, SS_ERR_INVALID, "ss_merkle_build: bit-reversed column order needs a power-of-two column count");
    }
    int bh, mask;
    byte_hash_of(kind, bh, mask);
    PedersenTable tab{nullptr};
    if (kind == SS_TREE_FRIENDLY && (n_cols == 1 || n_friendly > 0)) {
        ss_status rc = pedersen_table(ctx, &tab);
        if (rc) return rc;
    }
    ss_tree *t = new ss_tree{ctx, (int)kind, n_friendly, log_rows, n_cols, nullptr, nullptr};
    cudaError_t e1 = dev_alloc(ctx, reinterpret_cast<void **>(&t->d_leaves), n * 32), e2 = dev_alloc(ctx, reinterpret_cast<void **>(&t->d_nodes), n * 32);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        dev_free(ctx, t->d_leaves); dev_free(ctx, t->d_nodes); delete t;
        cudaGetLastError();
        return fail(ctx, SS_ERR_OOM, "ss_merkle_build: cannot allocate %llu bytes for the tree", n * 64ull);
    }
    SS_CUDA_CHECK(ctx, cudaMemsetAsync(t->d_nodes, 0, 32, st));
    const int height = log_rows;
    ss_status rc = SS_OK;
    if (n_cols == 1) {
        copy_column_kernel<<<grid_for(n, 256), 256, 0, st>>>(cols, log_rows, bitrev, t->d_leaves);
        ctx->launches++;
        const unsigned long long cnt = n / 2;
        uint8_t *lvl = t->d_nodes + 32ull * cnt;
        if (kind == SS_TREE_FRIENDLY) {
            pedersen_node_kernel<<<grid_for(cnt, 128), 128, 0, st>>>(t->d_leaves, cnt, 2, tab, lvl);
            ctx->launches++;
        } else {
            rc = by_byte_hash(bh, [&](auto BH) {
                leafpair_hash_kernel<decltype(BH)::value><<<grid_for(cnt, 128), 128, 0, st>>>(reinterpret_cast<const Fp *>(t->d_leaves), cnt, mask, lvl);
                ctx->launches++;
                return SS_OK;
            });
        }
        for (int d = height - 2; d >= 0; --d) {
            const unsigned long long c = 1ull << d;
            const uint8_t *children = t->d_nodes + 64ull * c;
            uint8_t *dst = t->d_nodes + 32ull * c;
            if (kind == SS_TREE_FRIENDLY) {
                pedersen_node_kernel<<<grid_for(c, 128), 128, 0, st>>>(children, c, 0, tab, dst);
                ctx->launches++;
            } else {
                by_byte_hash(bh, [&](auto BH) {
                    node_hash_kernel<decltype(BH)::value><<<grid_for(c, 128), 128, 0, st>>>(children, c, mask, dst);
                    ctx->launches++;
                    return SS_OK;
                });
            }
        }
    } else {
        by_byte_hash(bh, [&](auto BH) {
            leaf_hash_kernel<decltype(BH)::value><<<grid_for(n, 128), 128, 0, st>>>(cols, col_stride, n_cols, log_rows, bitrev, col_bits, mask, t->d_leaves);
            ctx->launches++;
            return SS_OK;
        });
        build_node_levels(ctx, t, bh, mask, tab, st);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        dev_free(ctx, t->d_leaves); dev_free(ctx, t->d_nodes); delete t;
        return fail(ctx, SS_ERR_CUDA, "ss_merkle_build: launch failed: %s", cudaGetErrorString(e));
    }
    (void)rc;
    register_tree(ctx, t);
    *out = t;
    return SS_OK;
}

ss_status ss_hash_rows(ss_ctx *ctx, ss_tree_kind kind, const void *d_cols, uint64_t col_stride, int n_cols, int log_rows,
                       ss_order order, uint64_t leaf_begin, uint64_t leaf_count, void *d_out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (!d_cols || !d_out || n_cols < 2 || log_rows < 0 || log_rows > 40 || col_stride < (1ull << log_rows) || (int)kind < 0 ||
        (int)kind > SS_TREE_SHA256 || leaf_begin + leaf_count > (1ull << log_rows))
        return fail(ctx, SS_ERR_INVALID, "ss_hash_rows: bad arguments");
    if (leaf_count == 0) return SS_OK;
    int col_bits = 0;
    if ((int)order & 2) {
        while ((1 << col_bits) < n_cols) ++col_bits;
        if ((1 << col_bits) != n_cols) return fail(ctx, SS_ERR_INVALID, "ss_hash_rows: bit-reversed column order needs a power-of-two column count");
    }
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    int bh, mask;
    byte_hash_of(kind, bh, mask);
    by_byte_hash(bh, [&](auto BH) {
        leaf_range_hash_kernel<decltype(BH)::value><<<grid_for(leaf_count, 128), 128, 0, pick_stream(ctx, stream)>>>(
            static_cast<const Fp *>(d_cols), col_stride, n_cols, log_rows, (int)order & 1, col_bits, mask, leaf_begin, leaf_count, static_cast<uint8_t *>(d_out));
        ctx->launches++;
        return SS_OK;
    });
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

ss_status ss_merkle_build_from_leaves(ss_ctx *ctx, ss_tree_kind kind, int n_friendly, const void *d_leaf_digests, int log_rows,
                                      ss_tree **out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (!out || !d_leaf_digests || log_rows < 1 || log_rows > 40 || (int)kind < 0 || (int)kind > SS_TREE_SHA256 || n_friendly < 0)
        return fail(ctx, SS_ERR_INVALID, "ss_merkle_build_from_leaves: bad arguments");
    *out = nullptr;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    const unsigned long long n = 1ull << log_rows;
    int bh, mask;
    byte_hash_of(kind, bh, mask);
    PedersenTable tab{nullptr};
    if (kind == SS_TREE_FRIENDLY && n_friendly > 0) {
        ss_status rc = pedersen_table(ctx, &tab);
        if (rc) return rc;
    }
    ss_tree *t = new ss_tree{ctx, (int)kind, n_friendly, log_rows, 2, nullptr, nullptr};
    cudaError_t e1 = dev_alloc(ctx, reinterpret_cast<void **>(&t->d_leaves), n * 32), e2 = dev_alloc(ctx, reinterpret_cast<void **>(&t->d_nodes), n * 32);
    if (e1 != cudaSuccess || e2 != cudaSuccess) {
        dev_free(ctx, t->d_leaves); dev_free(ctx, t->d_nodes); delete t;
        cudaGetLastError();
        return fail(ctx, SS_ERR_OOM, "ss_merkle_build_from_leaves: cannot allocate %llu bytes for the tree", n * 64ull);
    }
    cudaMemsetAsync(t->d_nodes, 0, 32, st);
    cudaMemcpyAsync(t->d_leaves, d_leaf_digests, n * 32, cudaMemcpyDeviceToDevice, st);
    build_node_levels(ctx, t, bh, mask, tab, st);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        dev_free(ctx, t->d_leaves); dev_free(ctx, t->d_nodes); delete t;
        return fail(ctx, SS_ERR_CUDA, "ss_merkle_build_from_leaves: launch failed: %s", cudaGetErrorString(e));
    }
    register_tree(ctx, t);
    *out = t;
    return SS_OK;
}

// In-place bit-reversal permutation of 2^log_n 32-byte items (leaf digests moving between row order and tree order)
__global__ void bitrev32_kernel(uint4 *items, int log_n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    const unsigned long long j = log_n ? (__brevll(i) >> (64 - log_n)) : 0ull;
    if (i < j) {
        const uint4 a0 = items[2 * i], a1 = items[2 * i + 1], b0 = items[2 * j], b1 = items[2 * j + 1];
        items[2 * i] = b0; items[2 * i + 1] = b1; items[2 * j] = a0; items[2 * j + 1] = a1;
    }
}
ss_status ss_bitrev_permute32(ss_ctx *ctx, void *d_items, int log_n, void *stream) {
    if (!ctx || !d_items || log_n < 0 || log_n > 40) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    bitrev32_kernel<<<grid_for(1ull << log_n, 256), 256, 0, pick_stream(ctx, stream)>>>(static_cast<uint4 *>(d_items), log_n);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

// PedersenDigest::as_bytes: big-endian canonical integer (crypto/src/hash/pedersen.rs:23-28) of a stored felt
static void felt_to_be_bytes(const uint8_t raw[32], uint8_t out[32]) {
    Fp m, one_int = fp::zero();
    one_int.l[0] = 1;
    for (int i = 0; i < 8; ++i) m.l[i] = (uint32_t)raw[4 * i] | ((uint32_t)raw[4 * i + 1] << 8) | ((uint32_t)raw[4 * i + 2] << 16) | ((uint32_t)raw[4 * i + 3] << 24);
    const Fp c = fp::canon(fp::mul(m, one_int));
    for (int w = 0; w < 8; ++w) {
        const uint32_t v = c.l[7 - w];
        out[4 * w] = (uint8_t)(v >> 24); out[4 * w + 1] = (uint8_t)(v >> 16); out[4 * w + 2] = (uint8_t)(v >> 8); out[4 * w + 3] = (uint8_t)v;
    }
}

ss_status ss_merkle_root(ss_ctx *ctx, const ss_tree *tree, uint8_t root[32]) {
    if (!ctx || !tree || !root) return SS_ERR_INVALID;
    uint8_t raw[32];
    SS_CUDA_CHECK(ctx, cudaDeviceSynchronize());
    SS_CUDA_CHECK(ctx, cudaMemcpy(raw, tree->d_nodes + 32, 32, cudaMemcpyDeviceToHost));
    const bool algebraic = tree->kind == SS_TREE_FRIENDLY && (tree->n_cols == 1 || tree->n_friendly > 0);
    if (!algebraic) {
        for (int i = 0; i < 32; ++i) root[i] = raw[i];
        return SS_OK;
    }
    felt_to_be_bytes(raw, root);
    return SS_OK;
}

static ss_status gather_virtual(ss_ctx *ctx, const ss_tree *tree, const std::vector<unsigned long long> &v, uint8_t *h_out) {
    if (v.empty()) return SS_OK;
    unsigned long long *d_idx = nullptr;
    uint8_t *d_out = nullptr;
    SS_CUDA_CHECK(ctx, cudaMalloc(&d_idx, v.size() * 8));
    cudaError_t e = cudaMalloc(&d_out, v.size() * 32);
    if (e != cudaSuccess) { cudaFree(d_idx); return fail(ctx, SS_ERR_OOM, "gather: out of memory"); }
    cudaMemcpy(d_idx, v.data(), v.size() * 8, cudaMemcpyHostToDevice);
    gather32_kernel<<<grid_for(v.size(), 128), 128>>>(tree->d_leaves, tree->d_nodes, 1ull << tree->log_rows, d_idx, v.size(), d_out);
    tree->ctx->launches++;
    e = cudaMemcpy(h_out, d_out, v.size() * 32, cudaMemcpyDeviceToHost);
    cudaFree(d_idx); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "gather: %s", cudaGetErrorString(e));
    return SS_OK;
}

ss_status ss_merkle_nodes(ss_ctx *ctx, const ss_tree *tree, const uint64_t *h_indices, size_t n, uint8_t *h_out) {
    if (!ctx || !tree || (!h_indices && n) || (!h_out && n)) return SS_ERR_INVALID;
    const unsigned long long N = 1ull << tree->log_rows;
    std::vector<unsigned long long> v(n);
    for (size_t i = 0; i < n; ++i) {
        if (h_indices[i] < 1 || h_indices[i] >= N) return fail(ctx, SS_ERR_INVALID, "ss_merkle_nodes: index %llu out of range", (unsigned long long)h_indices[i]);
        v[i] = N + h_indices[i];
    }
    return gather_virtual(ctx, tree, v, h_out);
}

ss_status ss_merkle_leaves(ss_ctx *ctx, const ss_tree *tree, const uint64_t *h_indices, size_t n, uint8_t *h_out) {
    if (!ctx || !tree || (!h_indices && n) || (!h_out && n)) return SS_ERR_INVALID;
    const unsigned long long N = 1ull << tree->log_rows;
    std::vector<unsigned long long> v(n);
    for (size_t i = 0; i < n; ++i) {
        if (h_indices[i] >= N) return fail(ctx, SS_ERR_INVALID, "ss_merkle_leaves: index %llu out of range", (unsigned long long)h_indices[i]);
        v[i] = h_indices[i];
    }
    return gather_virtual(ctx, tree, v, h_out);
}

ss_status ss_merkle_open(ss_ctx *ctx, const ss_tree *tree, const uint64_t *h_indices, size_t n, uint8_t *h_paths) {
    if (!ctx || !tree || (!h_indices && n) || (!h_paths && n)) return SS_ERR_INVALID;
    const unsigned long long N = 1ull << tree->log_rows;
    std::vector<unsigned long long> v;
    v.reserve(n * tree->log_rows);
    for (size_t i = 0; i < n; ++i) {
        const unsigned long long idx = h_indices[i];
        if (idx >= N) return fail(ctx, SS_ERR_INVALID, "ss_merkle_open: index %llu out of range", idx);
        v.push_back(idx ^ 1ull);                          // sibling leaf
        for (unsigned long long pos = (N + idx) >> 1; pos > 1; pos >>= 1) v.push_back(N + (pos ^ 1ull));
    }
    return gather_virtual(ctx, tree, v, h_paths);
}

// Top of a row-sharded tree: `count` = 2^log_count sub-tree roots (as ss_merkle_root returns them) -> the nodes of the
// tree whose leaves they are, on the device: d[count .. 2 count) = the sub-roots as given, d[c .. 2c) the level of c nodes
// in storage form, d[1] the root.  Used when each GPU commits a row range.  The caller dev_free()s *out.
static ss_status combine_levels(ss_ctx *ctx, ss_tree_kind kind, const uint8_t *h_subroots, int log_count, uint8_t **out) {
    const unsigned long long count = 1ull << log_count;
    int bh, mask;
    byte_hash_of(kind, bh, mask);
    PedersenTable tab{nullptr};
    if (kind == SS_TREE_FRIENDLY) {
        // The top log_count levels of a friendly tree are algebraic as long as log_count <= N_FRIENDLY (22): the
        // sub-roots arrive as ss_merkle_root returns them (big-endian bytes: a masked digest or a canonical felt,
        // both below p), so the first level is hash_boundary + Pedersen and the rest plain Pedersen merges.
        ss_status rc = pedersen_table(ctx, &tab);
        if (rc) return rc;
    }
    uint8_t *d = nullptr;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&d), 2 * count * 32));
    cudaMemcpy(d + 32 * count, h_subroots, count * 32, cudaMemcpyHostToDevice);
    for (int lvl = log_count - 1; lvl >= 0; --lvl) {
        const unsigned long long c = 1ull << lvl;
        if (kind == SS_TREE_FRIENDLY) {
            pedersen_node_kernel<<<grid_for(c, 128), 128>>>(d + 64ull * c, c, lvl == log_count - 1 ? 1 : 0, tab, d + 32ull * c);
            ctx->launches++;
        } else {
            by_byte_hash(bh, [&](auto BH) {
                node_hash_kernel<decltype(BH)::value><<<grid_for(c, 128), 128>>>(d + 64ull * c, c, mask, d + 32ull * c);
                ctx->launches++;
                return SS_OK;
            });
        }
    }
    *out = d;
    return SS_OK;
}

ss_status ss_merkle_combine(ss_ctx *ctx, ss_tree_kind kind, const uint8_t *h_subroots, int log_count, uint8_t root[32]) {
    if (!ctx || !h_subroots || !root || log_count < 0 || log_count > 16) return SS_ERR_INVALID;
    if (log_count == 0) { for (int i = 0; i < 32; ++i) root[i] = h_subroots[i]; return SS_OK; }
    uint8_t *d = nullptr;
    ss_status rc = combine_levels(ctx, kind, h_subroots, log_count, &d);
    if (rc) return rc;
    uint8_t raw[32];
    cudaError_t e = cudaMemcpy(raw, d + 32, 32, cudaMemcpyDeviceToHost);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_merkle_combine: %s", cudaGetErrorString(e));
    if (kind == SS_TREE_FRIENDLY) felt_to_be_bytes(raw, root);
    else for (int i = 0; i < 32; ++i) root[i] = raw[i];
    return SS_OK;
}

// The upper part of an authentication path through a row-sharded tree: the log_count siblings above sub-tree `index`, bottom
// first, in the storage form ss_merkle_open uses (so that  ss_merkle_open(sub-tree) ++ this  is the path of the whole tree).
ss_status ss_merkle_combine_open(ss_ctx *ctx, ss_tree_kind kind, const uint8_t *h_subroots, int log_count, int subroots_algebraic,
                                 uint64_t index, uint8_t *h_path) {
    if (!ctx || !h_subroots || (!h_path && log_count) || log_count < 0 || log_count > 16 || index >> log_count) return SS_ERR_INVALID;
    if (log_count == 0) return SS_OK;
    const unsigned long long count = 1ull << log_count;
    uint8_t *d = nullptr;
    ss_status rc = combine_levels(ctx, kind, h_subroots, log_count, &d);
    if (rc) return rc;
    std::vector<uint8_t> nodes(2 * count * 32);
    cudaError_t e = cudaMemcpy(nodes.data(), d, nodes.size(), cudaMemcpyDeviceToHost);
    dev_free(ctx, d);
    if (e != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_merkle_combine_open: %s", cudaGetErrorString(e));
    unsigned long long pos = count + index;
    for (int k = 0; k < log_count; ++k, pos >>= 1) {
        const uint8_t *sib = nodes.data() + 32 * (pos ^ 1ull);
        uint8_t *out = h_path + 32 * k;
        if (k == 0 && kind == SS_TREE_FRIENDLY && subroots_algebraic) {
            // a Pedersen sub-root came in as its big-endian canonical integer: back to Montgomery limbs
            Fp v;
            for (int w = 0; w < 8; ++w)
                v.l[7 - w] = ((uint32_t)sib[4 * w] << 24) | ((uint32_t)sib[4 * w + 1] << 16) | ((uint32_t)sib[4 * w + 2] << 8) | (uint32_t)sib[4 * w + 3];
            const Fp m = fp::canon(fp::mul(v, fp::r2()));
            for (int i = 0; i < 8; ++i) {
                out[4 * i] = (uint8_t)m.l[i]; out[4 * i + 1] = (uint8_t)(m.l[i] >> 8); out[4 * i + 2] = (uint8_t)(m.l[i] >> 16); out[4 * i + 3] = (uint8_t)(m.l[i] >> 24);
            }
        } else {
            memcpy(out, sib, 32);
        }
    }
    return SS_OK;
}

int ss_tree_log_rows(const ss_tree *tree) { return tree ? tree->log_rows : -1; }

void ss_tree_free(ss_tree *tree) {
    if (!tree) return;
    if (!tree->ctx) { delete tree; return; }                 // the context was destroyed first: its memory is already gone
    tree->ctx->live_trees.erase(tree);
    cudaSetDevice(tree->ctx->device);
    dev_free(tree->ctx, tree->d_leaves);
    dev_free(tree->ctx, tree->d_nodes);
    delete tree;
}

ss_status ss_pedersen_hash(ss_ctx *ctx, const void *d_a, const void *d_b, void *d_out, size_t n, void *stream) {
    if (!ctx || (n && (!d_a || !d_b || !d_out))) return SS_ERR_INVALID;
    if (n == 0) return SS_OK;
    PedersenTable tab;
    ss_status rc = pedersen_table(ctx, &tab);
    if (rc) return rc;
    pedersen_batch_kernel<<<grid_for(n, 128), 128, 0, pick_stream(ctx, stream)>>>(static_cast<const Fp *>(d_a), static_cast<const Fp *>(d_b), static_cast<Fp *>(d_out), n, tab);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

ss_status ss_rows_gather(ss_ctx *ctx, const void *d_cols, uint64_t col_stride, int n_cols, const uint64_t *h_indices,
                         size_t n, void *h_rows) {
    if (!ctx || !d_cols || n_cols < 1 || (n && (!h_indices || !h_rows))) return SS_ERR_INVALID;
    if (n == 0) return SS_OK;
    unsigned long long *d_idx = nullptr;
    Fp *d_out = nullptr;
    SS_CUDA_CHECK(ctx, cudaMalloc(&d_idx, n * 8));
    cudaError_t e = cudaMalloc(&d_out, n * n_cols * sizeof(Fp));
    if (e != cudaSuccess) { cudaFree(d_idx); return fail(ctx, SS_ERR_OOM, "ss_rows_gather: out of memory"); }
    cudaMemcpy(d_idx, h_indices, n * 8, cudaMemcpyHostToDevice);
    rows_gather_kernel<<<grid_for(n * n_cols, 128), 128>>>(static_cast<const Fp *>(d_cols), col_stride, n_cols, d_idx, n, d_out);
    ctx->launches++;
    e = cudaMemcpy(h_rows, d_out, n * n_cols * sizeof(Fp), cudaMemcpyDeviceToHost);
    cudaFree(d_idx); cudaFree(d_out);
    if (e != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_rows_gather: %s", cudaGetErrorString(e));
    return SS_OK;
}

}  // extern "C"
