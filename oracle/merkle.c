/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header).
 *
 * Restates the reference's Merkle commitment of a column-major matrix:
 *   - variant selection        crypto/src/merkle/mod.rs:110-123 (Friendly), :289-304 (LeafVariant)
 *   - row hashing              crypto/src/merkle/utils.rs:9-46
 *   - level / hash selection   crypto/src/merkle/mixed.rs:110-125, boundary :148-155
 *   - single-column first level crypto/src/merkle/mod.rs:422-437 (UnhashedLeafConfig)
 * The node array layout and the `depth` argument come from ministark's
 * `MerkleTreeImpl::new` (NOT in /root/reference; [RECALLED], parity unpinned):
 * nodes[1] is the root, nodes[i] = H(nodes[2i], nodes[2i+1]) with depth = floor(log2 i),
 * nodes[n/2 + i] = hash_leaves(depth = log2(n) - 1, leaf[2i], leaf[2i+1]).
 *
 * Storage form of a node (32 bytes): byte digests as produced by the hash;
 * algebraic (Pedersen) digests as the 4 x u64 LE Montgomery limbs of the felt.
 */
#include "hash.h"
#include <string.h>
#include <stdlib.h>

void oracle_pedersen_hash_elements(fp_t *r, const fp_t *elems, size_t n);

enum {
    TREE_KECCAK = 0,          /* LeafVariantMerkleTree<Keccak256HashFn>            src/claims.rs:29-30 */
    TREE_KECCAK_M20 = 1,      /* LeafVariantMerkleTree<MaskedKeccak256HashFn<20>>  src/claims.rs:18-19 */
    TREE_FRIENDLY = 2,        /* FriendlyMerkleTree<N, PedersenHashFn>             src/claims.rs:20-21,31-32 */
    TREE_BLAKE2S_M20 = 3,     /* LeafVariant over MaskedBlake2s (not a reference claim; Friendly with N=0 on >=2 cols) */
    TREE_SHA256 = 4
};

static size_t bitrev_sz(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

static int byte_hash_of(int kind) {
    switch (kind) {
    case TREE_KECCAK: return ORACLE_HASH_KECCAK;
    case TREE_KECCAK_M20: return ORACLE_HASH_KECCAK_M20;
    case TREE_FRIENDLY: case TREE_BLAKE2S_M20: return ORACLE_HASH_BLAKE2S_M20;
    default: return ORACLE_HASH_SHA256;
    }
}

/* utils.rs:19-46 `hash_rows`: digest_i = H(BE32(m[0][i]) || BE32(m[1][i]) || ...) */
void oracle_hash_rows(int hash_kind, const fp_t *cols, int n_cols, size_t n_rows, int bitrev_rows,
                      uint8_t *out /* n_rows * 32 */) {
    int bits = 0;
    while (((size_t)1 << bits) < n_rows) ++bits;
    #pragma omp parallel
    {
        uint8_t *buf = (uint8_t *)malloc((size_t)n_cols * 32);
        #pragma omp for schedule(static)
        for (size_t i = 0; i < n_rows; ++i) {
            const size_t src = bitrev_rows ? bitrev_sz(i, bits) : i;
            for (int j = 0; j < n_cols; ++j) oracle_felt_to_be32(&cols[(size_t)j * n_rows + src], buf + 32 * j);
            oracle_hash_bytes(hash_kind, buf, (size_t)n_cols * 32, out + 32 * i);
        }
        free(buf);
    }
}

/* mixed.rs:148-155: digest bytes -> BE integer -> Fp */
static void digest_to_felt(fp_t *r, const uint8_t d[32]) {
    fp_t c;
    for (int i = 0; i < 4; ++i) {
        uint64_t limb = 0;
        for (int b = 0; b < 8; ++b) limb = (limb << 8) | d[8 * (3 - i) + b];
        c.l[i] = limb;
    }
    /* Fp::from(BigUint) reduces mod p; masked-20 digests are < 2^160 so this never
     * fires on the reference's path, but keep the semantics exact. */
    fp_to_mont(r, &c);   /* Montgomery multiplication by R^2 reduces any 256-bit input */
}

static void merge_bytes(int hash_kind, const uint8_t *a, const uint8_t *b, uint8_t *out) {
    uint8_t buf[64];
    memcpy(buf, a, 32);
    memcpy(buf + 32, b, 32);
    oracle_hash_bytes(hash_kind, buf, 64, out);
}

/* Builds the full node array. nodes: n*32 bytes (slot 0 unused, zeroed).
 * leaves: n*32 bytes, receives the row hashes (n_cols >= 2) or the raw column (n_cols == 1).
 * Returns 0 on success. */
int oracle_merkle_build(int kind, int n_friendly, const fp_t *cols, int n_cols, int log_rows,
                        int bitrev_rows, uint8_t *nodes, uint8_t *leaves) {
    const size_t n = (size_t)1 << log_rows;
    if (log_rows < 1 || n_cols < 1) return -1;
    const int hk = byte_hash_of(kind);
    const int height = log_rows;
    memset(nodes, 0, 32);
    if (n_cols == 1) {
        /* raw leaves: mod.rs:113-116 / :292-295 */
        for (size_t i = 0; i < n; ++i) {
            const size_t src = bitrev_rows ? bitrev_sz(i, log_rows) : i;
            memcpy(leaves + 32 * i, &cols[src], 32);
        }
        #pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n / 2; ++i) {
            const fp_t *l = (const fp_t *)(leaves + 64 * i);
            uint8_t *dst = nodes + 32 * (n / 2 + i);
            if (kind == TREE_FRIENDLY) {
                fp_t pair[2], h;
                memcpy(pair, l, 64);
                oracle_pedersen_hash_elements(&h, pair, 2);       /* mod.rs:426-428 with H = Pedersen */
                memcpy(dst, &h, 32);
            } else {
                uint8_t buf[64];
                fp_t pair[2];
                memcpy(pair, l, 64);
                oracle_felt_to_be32(&pair[0], buf);
                oracle_felt_to_be32(&pair[1], buf + 32);
                oracle_hash_bytes(hk, buf, 64, dst);               /* H::hash_elements([l0,l1]) */
            }
        }
        for (int d = height - 2; d >= 0; --d) {
            const size_t lo = (size_t)1 << d;
            #pragma omp parallel for schedule(static)
            for (size_t i = lo; i < 2 * lo; ++i) {
                if (kind == TREE_FRIENDLY) {
                    fp_t a, b, h;
                    memcpy(&a, nodes + 64 * i, 32);
                    memcpy(&b, nodes + 64 * i + 32, 32);
                    oracle_pedersen_hash(&h, &a, &b);
                    memcpy(nodes + 32 * i, &h, 32);
                } else {
                    merge_bytes(hk, nodes + 64 * i, nodes + 64 * i + 32, nodes + 32 * i);
                }
            }
        }
        return 0;
    }
    oracle_hash_rows(hk, cols, n_cols, n, bitrev_rows, leaves);
    const int transition = (kind == TREE_FRIENDLY) ? n_friendly : 0;
    for (int d = height - 1; d >= 0; --d) {
        const size_t lo = (size_t)1 << d;
        const uint8_t *child = (d == height - 1) ? leaves : nodes;
        const size_t child_base = (d == height - 1) ? 0 : 2 * lo;
        /* children of node i (i in [lo,2lo)) are child[(2*(i-lo)) + child_base], +1 */
        const int high = d < transition;
        const int child_high = (d + 1 < transition) && (d != height - 1);
        #pragma omp parallel for schedule(static)
        for (size_t i = lo; i < 2 * lo; ++i) {
            const uint8_t *c0 = child + 32 * (child_base + 2 * (i - lo));
            const uint8_t *c1 = c0 + 32;
            uint8_t *dst = nodes + 32 * i;
            if (!high) {
                merge_bytes(hk, c0, c1, dst);                                 /* mixed.rs:113,120 */
            } else {
                fp_t a, b, h;
                if (child_high) { memcpy(&a, c0, 32); memcpy(&b, c1, 32); }   /* mixed.rs:122 */
                else { digest_to_felt(&a, c0); digest_to_felt(&b, c1); }      /* mixed.rs:112,121 */
                oracle_pedersen_hash(&h, &a, &b);
                memcpy(dst, &h, 32);
            }
        }
    }
    return 0;
}

/* Digest::as_bytes of the root: byte digests verbatim; PedersenDigest = BE32 of the
 * canonical integer (crypto/src/hash/pedersen.rs:23-28). */
void oracle_merkle_root_bytes(int kind, int n_friendly, int n_cols, int log_rows, const uint8_t *nodes, uint8_t out[32]) {
    const int algebraic = (kind == TREE_FRIENDLY) && (n_cols == 1 || n_friendly > 0);
    (void)log_rows;
    if (!algebraic) { memcpy(out, nodes + 32, 32); return; }
    fp_t m, c;
    memcpy(&m, nodes + 32, 32);
    fp_from_mont(&c, &m);
    oracle_felt_to_be32(&c, out);
}

/* Test helper for the reference's proof artefacts (tests/test_reference_proof.py): the leaf index of a Merkle opening is
 * not part of a serialized proof (the verifier derives it from the public coin); find it by trying both child orders at
 * every level.  leaf / siblings: 32-byte digests (siblings[0] = sibling leaf, then the path, leaf level first).
 * Returns the index whose recomputed root equals `root`, or -1. */
long long oracle_merkle_find_index(int hash_kind, const uint8_t *leaf, const uint8_t *siblings, int depth, const uint8_t *root) {
    size_t count = 1;
    uint8_t *cur = (uint8_t *)malloc(32), *nxt;
    memcpy(cur, leaf, 32);
    for (int lvl = 0; lvl < depth; ++lvl) {
        nxt = (uint8_t *)malloc(count * 2 * 32);
        const uint8_t *s = siblings + 32 * lvl;
        #pragma omp parallel for schedule(static) if (count >= 1024)
        for (size_t i = 0; i < count; ++i) {
            merge_bytes(hash_kind, cur + 32 * i, s, nxt + 32 * i);                    /* index bit lvl = 0: we are the left child */
            merge_bytes(hash_kind, s, cur + 32 * i, nxt + 32 * (i + count));          /* index bit lvl = 1 */
        }
        free(cur);
        cur = nxt;
        count *= 2;
    }
    long long found = -1;
    for (size_t i = 0; i < count; ++i)
        if (!memcmp(cur + 32 * i, root, 32)) { found = (long long)i; break; }
    free(cur);
    return found;
}
