/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header).
 *
 * Keccak-256 (legacy 0x01 padding, rate 136; crate sha3 0.10.8 `Keccak256`,
 * Cargo.lock:1533) and Blake2s-256 (crate blake2 0.10.6, Cargo.lock:308) as the
 * reference's HashFn impls use them: crypto/src/hash/keccak.rs:13-58,
 * crypto/src/hash/blake2s.rs:10-61; masks crypto/src/hash/mod.rs:5-23.
 * Pinned by python hashlib (blake2s) and the Solidity-coin KAT
 * (crypto/src/public_coin/solidity.rs:173-192) in tests/.
 */
#include "hash.h"
#include <string.h>

/* ------------------------------------------------------------ Keccak-f[1600] */
static const uint64_t KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[24] = {1, 3, 6, 10, 15, 21, 28, 36, 45, 55, 2, 14,
                                   27, 41, 56, 8, 25, 43, 62, 18, 39, 61, 20, 44};
static const int KECCAK_PIL[24] = {10, 7, 11, 17, 18, 3, 5, 16, 8, 21, 24, 4,
                                   15, 23, 19, 13, 12, 2, 20, 14, 22, 9, 6, 1};

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static void keccak_f(uint64_t st[25]) {
    for (int round = 0; round < 24; ++round) {
        uint64_t bc[5];
        for (int i = 0; i < 5; ++i) bc[i] = st[i] ^ st[i + 5] ^ st[i + 10] ^ st[i + 15] ^ st[i + 20];
        for (int i = 0; i < 5; ++i) {
            uint64_t t = bc[(i + 4) % 5] ^ rotl64(bc[(i + 1) % 5], 1);
            for (int j = 0; j < 25; j += 5) st[j + i] ^= t;
        }
        uint64_t t = st[1];
        for (int i = 0; i < 24; ++i) {
            int j = KECCAK_PIL[i];
            uint64_t b = st[j];
            st[j] = rotl64(t, KECCAK_ROT[i]);
            t = b;
        }
        for (int j = 0; j < 25; j += 5) {
            for (int i = 0; i < 5; ++i) bc[i] = st[j + i];
            for (int i = 0; i < 5; ++i) st[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
        }
        st[0] ^= KECCAK_RC[round];
    }
}

void oracle_keccak256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint64_t st[25];
    memset(st, 0, sizeof st);
    const size_t rate = 136;
    while (len >= rate) {
        for (size_t i = 0; i < rate / 8; ++i) {
            uint64_t w;
            memcpy(&w, in + 8 * i, 8);
            st[i] ^= w;
        }
        keccak_f(st);
        in += rate;
        len -= rate;
    }
    uint8_t blk[136];
    memset(blk, 0, sizeof blk);
    memcpy(blk, in, len);
    blk[len] ^= 0x01;
    blk[rate - 1] ^= 0x80;
    for (size_t i = 0; i < rate / 8; ++i) {
        uint64_t w;
        memcpy(&w, blk + 8 * i, 8);
        st[i] ^= w;
    }
    keccak_f(st);
    memcpy(out, st, 32);
}

/* ------------------------------------------------------------------ Blake2s */
static const uint32_t B2S_IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                                   0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t B2S_SIGMA[10][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3},
    {11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4}, {7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8},
    {9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13}, {2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9},
    {12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11}, {13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10},
    {6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5}, {10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0}};

static inline uint32_t rotr32(uint32_t x, int r) { return (x >> r) | (x << (32 - r)); }

static void b2s_compress(uint32_t h[8], const uint8_t block[64], uint64_t t, int last) {
    uint32_t m[16], v[16];
    for (int i = 0; i < 16; ++i) memcpy(&m[i], block + 4 * i, 4);
    for (int i = 0; i < 8; ++i) { v[i] = h[i]; v[i + 8] = B2S_IV[i]; }
    v[12] ^= (uint32_t)t;
    v[13] ^= (uint32_t)(t >> 32);
    if (last) v[14] = ~v[14];
#define G(a, b, c, d, x, y)                                        \
    v[a] = v[a] + v[b] + (x); v[d] = rotr32(v[d] ^ v[a], 16);      \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 12);      \
    v[a] = v[a] + v[b] + (y); v[d] = rotr32(v[d] ^ v[a], 8);       \
    v[c] = v[c] + v[d];       v[b] = rotr32(v[b] ^ v[c], 7);
    for (int r = 0; r < 10; ++r) {
        const uint8_t *s = B2S_SIGMA[r];
        G(0, 4, 8, 12, m[s[0]], m[s[1]]);
        G(1, 5, 9, 13, m[s[2]], m[s[3]]);
        G(2, 6, 10, 14, m[s[4]], m[s[5]]);
        G(3, 7, 11, 15, m[s[6]], m[s[7]]);
        G(0, 5, 10, 15, m[s[8]], m[s[9]]);
        G(1, 6, 11, 12, m[s[10]], m[s[11]]);
        G(2, 7, 8, 13, m[s[12]], m[s[13]]);
        G(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
#undef G
    for (int i = 0; i < 8; ++i) h[i] ^= v[i] ^ v[i + 8];
}

void oracle_blake2s256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) h[i] = B2S_IV[i];
    h[0] ^= 0x01010020u;    /* digest 32, no key, fanout 1, depth 1 */
    uint64_t t = 0;
    while (len > 64) {
        t += 64;
        b2s_compress(h, in, t, 0);
        in += 64;
        len -= 64;
    }
    uint8_t blk[64];
    memset(blk, 0, sizeof blk);
    memcpy(blk, in, len);
    t += len;
    b2s_compress(h, blk, t, 1);
    memcpy(out, h, 32);
}

/* ------------------------------------------------------------------ SHA-256 */
static const uint32_t SHA_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

static void sha256_block(uint32_t h[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; ++i)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; ++i) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; ++i) {
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA_K[i] + w[i];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

void oracle_sha256(const uint8_t *in, size_t len, uint8_t out[32]) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const uint64_t bitlen = (uint64_t)len * 8;
    while (len >= 64) { sha256_block(h, in); in += 64; len -= 64; }
    uint8_t blk[128];
    memset(blk, 0, sizeof blk);
    memcpy(blk, in, len);
    blk[len] = 0x80;
    size_t total = (len + 9 <= 64) ? 64 : 128;
    for (int i = 0; i < 8; ++i) blk[total - 1 - i] = (uint8_t)(bitlen >> (8 * i));
    sha256_block(h, blk);
    if (total == 128) sha256_block(h, blk + 64);
    for (int i = 0; i < 8; ++i) {
        out[4 * i] = (uint8_t)(h[i] >> 24); out[4 * i + 1] = (uint8_t)(h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(h[i] >> 8); out[4 * i + 3] = (uint8_t)h[i];
    }
}

/* -------------------------------------------------- element (row) encodings */
/* 32-byte big-endian of the Montgomery limbs: `to_montgomery(e).to_be_bytes::<32>()`
 * (crypto/src/hash/keccak.rs:54, blake2s.rs:54, crypto/src/utils.rs:15-17). */
void oracle_felt_to_be32(const fp_t *e, uint8_t out[32]) {
    for (int i = 0; i < 4; ++i) {
        uint64_t limb = e->l[3 - i];
        for (int b = 0; b < 8; ++b) out[8 * i + b] = (uint8_t)(limb >> (56 - 8 * b));
    }
}

/* Keeps bytes [0, keep): mask_least_significant_bytes (hash/mod.rs:5-13) */
void oracle_mask_lsb(uint8_t d[32], int keep) { for (int i = keep; i < 32; ++i) d[i] = 0; }
/* Keeps the last `keep` bytes: mask_most_significant_bytes (hash/mod.rs:15-23) */
void oracle_mask_msb(uint8_t d[32], int keep) { for (int i = 0; i < 32 - keep; ++i) d[i] = 0; }

void oracle_hash_bytes(int hash_kind, const uint8_t *in, size_t len, uint8_t out[32]) {
    switch (hash_kind) {
    case ORACLE_HASH_KECCAK:      oracle_keccak256(in, len, out); break;
    case ORACLE_HASH_KECCAK_M20:  oracle_keccak256(in, len, out); oracle_mask_lsb(out, 20); break;
    case ORACLE_HASH_BLAKE2S:     oracle_blake2s256(in, len, out); break;
    case ORACLE_HASH_BLAKE2S_M20: oracle_blake2s256(in, len, out); oracle_mask_msb(out, 20); break;
    case ORACLE_HASH_SHA256:      oracle_sha256(in, len, out); break;
    default: memset(out, 0xEE, 32);
    }
}
