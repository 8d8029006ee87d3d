// Device hash cores used by the Merkle kernels: Keccak-256 (legacy padding, rate 136), Blake2s-256,
// SHA-256.  They reproduce the reference's HashFn / ElementHashFn impls byte for byte
// (crypto/src/hash/keccak.rs:13-97, blake2s.rs:10-100, masks hash/mod.rs:5-23).  All functions are
// __host__ __device__ so the CPU-only test suite can check them (tests/test_host_hashes.py).
#pragma once
#include <cstdint>
#include "arith.cuh"

namespace ss {
namespace hash {

SS_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
SS_HD uint32_t rotr32(uint32_t x, int r) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(x, x, r);
#else
    return (x >> r) | (x << (32 - r));
#endif
}
SS_HD uint32_t bswap32(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, 0, 0x0123);
#else
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
#endif
}
SS_HD uint64_t bswap64(uint64_t x) {
    return ((uint64_t)bswap32((uint32_t)x) << 32) | bswap32((uint32_t)(x >> 32));
}

// ------------------------------------------------------------------------------ Keccak-f[1600]
SS_HD void keccak_f(uint64_t (&s)[25]) {
    static constexpr uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#pragma unroll 1
    for (int round = 0; round < 24; ++round) {
        uint64_t c0 = s[0] ^ s[5] ^ s[10] ^ s[15] ^ s[20];
        uint64_t c1 = s[1] ^ s[6] ^ s[11] ^ s[16] ^ s[21];
        uint64_t c2 = s[2] ^ s[7] ^ s[12] ^ s[17] ^ s[22];
        uint64_t c3 = s[3] ^ s[8] ^ s[13] ^ s[18] ^ s[23];
        uint64_t c4 = s[4] ^ s[9] ^ s[14] ^ s[19] ^ s[24];
        const uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1),
                       d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
        // theta + rho + pi into b[]
        const uint64_t b0 = s[0] ^ d0;
        const uint64_t b1 = rotl64(s[6] ^ d1, 44), b2 = rotl64(s[12] ^ d2, 43), b3 = rotl64(s[18] ^ d3, 21), b4 = rotl64(s[24] ^ d4, 14);
        const uint64_t b5 = rotl64(s[3] ^ d3, 28), b6 = rotl64(s[9] ^ d4, 20), b7 = rotl64(s[10] ^ d0, 3), b8 = rotl64(s[16] ^ d1, 45), b9 = rotl64(s[22] ^ d2, 61);
        const uint64_t b10 = rotl64(s[1] ^ d1, 1), b11 = rotl64(s[7] ^ d2, 6), b12 = rotl64(s[13] ^ d3, 25), b13 = rotl64(s[19] ^ d4, 8), b14 = rotl64(s[20] ^ d0, 18);
        const uint64_t b15 = rotl64(s[4] ^ d4, 27), b16 = rotl64(s[5] ^ d0, 36), b17 = rotl64(s[11] ^ d1, 10), b18 = rotl64(s[17] ^ d2, 15), b19 = rotl64(s[23] ^ d3, 56);
        const uint64_t b20 = rotl64(s[2] ^ d2, 62), b21 = rotl64(s[8] ^ d3, 55), b22 = rotl64(s[14] ^ d4, 39), b23 = rotl64(s[15] ^ d0, 41), b24 = rotl64(s[21] ^ d1, 2);
        // chi
        s[0] = b0 ^ (~b1 & b2); s[1] = b1 ^ (~b2 & b3); s[2] = b2 ^ (~b3 & b4); s[3] = b3 ^ (~b4 & b0); s[4] = b4 ^ (~b0 & b1);
        s[5] = b5 ^ (~b6 & b7); s[6] = b6 ^ (~b7 & b8); s[7] = b7 ^ (~b8 & b9); s[8] = b8 ^ (~b9 & b5); s[9] = b9 ^ (~b5 & b6);
        s[10] = b10 ^ (~b11 & b12); s[11] = b11 ^ (~b12 & b13); s[12] = b12 ^ (~b13 & b14); s[13] = b13 ^ (~b14 & b10); s[14] = b14 ^ (~b10 & b11);
        s[15] = b15 ^ (~b16 & b17); s[16] = b16 ^ (~b17 & b18); s[17] = b17 ^ (~b18 & b19); s[18] = b18 ^ (~b19 & b15); s[19] = b19 ^ (~b15 & b16);
        s[20] = b20 ^ (~b21 & b22); s[21] = b21 ^ (~b22 & b23); s[22] = b22 ^ (~b23 & b24); s[23] = b23 ^ (~b24 & b20); s[24] = b24 ^ (~b20 & b21);
        s[0] ^= RC[round];
    }
}

// Keccak-256 over a message given as a stream of little-endian 64-bit lanes (message length is a
// multiple of 8 bytes on every path we hash: rows of 32-byte elements and 64-byte node pairs).
// `lane(i)` returns message lane i, n_lanes = message bytes / 8.
template <typename LaneFn>
SS_HD void keccak256_lanes(LaneFn lane, int n_lanes, uint64_t (&digest)[4]) {
    uint64_t s[25];
#pragma unroll
    for (int i = 0; i < 25; ++i) s[i] = 0;
    int done = 0;
    while (true) {
        const int remaining = n_lanes - done;
        if (remaining >= 17) {
#pragma unroll
            for (int l = 0; l < 17; ++l) s[l] ^= lane(done + l);
            keccak_f(s);
            done += 17;
        } else {
            // final block: message lanes, then 0x01 at the next byte, 0x80 at byte 135
#pragma unroll
            for (int l = 0; l < 17; ++l) {
                uint64_t v = 0;
                if (l < remaining) v = lane(done + l);
                else if (l == remaining) v = 0x01ULL;
                if (l == 16) v ^= 0x8000000000000000ULL;
                s[l] ^= v;
            }
            keccak_f(s);
            break;
        }
    }
    digest[0] = s[0]; digest[1] = s[1]; digest[2] = s[2]; digest[3] = s[3];
}

// ------------------------------------------------------------------------------------ Blake2s
SS_HD void blake2s_compress(uint32_t (&h)[8], const uint32_t (&m)[16], uint32_t t, bool last) {
    static constexpr uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
    uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
    uint32_t v8 = IV[0], v9 = IV[1], v10 = IV[2], v11 = IV[3], v12 = IV[4] ^ t, v13 = IV[5], v14 = last ? ~IV[6] : IV[6], v15 = IV[7];
#define SS_B2G(a, b, c, d, x, y)                                  \
    a = a + b + (x); d = rotr32(d ^ a, 16); c = c + d; b = rotr32(b ^ c, 12); \
    a = a + b + (y); d = rotr32(d ^ a, 8);  c = c + d; b = rotr32(b ^ c, 7);
#define SS_B2ROUND(s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15) \
    SS_B2G(v0, v4, v8, v12, m[s0], m[s1]) SS_B2G(v1, v5, v9, v13, m[s2], m[s3])          \
    SS_B2G(v2, v6, v10, v14, m[s4], m[s5]) SS_B2G(v3, v7, v11, v15, m[s6], m[s7])        \
    SS_B2G(v0, v5, v10, v15, m[s8], m[s9]) SS_B2G(v1, v6, v11, v12, m[s10], m[s11])      \
    SS_B2G(v2, v7, v8, v13, m[s12], m[s13]) SS_B2G(v3, v4, v9, v14, m[s14], m[s15])
    SS_B2ROUND(0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    SS_B2ROUND(14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)
    SS_B2ROUND(11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)
    SS_B2ROUND(7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)
    SS_B2ROUND(9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)
    SS_B2ROUND(2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)
    SS_B2ROUND(12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)
    SS_B2ROUND(13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)
    SS_B2ROUND(6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)
    SS_B2ROUND(10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)
#undef SS_B2ROUND
#undef SS_B2G
    h[0] ^= v0 ^ v8; h[1] ^= v1 ^ v9; h[2] ^= v2 ^ v10; h[3] ^= v3 ^ v11;
    h[4] ^= v4 ^ v12; h[5] ^= v5 ^ v13; h[6] ^= v6 ^ v14; h[7] ^= v7 ^ v15;
}

// Blake2s-256 over a message of n_words little-endian 32-bit words (multiple of 8 words, > 0).
template <typename WordFn>
SS_HD void blake2s256_words(WordFn word, int n_words, uint32_t (&h)[8]) {
    h[0] = 0x6A09E667u ^ 0x01010020u; h[1] = 0xBB67AE85u; h[2] = 0x3C6EF372u; h[3] = 0xA54FF53Au;
    h[4] = 0x510E527Fu; h[5] = 0x9B05688Cu; h[6] = 0x1F83D9ABu; h[7] = 0x5BE0CD19u;
    int done = 0;
    while (true) {
        const int remaining = n_words - done;
        uint32_t m[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = (i < remaining) ? word(done + i) : 0u;
        const bool last = remaining <= 16;
        const uint32_t t = (uint32_t)(last ? n_words : done + 16) * 4u;
        blake2s_compress(h, m, t, last);
        if (last) break;
        done += 16;
    }
}

// ------------------------------------------------------------------------------------- SHA-256
SS_HD void sha256_block(uint32_t (&h)[8], uint32_t (&w)[16]) {
    static constexpr uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
        0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
        0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
        0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
        0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
        0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
        0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; ++i) {
        if (i >= 16) {
            const uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            const uint32_t s0 = rotr32(w15, 7) ^ rotr32(w15, 18) ^ (w15 >> 3);
            const uint32_t s1 = rotr32(w2, 17) ^ rotr32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        const uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        const uint32_t ch = (e & f) ^ (~e & g);
        const uint32_t t1 = hh + S1 + ch + K[i] + w[i & 15];
        const uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        const uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + S0 + mj;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

// SHA-256 over n_words big-endian 32-bit words (multiple of 8 words).
template <typename WordFn>
SS_HD void sha256_words(WordFn word, int n_words, uint32_t (&h)[8]) {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
    int done = 0;
    while (done + 16 <= n_words) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = word(done + i);
        sha256_block(h, w);
        done += 16;
    }
    // n_words is a multiple of 8, so the tail holds 0 or 8 message words: padding always fits
    const int r = n_words - done;
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 14; ++i) w[i] = (i < r) ? word(done + i) : (i == r ? 0x80000000u : 0u);
    w[14] = 0;
    w[15] = (uint32_t)n_words * 32u;
    sha256_block(h, w);
}

}  // namespace hash
}  // namespace ss
