// Proof-of-work grinding: PublicCoin::grind_proof_of_work of the reference's two coins,
//   Solidity coin  crypto/src/public_coin/solidity.rs:120-141   (Keccak-256)
//   Cairo coin     crypto/src/public_coin/cairo.rs:133-154      (Blake2s-256)
// prefix = H(0x0123456789ABCDED_be || digest || bits);  the nonce is valid when H(prefix || nonce_be) starts with
// `bits` zero bits (ministark::random::leading_zeros, most significant bit of byte 0 first — not vendored, this is
// StarkWare's convention).  The reference returns the smallest nonce >= 1 in its sequential build and ANY valid nonce
// with `-F parallel` (find_any), which makes its proofs non-reproducible; this search is parallel AND returns the
// smallest one: windows of 2^24 nonces are scanned in increasing order, atomicMin inside the first window that hits.
#include "ctx.h"
#include "hashes.cuh"

using namespace ss;

namespace {

struct PowArgs {
    uint64_t prefix[4];      // H(...) as little-endian lanes (Keccak) / packed words (Blake2s)
    int bits;
};

__device__ __forceinline__ int leading_zero_bits(uint64_t first8_le, uint64_t next8_le) {
    const uint64_t a = hash::bswap64(first8_le);
    if (a) return __clzll((long long)a);
    const uint64_t b = hash::bswap64(next8_le);
    return 64 + (b ? __clzll((long long)b) : 64);
}

template <int KIND>
__global__ void __launch_bounds__(256) pow_kernel(PowArgs A, unsigned long long first, unsigned long long *best) {
    const unsigned long long nonce = first + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    int lz;
    if (KIND == 0) {
        auto lane = [&](int g) -> uint64_t { return g < 4 ? A.prefix[g] : hash::bswap64(nonce); };
        uint64_t d[4];
        hash::keccak256_lanes(lane, 5, d);
        lz = leading_zero_bits(d[0], d[1]);
    } else {
        auto word = [&](int g) -> uint32_t {
            if (g < 8) return (uint32_t)(A.prefix[g >> 1] >> (32 * (g & 1)));
            return hash::bswap32(g == 8 ? (uint32_t)(nonce >> 32) : (uint32_t)nonce);
        };
        uint32_t h[8];
        hash::blake2s256_words(word, 10, h);
        lz = leading_zero_bits((uint64_t)h[0] | ((uint64_t)h[1] << 32), (uint64_t)h[2] | ((uint64_t)h[3] << 32));
    }
    if (lz >= A.bits) atomicMin(best, nonce);
}

}  // namespace

extern "C" ss_status ss_pow_grind(ss_ctx *ctx, int hash_kind, const uint8_t digest[32], int bits, uint64_t *nonce_out) {
    if (!ctx || !digest || !nonce_out || bits < 0 || bits > 64 || (hash_kind != 0 && hash_kind != 1)) return SS_ERR_INVALID;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    // prefix hash on the host: 8 + 32 + 1 = 41 bytes, one block of either hash
    uint8_t msg[64] = {0x01, 0x23, 0x45, 0x67, 0x89, 0xAB, 0xCD, 0xED};
    for (int i = 0; i < 32; ++i) msg[8 + i] = digest[i];
    msg[40] = (uint8_t)bits;
    PowArgs A{};
    A.bits = bits;
    if (hash_kind == 0) {
        uint64_t s[25] = {0};
        msg[41] = 0x01;                                   // legacy Keccak padding (sha3 crate Keccak256)
        for (int l = 0; l < 6; ++l)
            for (int b = 0; b < 8; ++b) s[l] |= (uint64_t)msg[8 * l + b] << (8 * b);
        s[16] ^= 0x8000000000000000ULL;
        hash::keccak_f(s);
        for (int l = 0; l < 4; ++l) A.prefix[l] = s[l];
    } else {
        uint32_t h[8] = {0x6A09E667u ^ 0x01010020u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
        uint32_t m[16];
        for (int w = 0; w < 16; ++w) m[w] = (uint32_t)msg[4 * w] | ((uint32_t)msg[4 * w + 1] << 8) | ((uint32_t)msg[4 * w + 2] << 16) | ((uint32_t)msg[4 * w + 3] << 24);
        hash::blake2s_compress(h, m, 41u, true);
        for (int l = 0; l < 4; ++l) A.prefix[l] = (uint64_t)h[2 * l] | ((uint64_t)h[2 * l + 1] << 32);
    }
    unsigned long long *d_best = nullptr;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&d_best), 8));
    const unsigned long long none = ~0ull, window = 1ull << 24;
    cudaError_t e = cudaMemcpy(d_best, &none, 8, cudaMemcpyHostToDevice);
    unsigned long long best = none;
    for (unsigned long long first = 1; e == cudaSuccess && best == none && first < (1ull << 62); first += window) {
        if (hash_kind == 0) pow_kernel<0><<<(unsigned)(window / 256), 256>>>(A, first, d_best);
        else pow_kernel<1><<<(unsigned)(window / 256), 256>>>(A, first, d_best);
        ctx->launches++;
        e = cudaMemcpy(&best, d_best, 8, cudaMemcpyDeviceToHost);
    }
    dev_free(ctx, d_best);
    if (e != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_pow_grind: %s", cudaGetErrorString(e));
    if (best == none) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_pow_grind: no nonce below 2^62");
    *nonce_out = best;
    return SS_OK;
}
