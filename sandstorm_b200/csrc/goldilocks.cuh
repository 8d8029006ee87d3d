// Goldilocks field p = 2^64 - 2^32 + 1: single-limb arithmetic for the NTT path of BASELINE config 4
// (ministark-gpu `fields::p18446744069414584321::ark::Fp`, wired at reference cli/src/main.rs:103-124).
//
// Memory format: ark-ff `Fp64<MontBackend<_, 1>>`, i.e. one u64 = x * 2^64 mod p, canonical.  The transforms are
// LINEAR, so the kernels work directly on the stored words with plain twiddles and the plain reduction
// 2^64 = 2^32 - 1, 2^96 = -1 (mod p): NTT(stored) is the stored form of NTT(values); no Montgomery step is needed.
#pragma once
#include <cstdint>
#include "arith.cuh"

namespace ss {
namespace gl {

constexpr uint64_t P = 0xFFFFFFFF00000001ull;
constexpr uint64_t EPS = 0xFFFFFFFFull;        // 2^64 mod p
constexpr uint64_t GENERATOR = 7;              // multiplicative generator = LDE coset offset

SS_HD uint64_t add(uint64_t a, uint64_t b) {
    uint64_t s = a + b;
    if (s < a) s += EPS;                        // wrapped: + 2^64 = + EPS
    if (s >= P) s -= P;
    return s;
}
SS_HD uint64_t sub(uint64_t a, uint64_t b) {
    uint64_t d = a - b;
    if (a < b) d -= EPS;                        // borrowed: - 2^64 = - EPS
    return d;
}
// (lo + 2^64 hi) mod p
SS_HD uint64_t reduce128(uint64_t lo, uint64_t hi) {
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
    uint64_t t0 = lo - hi_hi;
    if (lo < hi_hi) t0 -= EPS;                  // 2^96 = -1
    const uint64_t t1 = hi_lo * EPS;            // 2^64 = 2^32 - 1
    uint64_t t2 = t0 + t1;
    if (t2 < t1) t2 += EPS;
    if (t2 >= P) t2 -= P;
    return t2;
}
SS_HD uint64_t mul(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return reduce128(a * b, __umul64hi(a, b));
#else
    const unsigned __int128 x = (unsigned __int128)a * b;
    return reduce128((uint64_t)x, (uint64_t)(x >> 64));
#endif
}
SS_HD uint64_t pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) {
        if (e & 1) r = mul(r, a);
        a = mul(a, a);
        e >>= 1;
    }
    return r;
}
SS_HD uint64_t inv(uint64_t a) { return pow(a, P - 2); }
SS_HD uint64_t root_of_unity(int log_n) { return pow(GENERATOR, (P - 1) >> log_n); }

}  // namespace gl
}  // namespace ss
