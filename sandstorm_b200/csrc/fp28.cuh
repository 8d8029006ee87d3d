// Carry-free Fp252 arithmetic for the NTT inner loops: 9 limbs of 28 bits held in 32-bit registers.
//
// Why: on B200 a carry-chained IMAD.WIDE.U32.X issues at half the rate of a plain IMAD.WIDE.U32 and
// also occupies the ALU pipe, so the radix-2^32 Montgomery product (fp252.cuh) and the carry-propagating
// adds of a butterfly serialise (profiles/r01_ntt_tile_ab.md: 275 + 98 -> 410 cycles per warp-butterfly).
// With 28-bit limbs a 9 x 9 product is 81 plain IMAD.WIDE into 64-bit column sums that cannot overflow,
// the Montgomery reduction is 10 x 3 more (p = 1 + 2^24*2^168 + 2^196 + 2^27*2^224 in radix 2^28, and
// -p^-1 = -1 mod 2^28), and additions are nine independent IADD3 with no carries at all.
//
//   F28 value = sum l[i] << (28 i).   "normalised": l[0..7] < 2^28 (l[8] takes what is left).
//   Lazy values keep limbs < 2^32 and value < 2^256; see the bounds at each function.
//
// Memory format is untouched (x * 2^256 mod p, canonical).  Every multiplication in the NTT is by a table
// constant, stored pre-scaled as c_int = c * 2^280 mod p (= stored Montgomery constant * 2^24), so that the
// radix-2^280 reduction of  x_stored * c_int  gives  (x c) * 2^256  again: data never leaves the R = 2^256 form.
#pragma once
#include "fp252.cuh"

namespace ss {

struct F28 {
    uint32_t l[9];
};

namespace f28 {

constexpr uint32_t M28 = (1u << 28) - 1u;

// c + a * b as one IMAD.WIDE.U32 (kept opaque so that multiplications by 2^24, 2^27 and 1 are not
// strength-reduced into 64-bit shift/add sequences on the ALU pipe, which is the busy one)
SS_HD uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
#if defined(__CUDA_ARCH__)
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return c + (uint64_t)a * b;
#endif
}

// 8 x u32 canonical words -> normalised limbs
SS_HD F28 from_fp(const Fp &a) {
    const uint32_t *w = a.l;
    F28 r;
    r.l[0] = w[0] & M28;
    r.l[1] = ((w[0] >> 28) | (w[1] << 4)) & M28;
    r.l[2] = ((w[1] >> 24) | (w[2] << 8)) & M28;
    r.l[3] = ((w[2] >> 20) | (w[3] << 12)) & M28;
    r.l[4] = ((w[3] >> 16) | (w[4] << 16)) & M28;
    r.l[5] = ((w[4] >> 12) | (w[5] << 20)) & M28;
    r.l[6] = ((w[5] >> 8) | (w[6] << 24)) & M28;
    r.l[7] = (w[6] >> 4) & M28;
    r.l[8] = w[7];
    return r;
}

// normalised limbs (l[0..7] < 2^28, value < 2^256) -> 8 x u32 words
SS_HD Fp to_fp(const F28 &a) {
    const uint32_t *l = a.l;
    Fp r;
    r.l[0] = l[0] | (l[1] << 28);
    r.l[1] = (l[1] >> 4) | (l[2] << 24);
    r.l[2] = (l[2] >> 8) | (l[3] << 20);
    r.l[3] = (l[3] >> 12) | (l[4] << 16);
    r.l[4] = (l[4] >> 16) | (l[5] << 12);
    r.l[5] = (l[5] >> 20) | (l[6] << 8);
    r.l[6] = (l[6] >> 24) | (l[7] << 4);
    r.l[7] = l[8];
    return r;
}

// a (limbs < 2^32, value < 2^256) times w (normalised, value < p):  result normalised,
// value = (a*w + m*p) / 2^280 in (a*w/2^280, a*w/2^280 + p]  <  p + 2^228.
// Column sums: <= 9 products < 2^60 each plus reduction terms < 2^57: below 2^63.3, no overflow.
// The three multipliers of the reduction (2^24, 1, 2^27) are passed in from a kernel parameter so that
// ptxas cannot strength-reduce the IMAD.WIDE into LEA / shift sequences on the (busier) ALU pipe.
struct MulK { uint32_t k24, k1, k27; };
SS_HD MulK mulk_literal() { MulK k; k.k24 = 0x1000000u; k.k1 = 1u; k.k27 = 0x8000000u; return k; }

SS_HD F28 mulc(const F28 &a, const F28 &w, const MulK &K) {
    uint64_t c[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)a.l[i] * w.l[j];          // IMAD.WIDE.U32, no carries
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint64_t t = c[i] + carry;
        const uint32_t m = (0u - (uint32_t)t) & M28;                                 // m = -t mod 2^28
        carry = (t + m) >> 28;                                                        // low 28 bits cancel
        c[i + 6] = madw(m, K.k24, c[i + 6]);                                          // m * p:  p = 1 + 2^24 B^6 + B^7 + 2^27 B^8
        c[i + 7] = madw(m, K.k1, c[i + 7]);
        c[i + 8] = madw(m, K.k27, c[i + 8]);
    }
    F28 r;
#pragma unroll
    for (int k = 10; k < 18; ++k) {
        const uint64_t t = c[k] + carry;
        r.l[k - 10] = (uint32_t)t & M28;
        carry = t >> 28;
    }
    r.l[8] = (uint32_t)carry;
    return r;
}

// limb-wise sum: no carries.  Caller keeps limbs < 2^32 and the value < 2^256.
SS_HD F28 add(const F28 &a, const F28 &b) {
    F28 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + b.l[i];
    return r;
}

// Multiples of p written with every limb >= a floor, so that a - b + K never borrows:
//   K_DIF: limbs in [2^30, 2^30 + 2^28)   (b limbs < 2^30)      value ~ 2^254.00
//   K_DIT: limbs in [2^28, 2^29)          (b normalised)        value ~ 2^252.00
// K = T + ((-T) mod p) with T = floor * (B^9 - 1)/(B - 1); the limbs below are that sum, limb-wise.
#define SS_K_DIF {0x40000009u, 0x4ffffffcu, 0x4ffffffbu, 0x4ffffffbu, 0x4ffffffbu, 0x4ffffffbu, 0x48fffffbu, 0x40000005u, 0x47fffffcu}   /* 9p */
#define SS_K_DIT {0x10000003u, 0x1fffffffu, 0x1ffffffeu, 0x1ffffffeu, 0x1ffffffeu, 0x1ffffffeu, 0x12fffffeu, 0x10000002u, 0x17ffffffu}   /* 3p */
SS_HD F28 sub_bias(const F28 &a, const F28 &b, const uint32_t (&K)[9]) {
    F28 r;
#pragma unroll
    for (int i = 0; i < 9; ++i) r.l[i] = a.l[i] + K[i] - b.l[i];
    return r;
}

// Round-end reduction: limbs < 2^31.6, value < 2^256  ->  normalised, value < 2^252 + 9 * 2^224.
// q' = max((l8 >> 27) - 1, 0) copies of p are subtracted (never below zero: l8 * 2^224 >= (q'+1) 2^251 > q' p),
// with signed 64-bit carries so that individual limbs may go negative on the way.
SS_HD F28 weak_reduce(const F28 &a) {
    const uint32_t qe = a.l[8] >> 27;
    const int64_t q = qe ? (int64_t)qe - 1 : 0;
    F28 r;
    int64_t t = (int64_t)a.l[0] - q;
    r.l[0] = (uint32_t)t & M28; t >>= 28;
#pragma unroll
    for (int i = 1; i < 6; ++i) { t += (int64_t)a.l[i]; r.l[i] = (uint32_t)t & M28; t >>= 28; }
    t += (int64_t)a.l[6] - (q << 24); r.l[6] = (uint32_t)t & M28; t >>= 28;
    t += (int64_t)a.l[7] - q;         r.l[7] = (uint32_t)t & M28; t >>= 28;
    t += (int64_t)a.l[8] - (q << 27);
    r.l[8] = (uint32_t)t;
    return r;
}

SS_HD F28 sub_dif(const F28 &a, const F28 &b) { const uint32_t K[9] = SS_K_DIF; return sub_bias(a, b, K); }   // a - b + 9p
SS_HD F28 sub_dit(const F28 &a, const F28 &b) { const uint32_t K[9] = SS_K_DIT; return sub_bias(a, b, K); }   // a - b + 3p

// normalised value < 4p  ->  canonical words
SS_HD Fp to_canonical_fp(const F28 &a) { return fp::canon(to_fp(a)); }

// stored Montgomery constant (c * 2^256 mod p, canonical)  ->  NTT table form c * 2^280 mod p as limbs
SS_HD F28 const_from_mont(const Fp &c_mont) {
    // 2^24 in Montgomery form is 2^280 mod p; Montgomery-multiplying by it multiplies the integer by 2^24
    Fp two24 = fp::zero();
    two24.l[0] = 1u << 24;
    const Fp t = fp::canon(fp::mul(c_mont, fp::canon(fp::mul(two24, fp::r2()))));
    return from_fp(t);
}

}  // namespace f28
}  // namespace ss
