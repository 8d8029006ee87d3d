// Fp252 = Z / (2^251 + 17*2^192 + 1), Montgomery form with R = 2^256, 8 x u32 limbs.
//
// Memory format (identical bytes to the reference's ark-ff `Fp256<MontBackend<_,4>>`, i.e. the
// 4 x u64 little-endian limbs that reference crypto/src/utils.rs:15-17 exposes): canonical
// Montgomery residues in [0, p).  Inside kernels values live in a LAZY domain [0, M),
// M = 2^252 + 2^224 (a hair above 2p), which removes every multi-limb compare from the hot loops:
//
//   mul(a,b)   inputs < 2^255          -> (a*b/R, a*b/R + p]      (no final subtraction)
//   add(a,b)   inputs < 2^253          -> < M                      (two top-limb tests)
//   sub(a,b)   inputs < 2^253          -> < M
//   sub4p(a,b) inputs < 2^253          -> a - b + 4p  < 2^254.1    (feeds a mul directly)
//   canon(a)   input  < 4p + small     -> [0, p)                   (only at global stores)
//
// The reduction exploits p = 1 + k*2^192 with k = 2^59 + 17:  p^-1 = 1 - k*2^192 (mod 2^256), so
// q = T_lo * p^-1 mod 2^256 = T_lo - ((k*T_lo mod 2^64) << 192), and
// (T - q*p) / 2^256 = T_hi - delta - ((k*q) >> 64)  exactly (delta = borrow of the q fix-up).
// The "+p" that keeps the result positive is pre-loaded into the product accumulator for free.
#pragma once
#include "arith.cuh"

namespace ss {

struct alignas(16) Fp {
    uint32_t l[8];
};

namespace fp {

// p, 2p, 4p as 32-bit limbs
#define SS_P0 0x00000001u
#define SS_P6 0x00000011u
#define SS_P7 0x08000000u

SS_HD Fp zero() { Fp r; for (int i = 0; i < 8; ++i) r.l[i] = 0; return r; }
// R mod p  (Montgomery form of 1)
SS_HD Fp one() {
    Fp r;
    r.l[0] = 0xffffffe1u; r.l[1] = 0xffffffffu; r.l[2] = 0xffffffffu; r.l[3] = 0xffffffffu;
    r.l[4] = 0xffffffffu; r.l[5] = 0xffffffffu; r.l[6] = 0xfffffdf0u; r.l[7] = 0x07ffffffu;
    return r;
}
// R^2 mod p
SS_HD Fp r2() {
    Fp r;
    r.l[0] = 0x7e000401u; r.l[1] = 0xfffffd73u; r.l[2] = 0x330fffffu; r.l[3] = 0x00000001u;
    r.l[4] = 0xff6f8000u; r.l[5] = 0xffffffffu; r.l[6] = 0x5e008810u; r.l[7] = 0x07ffd4abu;
    return r;
}

// T[0..15] = a*b + p*2^256.  Operand scanning with two interleaved accumulators so that every
// 32x32 product is one (lo,hi) pair on aligned limbs => ptxas emits one IMAD.WIDE.U32 per product.
template <bool PLUS_P = true>
SS_HD void mul_wide_plus_p(uint32_t (&T)[16], const Fp &a, const Fp &b) {
    using namespace ptx;
    uint32_t E[16], O[16];   // O[k] sits at limb position k+1
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        const uint64_t pe = (uint64_t)a.l[j] * b.l[0];          // IMAD.WIDE.U32
        const uint64_t po = (uint64_t)a.l[j + 1] * b.l[0];
        E[j] = (uint32_t)pe;
        E[j + 1] = (uint32_t)(pe >> 32);
        O[j] = (uint32_t)po;
        O[j + 1] = (uint32_t)(po >> 32);
    }
    E[8] = PLUS_P ? SS_P0 : 0u; E[9] = 0; E[10] = 0; E[11] = 0; E[12] = 0; E[13] = 0; E[14] = PLUS_P ? SS_P6 : 0u; E[15] = PLUS_P ? SS_P7 : 0u;
#pragma unroll
    for (int j = 8; j < 16; ++j) O[j] = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        const uint32_t bi = b.l[i];
        if ((i & 1) == 0) {
            // even row: a[even j] -> E at i+j ; a[odd j] -> O at index i+j-1
            mad4_chain(E[i], E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8],
                       a.l[0], a.l[2], a.l[4], a.l[6], bi);
            mad4_chain(O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7], O[i + 8],
                       a.l[1], a.l[3], a.l[5], a.l[7], bi);
        } else {
            // odd row: a[odd j] -> E at i+j ; a[even j] -> O at index i+j-1
            if (i + 9 < 16)
                mad4_chain(E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8], E[(i + 9) & 15],
                           a.l[1], a.l[3], a.l[5], a.l[7], bi);
            else
                mad4_chain_top(E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8],
                               a.l[1], a.l[3], a.l[5], a.l[7], bi);
            mad4_chain(O[i - 1], O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7],
                       a.l[0], a.l[2], a.l[4], a.l[6], bi);
        }
    }
    T[0] = E[0];
    T[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 15; ++k) T[k] = addc_cc(E[k], O[k - 1]);
    T[15] = addc(E[15], O[14]);
}

// (T - q*p) / 2^256 for T = a*b + p*2^256.  See file header.
SS_HD Fp mont_reduce(const uint32_t (&T)[16]) {
    using namespace ptx;
    // kL0 = (2^59 + 17) * (T[0] + 2^32 T[1]) mod 2^64
    const uint64_t t17 = (uint64_t)T[0] * 17u;
    const uint32_t k0 = (uint32_t)t17;
    const uint32_t k1 = (uint32_t)(t17 >> 32) + T[1] * 17u + (T[0] << 27);
    // q = T_lo - (kL0 << 192)  mod 2^256 ; delta = borrow out
    uint32_t q[8];
#pragma unroll
    for (int i = 0; i < 6; ++i) q[i] = T[i];
    q[6] = sub_cc(T[6], k0);
    q[7] = subc_cc(T[7], k1);
    const uint32_t delta = subc(0, 0) & 1u;      // 0 - 0 - borrow = 0xffffffff when borrow
    // U = q * (17 + 2^27 * 2^32); only U >> 64 is needed (low 64 bits equal kL0 by construction)
    uint32_t u[8];
    uint64_t acc = (uint64_t)q[0] * 17u;
    acc = (acc >> 32) + (uint64_t)q[1] * 17u + (uint64_t)q[0] * 0x08000000u;
    acc >>= 32;
#pragma unroll
    for (int i = 2; i < 8; ++i) {
        acc += (uint64_t)q[i] * 17u + (uint64_t)q[i - 1] * 0x08000000u;
        u[i - 2] = (uint32_t)acc;
        acc >>= 32;
    }
    acc += (uint64_t)q[7] * 0x08000000u;
    u[6] = (uint32_t)acc;
    u[7] = (uint32_t)(acc >> 32);
    // r = T_hi - delta - u
    Fp r;
    (void)sub_cc(0, delta);                       // sets borrow = delta
    r.l[0] = subc_cc(T[8], u[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.l[i] = subc_cc(T[8 + i], u[i]);
    r.l[7] = subc(T[15], u[7]);
    return r;
}

SS_HD Fp mul(const Fp &a, const Fp &b) {
    uint32_t T[16];
    mul_wide_plus_p(T, a, b);
    return mont_reduce(T);
}

SS_HD Fp sqr(const Fp &a) { return mul(a, a); }

// x -= 2p  /  x -= 4p  when the top limb says it is safe (see header)
SS_HD void cond_sub_2p(Fp &x) {
    using namespace ptx;
    if (x.l[7] > 0x10000000u) {
        x.l[0] = sub_cc(x.l[0], 2u);
#pragma unroll
        for (int i = 1; i < 6; ++i) x.l[i] = subc_cc(x.l[i], 0u);
        x.l[6] = subc_cc(x.l[6], 0x22u);
        x.l[7] = subc(x.l[7], 0x10000000u);
    }
}
SS_HD void cond_sub_4p(Fp &x) {
    using namespace ptx;
    if (x.l[7] > 0x20000000u) {
        x.l[0] = sub_cc(x.l[0], 4u);
#pragma unroll
        for (int i = 1; i < 6; ++i) x.l[i] = subc_cc(x.l[i], 0u);
        x.l[6] = subc_cc(x.l[6], 0x44u);
        x.l[7] = subc(x.l[7], 0x20000000u);
    }
}

// raw a + b (caller tracks the bound)
SS_HD Fp add_raw(const Fp &a, const Fp &b) {
    using namespace ptx;
    Fp r;
    r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.l[i] = addc_cc(a.l[i], b.l[i]);
    r.l[7] = addc(a.l[7], b.l[7]);
    return r;
}

// a, b < 2^253  ->  < M   (closed invariant; two top-limb tests)
SS_HD Fp add(const Fp &a, const Fp &b) {
    Fp r = add_raw(a, b);
    cond_sub_4p(r);
    cond_sub_2p(r);
    return r;
}
// NTT inner loop: one test only.  a, b < 2^252 + c*2^224  ->  < 2^252 + 2c*2^224, i.e. the slack
// doubles per nested add; at most 12 adds separate two canonicalising global stores, so values stay
// < 2^252 + 2^237 and every bound used by mul / sub4p below holds with a wide margin.
SS_HD Fp add_fast(const Fp &a, const Fp &b) {
    Fp r = add_raw(a, b);
    cond_sub_2p(r);
    return r;
}

// a - b + c*p for c in {2,4}: c2 = 2c (limb0), c6 = 0x11c, c7 = c<<27
template <int C>
SS_HD Fp sub_plus_cp(const Fp &a, const Fp &b) {
    using namespace ptx;
    Fp r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.l[i] = subc_cc(a.l[i], b.l[i]);
    r.l[7] = subc(a.l[7], b.l[7]);
    r.l[0] = add_cc(r.l[0], (uint32_t)C);
#pragma unroll
    for (int i = 1; i < 6; ++i) r.l[i] = addc_cc(r.l[i], 0u);
    r.l[6] = addc_cc(r.l[6], 0x11u * (uint32_t)C);
    r.l[7] = addc(r.l[7], 0x08000000u * (uint32_t)C);
    return r;
}
// a, b < 2^253 -> a - b + 4p in (0, 2^254.1); valid multiplicand
SS_HD Fp sub4p(const Fp &a, const Fp &b) { return sub_plus_cp<4>(a, b); }
// a < anything, b < 2p -> a - b + 2p (positive), grows by 2p
SS_HD Fp sub2p(const Fp &a, const Fp &b) { return sub_plus_cp<2>(a, b); }
// a, b < 2^253 -> < M
SS_HD Fp sub(const Fp &a, const Fp &b) {
    Fp r = sub_plus_cp<4>(a, b);
    cond_sub_4p(r);
    cond_sub_2p(r);
    return r;
}
SS_HD Fp neg(const Fp &a) { return sub(zero(), a); }

// exact: x < 4p + small  ->  [0, p)
SS_HD Fp canon(const Fp &x) {
    using namespace ptx;
    Fp r = x, t;
    // trial subtract 2p
    t.l[0] = sub_cc(r.l[0], 2u);
#pragma unroll
    for (int i = 1; i < 6; ++i) t.l[i] = subc_cc(r.l[i], 0u);
    t.l[6] = subc_cc(r.l[6], 0x22u);
    t.l[7] = subc_cc(r.l[7], 0x10000000u);
    uint32_t borrow = subc(0, 0);
    if (!borrow) r = t;
    t.l[0] = sub_cc(r.l[0], 2u);
#pragma unroll
    for (int i = 1; i < 6; ++i) t.l[i] = subc_cc(r.l[i], 0u);
    t.l[6] = subc_cc(r.l[6], 0x22u);
    t.l[7] = subc_cc(r.l[7], 0x10000000u);
    borrow = subc(0, 0);
    if (!borrow) r = t;
    // trial subtract p
    t.l[0] = sub_cc(r.l[0], 1u);
#pragma unroll
    for (int i = 1; i < 6; ++i) t.l[i] = subc_cc(r.l[i], 0u);
    t.l[6] = subc_cc(r.l[6], 0x11u);
    t.l[7] = subc_cc(r.l[7], 0x08000000u);
    borrow = subc(0, 0);
    if (!borrow) r = t;
    return r;
}

// ---- primitives of the constraint-program interpreter (constraint_eval.cu) ------------------------
// The compiler (sandstorm_b200/air/program.py) tracks an exact upper bound of every value, so the
// device never tests magnitudes: it adds raw, subtracts with a pre-computed multiple of p and
// reduces only where the compiler asks.

// any x < 2^256  ->  congruent value < 2^252:  x -= max((x >> 251) - 1, 0) * p
SS_HD Fp red(const Fp &x) {
    using namespace ptx;
    const uint32_t q = x.l[7] >> 27;
    const uint32_t m = q - (q != 0u ? 1u : 0u);
    Fp r;
    r.l[0] = sub_cc(x.l[0], m);
#pragma unroll
    for (int i = 1; i < 6; ++i) r.l[i] = subc_cc(x.l[i], 0u);
    r.l[6] = subc_cc(x.l[6], 17u * m);
    r.l[7] = subc(x.l[7], m << 27);
    return r;
}
// a - b + k*p for a run-time k in [0, 31]; exact when k*p >= b and a + k*p < 2^256
SS_HD Fp sub_kp(const Fp &a, const Fp &b, uint32_t k) {
    using namespace ptx;
    Fp r;
    r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r.l[i] = subc_cc(a.l[i], b.l[i]);
    r.l[7] = subc(a.l[7], b.l[7]);
    r.l[0] = add_cc(r.l[0], k);
#pragma unroll
    for (int i = 1; i < 6; ++i) r.l[i] = addc_cc(r.l[i], 0u);
    r.l[6] = addc_cc(r.l[6], 17u * k);
    r.l[7] = addc(r.l[7], k << 27);
    return r;
}

// Unreduced sum of products  sum_k a_k * b_k + p * 2^256  (< 2^512, guaranteed by the compiler), kept as
// the two interleaved accumulators of mul_wide_plus_p plus carry counters: the carry out of every
// 4-product chain is counted at its limb position instead of being rippled to the top.
struct WideAcc {
    uint32_t E[16], O[16];   // O[k] sits at limb position k + 1
    uint32_t CE[4], CO[4];   // carries into E positions 8,10,12,14 and O indices 8,10,12,14
};
SS_HD void acc_init(WideAcc &w) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { w.E[i] = 0; w.O[i] = 0; }
    w.E[8] = SS_P0; w.E[14] = SS_P6; w.E[15] = SS_P7;
#pragma unroll
    for (int i = 0; i < 4; ++i) { w.CE[i] = 0; w.CO[i] = 0; }
}
SS_HD void acc_mac(WideAcc &w, const Fp &a, const Fp &b) {
    using namespace ptx;
    uint32_t (&E)[16] = w.E;
    uint32_t (&O)[16] = w.O;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t bi = b.l[i];
        if ((i & 1) == 0) {
            mad4_chain(E[i], E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], w.CE[i / 2],
                       a.l[0], a.l[2], a.l[4], a.l[6], bi);
            mad4_chain(O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7], w.CO[i / 2],
                       a.l[1], a.l[3], a.l[5], a.l[7], bi);
        } else {
            if (i < 7)
                mad4_chain(E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8], w.CE[(i + 1) / 2],
                           a.l[1], a.l[3], a.l[5], a.l[7], bi);
            else
                mad4_chain_top(E[8], E[9], E[10], E[11], E[12], E[13], E[14], E[15], a.l[1], a.l[3], a.l[5], a.l[7], bi);
            mad4_chain(O[i - 1], O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], w.CO[(i - 1) / 2],
                       a.l[0], a.l[2], a.l[4], a.l[6], bi);
        }
    }
}
// (acc - q*p) / 2^256 : in (S/R, S/R + p] for S = sum of the products
SS_HD Fp acc_reduce(const WideAcc &w) {
    using namespace ptx;
    uint32_t T[16];
    T[0] = w.E[0];
    T[1] = add_cc(w.E[1], w.O[0]);
#pragma unroll
    for (int k = 2; k < 15; ++k) T[k] = addc_cc(w.E[k], w.O[k - 1]);
    T[15] = addc(w.E[15], w.O[14]);
    // carry counters: CE[j] at limb 8 + 2j, CO[j] at limb 9 + 2j
    T[8] = add_cc(T[8], w.CE[0]);
    T[9] = addc_cc(T[9], w.CO[0]);
    T[10] = addc_cc(T[10], w.CE[1]);
    T[11] = addc_cc(T[11], w.CO[1]);
    T[12] = addc_cc(T[12], w.CE[2]);
    T[13] = addc_cc(T[13], w.CO[2]);
    T[14] = addc_cc(T[14], w.CE[3]);
    T[15] = addc(T[15], w.CO[3]);
    return mont_reduce(T);
}

SS_HD bool is_zero_canon(const Fp &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.l[i];
    return o == 0;
}
SS_HD bool eq_canon(const Fp &a, const Fp &b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) o |= a.l[i] ^ b.l[i];
    return o == 0;
}

SS_HD Fp from_u32(uint32_t v) {
    Fp t = zero();
    t.l[0] = v;
    return canon(mul(t, r2()));
}

// a^e, e as 8 x u32 little-endian limbs; result lazy (< M)
SS_HD Fp pow_limbs(const Fp &a, const uint32_t *e, int n_limbs) {
    Fp acc = one(), base = a;
    for (int i = 0; i < n_limbs * 32; ++i) {
        if ((e[i >> 5] >> (i & 31)) & 1u) acc = mul(acc, base);
        base = sqr(base);
    }
    return acc;
}
SS_HD Fp pow_u64(const Fp &a, uint64_t e) {
    Fp acc = one(), base = a;
    while (e) {
        if (e & 1) acc = mul(acc, base);
        base = sqr(base);
        e >>= 1;
    }
    return acc;
}
// a^(p-2)
SS_HD Fp inv(const Fp &a) {
    const uint32_t e[8] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00000010u, 0x08000000u};
    return pow_limbs(a, e, 8);
}

}  // namespace fp
}  // namespace ss
