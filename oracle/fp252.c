/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header). */
#include "fp252.h"
#include <string.h>
#include <stdlib.h>

typedef unsigned __int128 u128;

const fp_t FP_P    = {{0x1ULL, 0x0ULL, 0x0ULL, 0x0800000000000011ULL}};
const fp_t FP_ONE  = {{0xffffffffffffffe1ULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x07fffffffffffdf0ULL}};
const fp_t FP_R2   = {{0xfffffd737e000401ULL, 0x00000001330fffffULL, 0xffffffffff6f8000ULL, 0x07ffd4ab5e008810ULL}};
const fp_t FP_ZERO = {{0, 0, 0, 0}};

static inline int geq_p(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > FP_P.l[i]) return 1;
        if (a[i] < FP_P.l[i]) return 0;
    }
    return 1;
}

static inline void sub_p(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a[i] - FP_P.l[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
}

void fp_add(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[4];
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    /* a,b < p < 2^252 so no carry out of 256 bits */
    if (geq_p(t)) sub_p(t);
    memcpy(r->l, t, sizeof t);
}

void fp_sub(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[4];
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 d = (u128)a->l[i] - b->l[i] - borrow;
        t[i] = (uint64_t)d;
        borrow = (d >> 64) & 1;
    }
    if (borrow) {
        u128 c = 0;
        for (int i = 0; i < 4; ++i) {
            c += (u128)t[i] + FP_P.l[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, sizeof t);
}

void fp_neg(fp_t *r, const fp_t *a) {
    fp_sub(r, &FP_ZERO, a);
}

/* CIOS Montgomery multiplication, word size 2^64, -p^-1 mod 2^64 = 2^64-1
 * because p = 1 mod 2^64. */
void fp_mul(fp_t *r, const fp_t *a, const fp_t *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)t[j] + (u128)a->l[j] * b->l[i];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = (uint64_t)(0 - t[0]);
        c = (u128)t[0] + (u128)m * FP_P.l[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)t[j] + (u128)m * FP_P.l[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || geq_p(t)) sub_p(t);
    memcpy(r->l, t, 4 * sizeof(uint64_t));
}

void fp_sqr(fp_t *r, const fp_t *a) { fp_mul(r, a, a); }

void fp_pow(fp_t *r, const fp_t *a, const uint64_t e[4]) {
    fp_t acc = FP_ONE, base = *a;
    for (int i = 0; i < 256; ++i) {
        if ((e[i / 64] >> (i % 64)) & 1) fp_mul(&acc, &acc, &base);
        fp_sqr(&base, &base);
    }
    *r = acc;
}

void fp_pow_u64(fp_t *r, const fp_t *a, uint64_t e) {
    uint64_t ee[4] = {e, 0, 0, 0};
    fp_pow(r, a, ee);
}

void fp_inv(fp_t *r, const fp_t *a) {
    /* p - 2 */
    const uint64_t e[4] = {0xffffffffffffffffULL, 0xffffffffffffffffULL, 0xffffffffffffffffULL, 0x0800000000000010ULL};
    fp_pow(r, a, e);
}

void fp_to_mont(fp_t *r, const fp_t *canonical) { fp_mul(r, canonical, &FP_R2); }

void fp_from_mont(fp_t *r, const fp_t *mont) {
    fp_t one = {{1, 0, 0, 0}};
    fp_mul(r, mont, &one);
}

void fp_from_u64(fp_t *r, uint64_t v) {
    fp_t t = {{v, 0, 0, 0}};
    fp_to_mont(r, &t);
}

int fp_eq(const fp_t *a, const fp_t *b) { return memcmp(a->l, b->l, 32) == 0; }
int fp_is_zero(const fp_t *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }

/* Montgomery's trick; zeros are skipped and stay zero (ark-ff batch_inversion
 * semantics, used at reference layouts/src/recursive/trace.rs:718). */
void fp_batch_inv(fp_t *v, size_t n) {
    if (n == 0) return;
    fp_t *prefix = (fp_t *)malloc(n * sizeof(fp_t));
    fp_t acc = FP_ONE;
    for (size_t i = 0; i < n; ++i) {
        prefix[i] = acc;
        if (!fp_is_zero(&v[i])) fp_mul(&acc, &acc, &v[i]);
    }
    fp_inv(&acc, &acc);
    for (size_t i = n; i-- > 0;) {
        if (fp_is_zero(&v[i])) continue;
        fp_t tmp;
        fp_mul(&tmp, &acc, &prefix[i]);
        fp_mul(&acc, &acc, &v[i]);
        v[i] = tmp;
    }
    free(prefix);
}

void fp_generator(fp_t *r) { fp_from_u64(r, 3); }

void fp_root_of_unity(fp_t *r, int log_n) {
    /* (p-1) = 2^192 * (2^59 + 17); (p-1)/2^log_n as 256-bit exponent */
    uint64_t e[4] = {0, 0, 0, 0x0800000000000011ULL};   /* p - 1 */
    for (int s = 0; s < log_n; ++s) {                   /* shift right by one */
        for (int i = 0; i < 4; ++i) {
            e[i] >>= 1;
            if (i < 3) e[i] |= e[i + 1] << 63;
        }
    }
    fp_t g;
    fp_generator(&g);
    fp_pow(r, &g, e);
}
