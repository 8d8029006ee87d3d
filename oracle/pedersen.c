/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header).
 *
 * Pedersen hash with StarkWare's parameters, as called by the reference at
 * builtins/src/pedersen/mod.rs:31-36.  The arithmetic lives in the third-party
 * crate starknet-crypto 0.6.1 (Cargo.lock:1565-1567, not vendored); this
 * restates its published definition (mod.rs:24-30 doc comment):
 *
 *   H(a,b) = [ P0 + a_low*P1 + a_high*P2 + b_low*P3 + b_high*P4 ]_x
 *
 * a_low = low 248 bits, a_high = top 4 bits of the canonical integer.
 * Curve y^2 = x^3 + x + beta (builtins/src/utils.rs:141-152); P0..P4 from
 * builtins/src/pedersen/constants.rs:6-29.  Pinned by the reference KATs
 * builtins/src/pedersen/mod.rs:184-211 (tests/test_oracle_kat.py).
 */
#include "hash.h"
#include <string.h>
#include <stdlib.h>

typedef struct { fp_t x, y; } aff_t;           /* Montgomery-form coordinates */
typedef struct { fp_t x, y, z; } jac_t;         /* z == 0 <=> infinity          */

static const uint64_t PEDERSEN_BASE[5][2][4] = {
    /* P0 */ {{0x551fde4050ca6804ULL, 0x716b0b1022947733ULL, 0x00ee1b87eb599f16ULL, 0x049ee3eba8c16007ULL},
              {0xd0405d266e10268aULL, 0x4e621062c0e056c1ULL, 0xf346d49d06ea0ed3ULL, 0x03ca0cfe4b3bc6ddULL}},
    /* P1 */ {{0x1080d17957ebe47bULL, 0x8fa8120b6d56eb0cULL, 0x969c748655fca9e5ULL, 0x0234287dcbaffe7fULL},
              {0x6ed0268ee89e5615ULL, 0x940135dd7a6c94ccULL, 0x1e889527d41f4e39ULL, 0x03b056f100f96fb2ULL}},
    /* P2 */ {{0xb7a6932dba8aa378ULL, 0x99099ec1de5e3018ULL, 0x3f9dab2656558f33ULL, 0x04fa56f376c83db3ULL},
              {0x5168f4e80ff5b54dULL, 0x562761f92a7a23b4ULL, 0x8113e0c0e47e4401ULL, 0x03fa0984c931c9e3ULL}},
    /* P3 */ {{0x3aa372f0bd2d6997ULL, 0x40c690c74709e90fULL, 0x764910f75b45f74bULL, 0x04ba4cc166be8decULL},
              {0x48151f27b24b219cULL, 0xcac5c59a5ce5ae7cULL, 0x4b971e46c4ede85fULL, 0x0040301cf5c1751fULL}},
    /* P4 */ {{0xd36ff12c49a58202ULL, 0x2ca65048d53fb325ULL, 0x6e44cca8f61a63bbULL, 0x054302dcb0e6cc1cULL},
              {0x879dcc77e99c2426ULL, 0xce98ad783c25561aULL, 0xb348046268d8ae25ULL, 0x01b77b3e37d13504ULL}},
};

static void jac_double(jac_t *r, const jac_t *p) {
    if (fp_is_zero(&p->z) || fp_is_zero(&p->y)) { memset(r, 0, sizeof *r); return; }
    fp_t xx, yy, yyyy, zz, s, m, t;
    fp_sqr(&xx, &p->x);
    fp_sqr(&yy, &p->y);
    fp_sqr(&yyyy, &yy);
    fp_sqr(&zz, &p->z);
    fp_mul(&s, &p->x, &yy);
    fp_add(&s, &s, &s); fp_add(&s, &s, &s);            /* 4*X*YY */
    fp_sqr(&t, &zz);                                    /* a*ZZ^2, a = 1 */
    fp_add(&m, &xx, &xx); fp_add(&m, &m, &xx); fp_add(&m, &m, &t);
    fp_t x3, y3, z3;
    fp_sqr(&x3, &m); fp_sub(&x3, &x3, &s); fp_sub(&x3, &x3, &s);
    fp_sub(&t, &s, &x3); fp_mul(&y3, &m, &t);
    fp_add(&yyyy, &yyyy, &yyyy); fp_add(&yyyy, &yyyy, &yyyy); fp_add(&yyyy, &yyyy, &yyyy);
    fp_sub(&y3, &y3, &yyyy);
    fp_mul(&z3, &p->y, &p->z); fp_add(&z3, &z3, &z3);
    r->x = x3; r->y = y3; r->z = z3;
}

static void jac_add_affine(jac_t *r, const jac_t *p, const aff_t *q) {
    if (fp_is_zero(&p->z)) { r->x = q->x; r->y = q->y; r->z = FP_ONE; return; }
    fp_t z1z1, u2, s2, h, rr, hh, hhh, v, t;
    fp_sqr(&z1z1, &p->z);
    fp_mul(&u2, &q->x, &z1z1);
    fp_mul(&s2, &q->y, &p->z); fp_mul(&s2, &s2, &z1z1);
    fp_sub(&h, &u2, &p->x);
    fp_sub(&rr, &s2, &p->y);
    if (fp_is_zero(&h)) {
        if (fp_is_zero(&rr)) { jac_double(r, p); return; }
        memset(r, 0, sizeof *r); return;
    }
    fp_sqr(&hh, &h);
    fp_mul(&hhh, &hh, &h);
    fp_mul(&v, &p->x, &hh);
    fp_t x3, y3, z3;
    fp_sqr(&x3, &rr); fp_sub(&x3, &x3, &hhh); fp_sub(&x3, &x3, &v); fp_sub(&x3, &x3, &v);
    fp_sub(&t, &v, &x3); fp_mul(&y3, &rr, &t);
    fp_mul(&t, &p->y, &hhh); fp_sub(&y3, &y3, &t);
    fp_mul(&z3, &p->z, &h);
    r->x = x3; r->y = y3; r->z = z3;
}

static void jac_to_affine(aff_t *r, const jac_t *p) {
    fp_t zi, zi2, zi3;
    fp_inv(&zi, &p->z);
    fp_sqr(&zi2, &zi);
    fp_mul(&zi3, &zi2, &zi);
    fp_mul(&r->x, &p->x, &zi2);
    fp_mul(&r->y, &p->y, &zi3);
}

/* doubling tables: TAB[k][i] = 2^i * P_{k+1}; 252 entries are enough for any shift */
static aff_t PED_P0;
static aff_t PED_TAB[4][248];
static int PED_READY = 0;

static void pedersen_init(void) {
    #pragma omp critical(pedersen_init_lock)
    {
        if (!PED_READY) {
            fp_t c;
            memcpy(c.l, PEDERSEN_BASE[0][0], 32); fp_to_mont(&PED_P0.x, &c);
            memcpy(c.l, PEDERSEN_BASE[0][1], 32); fp_to_mont(&PED_P0.y, &c);
            for (int k = 0; k < 4; ++k) {
                aff_t cur;
                memcpy(c.l, PEDERSEN_BASE[k + 1][0], 32); fp_to_mont(&cur.x, &c);
                memcpy(c.l, PEDERSEN_BASE[k + 1][1], 32); fp_to_mont(&cur.y, &c);
                const int count = (k % 2 == 0) ? 248 : 4;
                for (int i = 0; i < count; ++i) {
                    PED_TAB[k][i] = cur;
                    jac_t j = {cur.x, cur.y, FP_ONE}, d;
                    jac_double(&d, &j);
                    jac_to_affine(&cur, &d);
                }
            }
            PED_READY = 1;
        }
    }
}

void oracle_pedersen_hash(fp_t *r, const fp_t *a, const fp_t *b) {
    if (!PED_READY) pedersen_init();
    jac_t acc = {PED_P0.x, PED_P0.y, FP_ONE};
    const fp_t *in[2] = {a, b};
    for (int e = 0; e < 2; ++e) {
        fp_t c;
        fp_from_mont(&c, in[e]);               /* canonical integer */
        for (int i = 0; i < 252; ++i) {
            if (!((c.l[i / 64] >> (i % 64)) & 1)) continue;
            const aff_t *pt = (i < 248) ? &PED_TAB[2 * e][i] : &PED_TAB[2 * e + 1][i - 248];
            jac_t t;
            jac_add_affine(&t, &acc, pt);
            acc = t;
        }
    }
    aff_t res;
    jac_to_affine(&res, &acc);
    *r = res.x;
}

/* PedersenHashFn::hash_elements (crypto/src/hash/pedersen.rs:67-76):
 * h = 0; for e: h = H(h, e); return H(h, count). */
void oracle_pedersen_hash_elements(fp_t *r, const fp_t *elems, size_t n) {
    fp_t h = FP_ZERO, cnt;
    for (size_t i = 0; i < n; ++i) oracle_pedersen_hash(&h, &h, &elems[i]);
    fp_from_u64(&cnt, (uint64_t)n);
    oracle_pedersen_hash(r, &h, &cnt);
}
