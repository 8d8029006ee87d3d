/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the Fp252 arithmetic the reference's hot path runs on.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library; the product (libsandstorm_b200.so)
 * never links or calls it.
 *
 * Field: p = 2^251 + 17*2^192 + 1  (reference: cli/src/main.rs:25-26).
 * In-memory element = ark-ff Fp256<MontBackend<_,4>>: 4 x u64 little-endian
 * limbs of x*2^256 mod p (reference: crypto/src/utils.rs:8-18 exposes (v.0).0
 * as "to_montgomery"; solidity.rs:173-192 KAT pins R = 2^256).
 * The arithmetic lives in ark-ff 0.4.2 (Cargo.lock:94-96), not in the
 * reference tree; exact integer arithmetic => any correct algorithm is
 * bit-identical.
 */
#ifndef ORACLE_FP252_H
#define ORACLE_FP252_H
#include <stdint.h>
#include <stddef.h>

typedef struct { uint64_t l[4]; } fp_t;

extern const fp_t FP_P;        /* modulus, plain integer limbs            */
extern const fp_t FP_ONE;      /* R mod p      (Montgomery form of 1)     */
extern const fp_t FP_R2;       /* R^2 mod p                               */
extern const fp_t FP_ZERO;

void fp_add(fp_t *r, const fp_t *a, const fp_t *b);
void fp_sub(fp_t *r, const fp_t *a, const fp_t *b);
void fp_neg(fp_t *r, const fp_t *a);
void fp_mul(fp_t *r, const fp_t *a, const fp_t *b);      /* Montgomery product */
void fp_sqr(fp_t *r, const fp_t *a);
void fp_pow_u64(fp_t *r, const fp_t *a, uint64_t e);
void fp_pow(fp_t *r, const fp_t *a, const uint64_t e[4]);
void fp_inv(fp_t *r, const fp_t *a);                      /* a^(p-2); inv(0)=0 */
void fp_from_u64(fp_t *r, uint64_t v);                    /* -> Montgomery     */
void fp_to_mont(fp_t *r, const fp_t *canonical);
void fp_from_mont(fp_t *r, const fp_t *mont);             /* -> canonical      */
int  fp_eq(const fp_t *a, const fp_t *b);
int  fp_is_zero(const fp_t *a);
void fp_batch_inv(fp_t *v, size_t n);                     /* in place, zeros stay zero */

/* omega_n = 3^((p-1)/n), n = 2^log_n  (reference KAT:
 * builtins/src/pedersen/periodic.rs:1184-1209 pins this root and natural order) */
void fp_root_of_unity(fp_t *r, int log_n);
void fp_generator(fp_t *r);                               /* Fp::GENERATOR = 3 */

#endif
