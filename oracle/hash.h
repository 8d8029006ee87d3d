/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header). */
#ifndef ORACLE_HASH_H
#define ORACLE_HASH_H
#include "fp252.h"

enum {
    ORACLE_HASH_KECCAK = 0,       /* Keccak256HashFn            keccak.rs:13   */
    ORACLE_HASH_KECCAK_M20 = 1,   /* MaskedKeccak256HashFn<20>  keccak.rs:61   */
    ORACLE_HASH_BLAKE2S = 2,      /* Blake2sHashFn              blake2s.rs:10  */
    ORACLE_HASH_BLAKE2S_M20 = 3,  /* MaskedBlake2sHashFn<20>    blake2s.rs:64  */
    ORACLE_HASH_SHA256 = 4        /* ministark Sha256HashFn (cli/src/main.rs:119) */
};

void oracle_keccak256(const uint8_t *in, size_t len, uint8_t out[32]);
void oracle_blake2s256(const uint8_t *in, size_t len, uint8_t out[32]);
void oracle_sha256(const uint8_t *in, size_t len, uint8_t out[32]);
void oracle_felt_to_be32(const fp_t *e, uint8_t out[32]);
void oracle_mask_lsb(uint8_t d[32], int keep);
void oracle_mask_msb(uint8_t d[32], int keep);
void oracle_hash_bytes(int hash_kind, const uint8_t *in, size_t len, uint8_t out[32]);

/* builtins/src/pedersen/mod.rs:31-36 (-> starknet-crypto 0.6.1 pedersen_hash) */
void oracle_pedersen_hash(fp_t *r, const fp_t *a, const fp_t *b);

#endif
