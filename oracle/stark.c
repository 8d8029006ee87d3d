/* ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * CPU restatement of the stages of ministark's prove loop that follow the commitments,
 * so that the WHOLE hot path has a CPU counterpart: it is the checker of the GPU pipeline at small sizes
 * (tests/test_prover_gpu.py: every root, OOD value, FRI root and the remainder bit for bit) and the
 * `cpu_baseline` / `--impl reference` leg of bench.py (all host threads, OpenMP over rows).
 *
 *   oracle_constraint_eval   evaluates a compiled composition / DEEP program (sandstorm_b200/air/program.py blob, the
 *                            flattened Expr DAG of AirConfig::composition_constraint, layouts/src/recursive/air.rs:1184-1200)
 *                            on LDE rows — ministark's `eval_constraint` [not vendored].  Same program, but every value is
 *                            kept canonical (modular add / sub instead of the device's lazy bounds): results are equal
 *                            as field elements, i.e. bit-identical after the final canonical store.
 *   oracle_inv_x_minus_c     1 / (3 w_N^i - c)  (batch inversion), the denominators of boundary constraints and DEEP
 *   oracle_horner            P(z) for natural-order coefficients — ministark's OOD evaluation (horner_evaluate)
 *   oracle_fri_fold          one FRI fold by 2^log_fold from the definition f(x) = sum_j x^j F_j(x^F)
 *                            (ministark FriProver::build_layers / fold_positions, options cli/src/main.rs:55-60)
 */
#include "fp252.h"
#include <stdlib.h>
#include <string.h>

enum { OP_NOP, OP_MOV, OP_ADD, OP_SUBK, OP_RED, OP_MUL, OP_DOT, OP_INV, OP_OUT };
enum { K_SLOT, K_CONST, K_TAP, K_TABLE, K_X };

typedef struct {
    const uint32_t *code, *tdesc, *taps;
    uint32_t n_words, n_consts, n_tables, n_slots, n_taps;
    const fp_t *consts;
    fp_t *tables;                 /* scaled copy */
    const fp_t *cols;
    uint64_t stride;
    int log_N;
} prog_t;

static inline fp_t fetch(const prog_t *p, uint32_t w, const fp_t *s, uint64_t i, const fp_t *x) {
    const uint32_t pay = w & 0x1fffffffu;
    switch (w >> 29) {
    case K_SLOT: return s[pay];
    case K_CONST: return p->consts[pay];
    case K_TAP: {
        const uint64_t row = (i + p->taps[2 * pay + 1]) & ((1ull << p->log_N) - 1);
        return p->cols[(uint64_t)p->taps[2 * pay] * p->stride + row];
    }
    case K_TABLE: {
        const uint32_t lp = p->tdesc[2 * pay] & 0xffu;
        return p->tables[p->tdesc[2 * pay + 1] + (i & ((1ull << lp) - 1))];
    }
    default: return *x;
    }
}

/* returns 0 on success */
int oracle_constraint_eval(const uint8_t *blob, size_t bytes, const fp_t *cols, uint64_t stride, int log_N,
                           uint64_t row_begin, uint64_t row_count, int log_step, fp_t *out) {
    const uint32_t *w = (const uint32_t *)blob;
    if (bytes < 64 || w[0] != 0x50435353u || w[1] != 3) return -1;
    prog_t p;
    p.n_words = w[2]; p.n_consts = w[3]; p.n_tables = w[4]; p.n_slots = w[5]; p.n_taps = w[8];
    const size_t nt = p.n_tables + (p.n_tables & 1), ntap = p.n_taps + (p.n_taps & 1);
    p.tdesc = w + 16; p.taps = p.tdesc + 2 * nt; p.code = p.taps + 2 * ntap;
    size_t head = (64 + 8 * nt + 8 * ntap + 16 * (size_t)p.n_words + 31) / 32 * 32;
    p.consts = (const fp_t *)(blob + head);
    const fp_t *tab_src = p.consts + p.n_consts;
    size_t tab_elems = 0;
    for (uint32_t t = 0; t < p.n_tables; ++t) tab_elems += (size_t)1 << (p.tdesc[2 * t] & 0xffu);
    if (head + 32 * (p.n_consts + tab_elems) != bytes) return -2;
    p.tables = (fp_t *)malloc(tab_elems ? 32 * tab_elems : 32);
    memcpy(p.tables, tab_src, 32 * tab_elems);
    for (uint32_t t = 0; t < p.n_tables; ++t) {          /* tables stored as (values, scale constant): multiply them out */
        const uint32_t scale = p.tdesc[2 * t] >> 8;
        if (!scale) continue;
        const size_t T = (size_t)1 << (p.tdesc[2 * t] & 0xffu), off = p.tdesc[2 * t + 1];
        for (size_t j = 0; j < T; ++j) fp_mul(&p.tables[off + j], &p.tables[off + j], &p.consts[scale - 1]);
    }
    p.cols = cols; p.stride = stride; p.log_N = log_N;
    const uint64_t N = 1ull << log_N;
    if (row_count == 0) { row_begin = 0; row_count = N >> log_step; }
    fp_t wN, g3;
    fp_root_of_unity(&wN, log_N);
    fp_generator(&g3);
    fp_t wstep;
    fp_pow_u64(&wstep, &wN, 1ull << log_step);
    int bad = 0;
#pragma omp parallel
    {
        fp_t *s = (fp_t *)malloc(sizeof(fp_t) * (p.n_slots ? p.n_slots : 1));
#pragma omp for schedule(static)
        for (int64_t chunk = 0; chunk < (int64_t)((row_count + 1023) / 1024); ++chunk) {
            const uint64_t t0 = (uint64_t)chunk * 1024, t1 = t0 + 1024 < row_count ? t0 + 1024 : row_count;
            fp_t x;
            fp_pow_u64(&x, &wN, row_begin + (t0 << log_step));
            fp_mul(&x, &x, &g3);
            for (uint64_t t = t0; t < t1; ++t) {
                const uint64_t i = row_begin + (t << log_step);
                for (uint32_t pc = 0; pc < p.n_words; ++pc) {
                    const uint32_t *ins = p.code + 4 * pc;
                    const uint32_t op = ins[0] & 0xffu, d = (ins[0] >> 8) & 0xffu, n = ins[0] >> 16;
                    fp_t a, b;
                    switch (op) {
                    case OP_MOV: s[d] = fetch(&p, ins[1], s, i, &x); break;
                    case OP_ADD: a = fetch(&p, ins[1], s, i, &x); b = fetch(&p, ins[2], s, i, &x); fp_add(&s[d], &a, &b); break;
                    case OP_SUBK: a = fetch(&p, ins[1], s, i, &x); b = fetch(&p, ins[2], s, i, &x); fp_sub(&s[d], &a, &b); break;
                    case OP_RED: break;
                    case OP_MUL: a = fetch(&p, ins[1], s, i, &x); b = fetch(&p, ins[2], s, i, &x); fp_mul(&s[d], &a, &b); break;
                    case OP_DOT: {
                        fp_t acc = FP_ZERO, prod;
                        for (uint32_t k = 0; k < n; ++k) {
                            const uint32_t *pr = p.code + 4 * (pc + 1 + k / 2) + 2 * (k & 1);
                            a = fetch(&p, pr[0], s, i, &x); b = fetch(&p, pr[1], s, i, &x);
                            fp_mul(&prod, &a, &b);
                            fp_add(&acc, &acc, &prod);
                        }
                        pc += (n + 1) / 2;
                        s[d] = acc;
                        break;
                    }
                    case OP_INV: a = fetch(&p, ins[1], s, i, &x); fp_inv(&s[d], &a); break;
                    case OP_OUT: out[i >> log_step] = fetch(&p, ins[1], s, i, &x); break;
                    case OP_NOP: break;
                    default: bad = 1;
                    }
                }
                fp_mul(&x, &x, &wstep);
            }
        }
        free(s);
    }
    free(p.tables);
    return bad ? -3 : 0;
}

/* out[i] = 1 / (3 w_N^i - c) for the rows i = k << log_step (other rows untouched); c in Montgomery form */
void oracle_inv_x_minus_c(int log_N, int log_step, const fp_t *c, fp_t *out) {
    const uint64_t N = 1ull << log_N, cnt = N >> log_step;
    fp_t wN, g3, wstep;
    fp_root_of_unity(&wN, log_N);
    fp_generator(&g3);
    fp_pow_u64(&wstep, &wN, 1ull << log_step);
    const uint64_t CH = 4096;
#pragma omp parallel for schedule(static)
    for (int64_t chunk = 0; chunk < (int64_t)((cnt + CH - 1) / CH); ++chunk) {
        const uint64_t t0 = (uint64_t)chunk * CH, t1 = t0 + CH < cnt ? t0 + CH : cnt;
        fp_t tmp[4096];
        fp_t x;
        fp_pow_u64(&x, &wN, t0 << log_step);
        fp_mul(&x, &x, &g3);
        for (uint64_t t = t0; t < t1; ++t) {
            fp_sub(&tmp[t - t0], &x, c);
            fp_mul(&x, &x, &wstep);
        }
        fp_batch_inv(tmp, t1 - t0);
        for (uint64_t t = t0; t < t1; ++t) out[t << log_step] = tmp[t - t0];
    }
}

/* r = sum_k coeffs[k] z^k, natural order (ministark horner_evaluate); parallel over blocks of coefficients */
void oracle_horner(const fp_t *coeffs, uint64_t n, const fp_t *z, fp_t *r) {
    const uint64_t CH = 1 << 14;
    const uint64_t blocks = (n + CH - 1) / CH;
    fp_t *part = (fp_t *)malloc(sizeof(fp_t) * blocks);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < (int64_t)blocks; ++b) {
        const uint64_t lo = (uint64_t)b * CH, hi = lo + CH < n ? lo + CH : n;
        fp_t acc = FP_ZERO;
        for (uint64_t k = hi; k-- > lo;) {
            fp_mul(&acc, &acc, z);
            fp_add(&acc, &acc, &coeffs[k]);
        }
        part[b] = acc;
    }
    fp_t zc, acc = FP_ZERO;
    fp_pow_u64(&zc, z, CH);
    for (uint64_t b = blocks; b-- > 0;) {
        fp_mul(&acc, &acc, &zc);
        fp_add(&acc, &acc, &part[b]);
    }
    free(part);
    *r = acc;
}

/* One FRI fold by F = 2^log_fold of evaluations on h<w_N> (natural order):  f(x) = sum_{j<F} x^j F_j(x^F),
 * out[i] = sum_j alpha^j F_j(x_i^F) (times F when no_inv_f).  The F values f(x_i w_F^k) sit at i + k N/F; F_j(y) = (1 / (F x_i^j)) sum_k f(x_i w_F^k) w_F^(-jk). */
void oracle_fri_fold(const fp_t *evals, int log_n, int log_fold, const fp_t *alpha, const fp_t *h, int no_inv_f, fp_t *out) {
    const uint64_t N = 1ull << log_n, F = 1ull << log_fold, M = N >> log_fold;
    fp_t wN, wF, wF_inv, Finv, hinv, wN_inv;
    fp_root_of_unity(&wN, log_n);
    fp_root_of_unity(&wF, log_fold);
    fp_inv(&wF_inv, &wF);
    fp_from_u64(&Finv, F);
    fp_inv(&Finv, &Finv);
    fp_inv(&hinv, h);
    fp_inv(&wN_inv, &wN);
#pragma omp parallel for schedule(static)
    for (int64_t chunk = 0; chunk < (int64_t)((M + 1023) / 1024); ++chunk) {
        const uint64_t i0 = (uint64_t)chunk * 1024, i1 = i0 + 1024 < M ? i0 + 1024 : M;
        fp_t xinv;                                   /* 1 / x_i = h^-1 w_N^-i */
        fp_pow_u64(&xinv, &wN_inv, i0);
        fp_mul(&xinv, &xinv, &hinv);
        for (uint64_t i = i0; i < i1; ++i) {
            fp_t acc = FP_ZERO, apow = FP_ONE, xj = FP_ONE;   /* apow = alpha^j, xj = x_i^-j */
            for (uint64_t j = 0; j < F; ++j) {
                fp_t fj = FP_ZERO, wk = FP_ONE, wj;
                fp_pow_u64(&wj, &wF_inv, j);
                for (uint64_t k = 0; k < F; ++k) {
                    fp_t t;
                    fp_mul(&t, &evals[i + k * M], &wk);
                    fp_add(&fj, &fj, &t);
                    fp_mul(&wk, &wk, &wj);
                }
                fp_mul(&fj, &fj, &xj);
                fp_mul(&fj, &fj, &apow);
                fp_add(&acc, &acc, &fj);
                fp_mul(&apow, &apow, alpha);
                fp_mul(&xj, &xj, &xinv);
            }
            if (no_inv_f) out[i] = acc;                      /* StarkWare / ministark: no 1/F (pinned by the reference's proofs) */
            else fp_mul(&out[i], &acc, &Finv);
            fp_mul(&xinv, &xinv, &wN_inv);
        }
    }
}

/* ark-poly coset_fft / coset_ifft on one column, in place (coset offset = Fp::GENERATOR = 3) */
void oracle_ntt_fp252(fp_t *a, int log_n, int inverse);
void oracle_distribute_powers(fp_t *a, size_t n, const fp_t *h);
void oracle_coset_ntt_fp252(fp_t *a, int log_n, int inverse) {
    fp_t g;
    fp_generator(&g);
    if (!inverse) {
        oracle_distribute_powers(a, (size_t)1 << log_n, &g);
        oracle_ntt_fp252(a, log_n, 0);
    } else {
        oracle_ntt_fp252(a, log_n, 1);
        fp_inv(&g, &g);
        oracle_distribute_powers(a, (size_t)1 << log_n, &g);
    }
}
