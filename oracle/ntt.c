/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header).
 *
 * Restates the semantics of ark-poly 0.4.2 `Radix2EvaluationDomain::{fft, ifft}`
 * and their coset forms (third-party: Cargo.lock:160-162; not vendored in
 * /root/reference), which is what ministark's `Matrix::interpolate` /
 * `Matrix::evaluate` run per column (call sites
 * layouts/src/recursive/air.rs:66-67, builtins/src/pedersen/periodic.rs:1186-1187):
 *
 *   fft(c)[i]  = sum_k c[k] * (h * w^i)^k          natural order, w = 3^((p-1)/n)
 *   ifft(e)[k] = h^-k * n^-1 * sum_i e[i] * w^(-ik)
 *
 * with h = 1 for the plain domain and h = Fp::GENERATOR = 3 for the LDE coset.
 * Pinned by the reference's periodic-column KATs (tests/test_oracle_kat.py).
 */
#include "fp252.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* in-place, natural in -> natural out.  inverse != 0 also scales by n^-1. */
void oracle_ntt_fp252(fp_t *a, int log_n, int inverse) {
    const size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    fp_t w;
    fp_root_of_unity(&w, log_n);
    if (inverse) fp_inv(&w, &w);
    /* twiddle table w^0 .. w^(n/2-1) */
    fp_t *tw = (fp_t *)malloc((n / 2) * sizeof(fp_t));
    tw[0] = FP_ONE;
    for (size_t i = 1; i < n / 2; ++i) fp_mul(&tw[i], &tw[i - 1], &w);
    for (size_t i = 0; i < n; ++i) {
        size_t j = bitrev(i, log_n);
        if (i < j) { fp_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int s = 1; s <= log_n; ++s) {
        const size_t half = (size_t)1 << (s - 1);
        const size_t step = n >> s;            /* twiddle stride */
        #pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t idx = 0; idx < n / 2; ++idx) {
            const size_t blk = idx / half, j = idx % half;
            fp_t *lo = &a[blk * 2 * half + j], *hi = lo + half;
            fp_t t;
            fp_mul(&t, hi, &tw[j * step]);
            fp_sub(hi, lo, &t);
            fp_add(lo, lo, &t);
        }
    }
    if (inverse) {
        fp_t ninv;
        fp_from_u64(&ninv, (uint64_t)n);
        fp_inv(&ninv, &ninv);
        #pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t i = 0; i < n; ++i) fp_mul(&a[i], &a[i], &ninv);
    }
    free(tw);
}

/* a[k] *= h^k  (ark-poly `distribute_powers`) */
void oracle_distribute_powers(fp_t *a, size_t n, const fp_t *h) {
    fp_t acc = FP_ONE;
    for (size_t k = 0; k < n; ++k) {
        fp_mul(&a[k], &a[k], &acc);
        fp_mul(&acc, &acc, h);
    }
}

/* Matrix::interpolate on one column (size n = 2^log_n), in place. */
void oracle_interpolate_fp252(fp_t *col, int log_n) { oracle_ntt_fp252(col, log_n, 1); }

/* Matrix::evaluate on the LDE coset: coeffs (n) -> evaluations on 3*<w_N>, N = n << log_blowup.
 * out has room for N elements. */
void oracle_evaluate_coset_fp252(const fp_t *coeffs, int log_n, int log_blowup, fp_t *out) {
    const size_t n = (size_t)1 << log_n, N = n << log_blowup;
    memcpy(out, coeffs, n * sizeof(fp_t));
    memset(out + n, 0, (N - n) * sizeof(fp_t));
    fp_t g;
    fp_generator(&g);
    oracle_distribute_powers(out, n, &g);
    oracle_ntt_fp252(out, log_n + log_blowup, 0);
}

/* Full per-column LDE: trace evaluations on <w_n> -> evaluations on 3*<w_N>. */
void oracle_lde_fp252(const fp_t *trace_col, int log_n, int log_blowup, fp_t *out) {
    const size_t n = (size_t)1 << log_n;
    fp_t *tmp = (fp_t *)malloc(n * sizeof(fp_t));
    memcpy(tmp, trace_col, n * sizeof(fp_t));
    oracle_interpolate_fp252(tmp, log_n);
    oracle_evaluate_coset_fp252(tmp, log_n, log_blowup, out);
    free(tmp);
}

/* Column-major matrix versions (columns contiguous, like ministark::Matrix). */
void oracle_ntt_fp252_batch(fp_t *cols, int n_cols, int log_n, int inverse) {
    const size_t n = (size_t)1 << log_n;
    for (int c = 0; c < n_cols; ++c) oracle_ntt_fp252(cols + (size_t)c * n, log_n, inverse);
}

void oracle_lde_fp252_batch(const fp_t *cols, int n_cols, int log_n, int log_blowup, fp_t *out) {
    const size_t n = (size_t)1 << log_n, N = n << log_blowup;
    for (int c = 0; c < n_cols; ++c)
        oracle_lde_fp252(cols + (size_t)c * n, log_n, log_blowup, out + (size_t)c * N);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
