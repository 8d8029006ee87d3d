/* ORACLE — TEST INFRASTRUCTURE ONLY (see fp252.h header).
 *
 * Goldilocks field p = 2^64 - 2^32 + 1 (ministark-gpu `fields::p18446744069414584321::ark::Fp`, used by the reference
 * at cli/src/main.rs:103-124; third-party, not vendored) and the ark-poly `Radix2EvaluationDomain` transforms over it:
 *
 *   fft(c)[i]  = sum_k c[k] * (h * w^i)^k          natural order, w = 7^((p-1)/n)
 *   ifft(e)[k] = h^-k * n^-1 * sum_i e[i] * w^(-ik)
 *
 * with h = 1 for the plain domain and h = GENERATOR = 7 for the LDE coset.  Elements are ark-ff
 * `Fp64<MontBackend<_, 1>>` values: one u64 holding x * 2^64 mod p.  The transforms are linear, so they are
 * computed directly on the stored words with plain (non-Montgomery) twiddles: the result is again in stored form.
 * PARITY UNPINNED for this field: the reference holds no Goldilocks known-answer test (README.md:60-64 marks the
 * field as work in progress); tests/test_goldilocks.py pins this file against a quadratic-time big-int DFT. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GL_P 0xFFFFFFFF00000001ULL
typedef unsigned __int128 u128;

static inline uint64_t gl_add(uint64_t a, uint64_t b) { u128 s = (u128)a + b; if (s >= GL_P) s -= GL_P; return (uint64_t)s; }
static inline uint64_t gl_sub(uint64_t a, uint64_t b) { return a >= b ? a - b : a + (GL_P - b); }
static inline uint64_t gl_mul(uint64_t a, uint64_t b) { return (uint64_t)(((u128)a * b) % GL_P); }
static uint64_t gl_pow(uint64_t a, uint64_t e) {
    uint64_t r = 1;
    while (e) { if (e & 1) r = gl_mul(r, a); a = gl_mul(a, a); e >>= 1; }
    return r;
}
static inline uint64_t gl_inv(uint64_t a) { return gl_pow(a, GL_P - 2); }
uint64_t oracle_gl_root_of_unity(int log_n) { return gl_pow(7, (GL_P - 1) >> log_n); }

static inline size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* in-place, natural in -> natural out; inverse != 0 also scales by n^-1 */
void oracle_gl_ntt(uint64_t *a, int log_n, int inverse) {
    const size_t n = (size_t)1 << log_n;
    if (log_n == 0) return;
    uint64_t w = oracle_gl_root_of_unity(log_n);
    if (inverse) w = gl_inv(w);
    uint64_t *tw = (uint64_t *)malloc((n / 2) * sizeof(uint64_t));
    tw[0] = 1;
    for (size_t i = 1; i < n / 2; ++i) tw[i] = gl_mul(tw[i - 1], w);
    for (size_t i = 0; i < n; ++i) {
        size_t j = bitrev(i, log_n);
        if (i < j) { uint64_t t = a[i]; a[i] = a[j]; a[j] = t; }
    }
    for (int s = 1; s <= log_n; ++s) {
        const size_t half = (size_t)1 << (s - 1), step = n >> s;
        #pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t idx = 0; idx < n / 2; ++idx) {
            const size_t blk = idx / half, j = idx % half;
            uint64_t *lo = &a[blk * 2 * half + j], *hi = lo + half;
            const uint64_t t = gl_mul(*hi, tw[j * step]);
            *hi = gl_sub(*lo, t);
            *lo = gl_add(*lo, t);
        }
    }
    free(tw);
    if (inverse) {
        const uint64_t ninv = gl_inv((uint64_t)n % GL_P);
        for (size_t i = 0; i < n; ++i) a[i] = gl_mul(a[i], ninv);
    }
}

/* a[k] *= h^k  (coset pre-scale of fft) or h^-k (post-scale of ifft) */
void oracle_gl_scale_powers(uint64_t *a, size_t n, uint64_t h) {
    uint64_t c = 1;
    for (size_t k = 0; k < n; ++k) { a[k] = gl_mul(a[k], c); c = gl_mul(c, h); }
}

/* columns of n = 2^log_n words (stride n) */
void oracle_gl_ntt_cols(uint64_t *cols, size_t n_cols, int log_n, int inverse, int coset) {
    const size_t n = (size_t)1 << log_n;
    for (size_t c = 0; c < n_cols; ++c) {
        uint64_t *a = cols + c * n;
        if (!inverse && coset) oracle_gl_scale_powers(a, n, 7);
        oracle_gl_ntt(a, log_n, inverse);
        if (inverse && coset) oracle_gl_scale_powers(a, n, gl_inv(7));
    }
}

/* Matrix::interpolate + Matrix::evaluate on the coset 7 * <w_N>: out has n << log_blowup words per column */
void oracle_gl_lde_cols(const uint64_t *trace, size_t n_cols, int log_n, int log_blowup, uint64_t *out) {
    const size_t n = (size_t)1 << log_n, N = n << log_blowup;
    for (size_t c = 0; c < n_cols; ++c) {
        uint64_t *o = out + c * N;
        memcpy(o, trace + c * n, n * sizeof(uint64_t));
        memset(o + n, 0, (N - n) * sizeof(uint64_t));
        oracle_gl_ntt(o, log_n, 1);
        oracle_gl_scale_powers(o, n, 7);
        oracle_gl_ntt(o, log_n + log_blowup, 0);
    }
}
