// FRI layer folding and out-of-domain polynomial evaluation.
//
// ss_fri_fold: one fold by F = 2^log_fold of evaluations on the coset h*<w_N> (natural order):
//     f(x) = sum_{j<F} x^j F_j(x^F)      ->      out[i] = sum_j alpha^j F_j(y_i),  y_i = x_i^F
// The F evaluations f(x_i * w_F^k), k < F, sit at indices i + k*N/F (the rows of the reference's FRI
// layer matrix, which is therefore the input buffer itself viewed column-major with stride N/F: the
// layer commit is ss_merkle_build(d_evals, col_stride = N/F, n_cols = F) with no data movement).
// Per output: size-F inverse DFT in registers, then Horner in alpha / x_i.  flags bit 0 multiplies by F,
// which is StarkWare's convention of three binary folds with alpha, alpha^2, alpha^4 and no 1/2 factor
// (ministark's FriProver is not vendored: convention unpinned).
//
// ss_poly_eval: P(z) for polynomials stored as ss_lde writes them (coefficient k scaled by 3^k, at
// position brev(k)):  P(z) = sum_pos a[pos] * w^brev(pos), w = z/3, evaluated by log n levels of
// adjacent-pair folds  b[m] = a[2m] + w^(n/2) a[2m+1],  c[m] = b[2m] + w^(n/4) b[2m+1], ...
#include "ctx.h"
#include "fp252.cuh"
#include <algorithm>

using namespace ss;

namespace {

__device__ __forceinline__ Fp ld_fp(const Fp *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}
__device__ __forceinline__ void st_fp(Fp *p, const Fp &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

struct FoldParams {
    Fp winv[8];          // w_F^-j, j < F/2  (inverse DFT twiddles)
    Fp alpha_hinv;       // alpha / h
    Fp scale;            // 1/F, or 1 when the StarkWare (x F) convention is requested
    const Fp *xinv_lo;   // w_N^-i,       i < 4096      (the inverse NTT twiddle tables)
    const Fp *xinv_hi;   // w_N^-(4096 i)
};

template <int LOG_F>
__global__ void __launch_bounds__(256) fri_fold_kernel(const Fp *__restrict__ evals, Fp *__restrict__ out, int log_n, FoldParams P,
                                                         unsigned long long out_begin, unsigned long long out_count) {
    constexpr int F = 1 << LOG_F;
    const unsigned long long m = 1ull << (log_n - LOG_F);
    const unsigned long long tid = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (tid >= out_count) return;
    const unsigned long long i = out_begin + tid;
    Fp v[F];
#pragma unroll
    for (int k = 0; k < F; ++k) v[k] = ld_fp(evals + i + (unsigned long long)k * m);
    // inverse DFT of size F, decimation in frequency: natural in -> bit-reversed out
#pragma unroll
    for (int s = LOG_F - 1; s >= 0; --s) {
        const int span = 1 << s;
#pragma unroll
        for (int k = 0; k < F; ++k) {
            if (k & span) continue;
            const int j = k & (span - 1);                 // twiddle exponent j * F / (2 span)
            const Fp d = fp::sub4p(v[k], v[k + span]);
            v[k] = fp::add(v[k], v[k + span]);
            if (j == 0) { v[k + span] = d; fp::cond_sub_4p(v[k + span]); fp::cond_sub_2p(v[k + span]); }
            else v[k + span] = fp::mul(d, P.winv[j << (LOG_F - 1 - s)]);
        }
    }
    // t = alpha / x_i = (alpha / h) * w_N^-i ; Horner over c_j = IDFT_j / x^j (v holds IDFT in bit-reversed order)
    Fp xinv = ld_fp(P.xinv_lo + (i & 4095ull));
    if (i >> 12) xinv = fp::mul(xinv, ld_fp(P.xinv_hi + (i >> 12)));
    const Fp t = fp::mul(xinv, P.alpha_hinv);
    auto brev = [](int x) { int r = 0; for (int b = 0; b < LOG_F; ++b) r |= ((x >> b) & 1) << (LOG_F - 1 - b); return r; };
    Fp acc = v[brev(F - 1)];
#pragma unroll
    for (int j = F - 2; j >= 0; --j) acc = fp::add(fp::mul(acc, t), v[brev(j)]);
    acc = fp::mul(acc, P.scale);
    st_fp(out + i, fp::canon(acc));
}

// ---- polynomial evaluation by pairwise folding --------------------------------------------------
// One block folds 2048 contiguous values of one (column, point) job down to 1.  mult[l] is the
// multiplier of fold level l of this launch (level 0 folds adjacent pairs).  Jobs along blockIdx.y.
struct EvalJob {
    const Fp *src;            // values of this job at this stage
    Fp *dst;                  // one value per block
    const Fp *mult;           // 11 multipliers for this stage (device memory)
};

__global__ void __launch_bounds__(256) poly_fold_kernel(const EvalJob *jobs, unsigned long long n_in, int levels) {
    __shared__ Fp sm[256];
    const EvalJob job = jobs[blockIdx.y];
    const unsigned long long base = (unsigned long long)blockIdx.x * 2048ull + threadIdx.x * 8ull;
    // levels: how many of the 11 levels are real (n_in may be < 2048 at the last stage)
    Fp v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (base + k < n_in) ? ld_fp(job.src + base + k) : fp::zero();
    int lvl = 0;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        if (lvl < levels) {
            const Fp w = ld_fp(job.mult + lvl);
#pragma unroll
            for (int k = 0; k < (8 >> (s + 1)); ++k) v[k] = fp::add(v[2 * k], fp::mul(v[2 * k + 1], w));
            ++lvl;
        }
    }
    sm[threadIdx.x] = v[0];
    __syncthreads();
    for (int active = 128; active >= 1; active >>= 1) {
        if (lvl < levels) {
            Fp r;
            if ((int)threadIdx.x < active) r = fp::add(sm[2 * threadIdx.x], fp::mul(sm[2 * threadIdx.x + 1], ld_fp(job.mult + lvl)));
            __syncthreads();
            if ((int)threadIdx.x < active) sm[threadIdx.x] = r;
            __syncthreads();
            ++lvl;
        }
    }
    if (threadIdx.x == 0) st_fp(job.dst + blockIdx.x, fp::canon(sm[0]));
}

// ---- out[i] = 1 / (h * w_N^i - c) for every i: Montgomery batch inversion, R rows per thread ----------
// Every denominator X - c * g^e of the AIR boundary terms and of the DEEP quotients is a shifted read of
// this one vector:  x_i - c g^e = g^e (x_{i - b e} - c)  (g = w_N^b generates the trace domain).
constexpr int INV_ROWS = 64;           // most elements per thread, walked sequentially (prefix products parked in the output buffer)
constexpr int INV_THREADS = 128;
// elements per thread for a launch over `count` elements: long sweeps amortise the block's one inversion, but a short vector
// (one rank's piece) must still fill the GPU with blocks
static int inv_rows_for(unsigned long long count) {
    const unsigned long long want = count / ((unsigned long long)INV_THREADS * 148 * 8) + 1;
    return want >= INV_ROWS ? INV_ROWS : (want < 8 ? 8 : (int)want);
}

// Inverse of every thread's value with ONE field inversion per block (Montgomery's trick as a product tree in
// shared memory): up-sweep of pairwise products, thread 0 inverts the root (a^(p-2): 250 squarings + 11
// multiplications — the other warps wait at the barrier and the SM runs other blocks), down-sweep
// inv(left) = inv(node) * right, inv(right) = inv(node) * left.  sm: 2 * INV_THREADS elements.
__device__ __forceinline__ Fp block_inverse(const Fp &v, Fp *sm) {
    const int tid = threadIdx.x;
    sm[INV_THREADS + tid] = v;
    __syncthreads();
    for (int w = INV_THREADS / 2; w >= 1; w >>= 1) {
        if (tid < w) sm[w + tid] = fp::mul(sm[2 * (w + tid)], sm[2 * (w + tid) + 1]);
        __syncthreads();
    }
    if (tid == 0) {
        const Fp acc = sm[1];
        auto sqn = [](Fp x, int m) { for (int t = 0; t < m; ++t) x = fp::sqr(x); return x; };
        const Fp e2 = fp::mul(sqn(acc, 1), acc), e4 = fp::mul(sqn(e2, 2), e2), e8 = fp::mul(sqn(e4, 4), e4);
        const Fp e16 = fp::mul(sqn(e8, 8), e8), e32 = fp::mul(sqn(e16, 16), e16), e64 = fp::mul(sqn(e32, 32), e32);
        const Fp e128 = fp::mul(sqn(e64, 64), e64), e192 = fp::mul(sqn(e128, 64), e64);
        const Fp gq = fp::mul(e192, acc);
        sm[1] = fp::mul(sqn(fp::mul(sqn(gq, 55), gq), 4), e192);
    }
    __syncthreads();
    for (int w = 1; w < INV_THREADS; w <<= 1) {
        if (tid < w) {
            const int k = w + tid;
            const Fp ik = sm[k], l = sm[2 * k], r = sm[2 * k + 1];
            sm[2 * k] = fp::mul(ik, r);
            sm[2 * k + 1] = fp::mul(ik, l);
        }
        __syncthreads();
    }
    return sm[INV_THREADS + tid];
}
// Two sweeps per thread over its INV_ROWS elements with ONE inversion per block of INV_THREADS * INV_ROWS = 8192 elements:
// forward, the running product before each element is parked in the element's own output slot; after the block inversion the
// backward sweep recomputes the element (a table product), reads its prefix back and writes the inverse.  No per-thread arrays:
// few registers, many resident blocks, so the serial a^(p-2) chain of one block hides under the sweeps of the others (the
// previous form held 16 elements + 16 prefixes in 255 registers, two blocks per SM, and spent 80 % of its time in that chain).
template <typename ElemFn>
__device__ __forceinline__ void batch_invert_block(Fp *out, unsigned long long count, int rows, ElemFn elem, Fp *sm) {
    const unsigned long long chunk = (unsigned long long)blockDim.x * rows;
    const unsigned long long base = blockIdx.x * chunk + threadIdx.x;
    Fp acc = fp::one();
#pragma unroll 1
    for (int k = 0; k < rows; ++k) {
        const unsigned long long t = base + (unsigned long long)k * blockDim.x;
        if (t < count) {
            unsigned long long slot;
            const Fp x = elem(t, slot);
            st_fp(out + slot, acc);
            acc = fp::mul(acc, x);
        }
    }
    Fp inv = block_inverse(acc, sm);
#pragma unroll 1
    for (int k = rows - 1; k >= 0; --k) {
        const unsigned long long t = base + (unsigned long long)k * blockDim.x;
        if (t < count) {
            unsigned long long slot;
            const Fp x = elem(t, slot);
            const Fp r = fp::mul(inv, ld_fp(out + slot));
            inv = fp::mul(inv, x);
            st_fp(out + slot, fp::canon(r));
        }
    }
}

__global__ void __launch_bounds__(INV_THREADS) inv_x_minus_c_kernel(Fp *out, int log_n, int log_step, unsigned long long first, unsigned long long count, int rows, Fp c,
                                                                      const Fp *xlo, const Fp *xhi) {
    __shared__ Fp sm[2 * INV_THREADS];
    const unsigned long long mask = (1ull << log_n) - 1;
    // element t = row ((first + t) << log_step) mod N: x_i - c, x_i = h w_N^i from two tables
    batch_invert_block(out, count, rows, [&](unsigned long long t, unsigned long long &slot) {
        const unsigned long long i = ((first + t) << log_step) & mask;
        slot = i;
        Fp x = ld_fp(xlo + (i & 4095ull));
        if (i >> 12) x = fp::mul(x, ld_fp(xhi + (i >> 12)));
        return fp::sub(x, c);
    }, sm);
}

// ---- out-of-domain evaluation from the trace itself (barycentric form) ------------------------------------
// With g the generator of the trace domain (order n) and t_j = T(g^j) the trace column,
//     T(z g^k) = (z^n - 1)/n * sum_j t_{(j+k) mod n} * W_j,      W_j = g^j / (z - g^j) = 1 / (z g^-j - 1):
// every tap (column, row offset k) of the AIR is a dot product of the SAME weight vector with a rotated
// column, so the step needs no coefficient vectors, every multiplication is a multiply-accumulate into an
// unreduced 512-bit sum (fp::WideAcc, one Montgomery reduction per thread), and — unlike Horner — the sum
// splits over row ranges (one per GPU).
//
// bary_weights_kernel: W_j for j in [row_begin, row_begin + count), Montgomery batch inversion per thread.
__global__ void __launch_bounds__(INV_THREADS) bary_weights_kernel(Fp *out, unsigned long long row_begin, unsigned long long count, int rows, Fp z,
                                                                     const Fp *ginv_lo, const Fp *ginv_hi) {
    __shared__ Fp sm[2 * INV_THREADS];
    batch_invert_block(out, count, rows, [&](unsigned long long t, unsigned long long &slot) {
        const unsigned long long j = row_begin + t;
        slot = t;
        Fp x = ld_fp(ginv_lo + (j & 4095ull));
        if (j >> 12) x = fp::mul(x, ld_fp(ginv_hi + (j >> 12)));
        return fp::sub(fp::mul(x, z), fp::one());
    }, sm);
}

// ood_dot_kernel: block (x = tap pair, y = row chunk) accumulates  sum_j t[(j + off) mod n] * W_j  over its chunk
// for two taps of one column (the weight is loaded once for both) and writes one partial per tap.  Taps vary
// fastest over the grid, so the CTAs in flight share a few chunks: their weights and column windows stay in L2.
constexpr int OOD_THREADS = 128;
constexpr int OOD_ROWS = 32;                            // rows per thread
constexpr int OOD_CHUNK = OOD_THREADS * OOD_ROWS;       // rows per block
struct OodTap { int col; unsigned int off0, off1; int n_taps; };   // one column, one or two row offsets

__global__ void __launch_bounds__(OOD_THREADS) ood_dot_kernel(const Fp *__restrict__ cols, unsigned long long stride, int log_n,
                                                               const Fp *__restrict__ weights, unsigned long long row_begin,
                                                               unsigned long long count, const OodTap *__restrict__ taps, unsigned int n_pairs,
                                                               unsigned int n_chunks, Fp *__restrict__ partials) {
    __shared__ Fp sm[2][OOD_THREADS];
    const unsigned int pair = blockIdx.x % n_pairs, chunk = blockIdx.x / n_pairs;      // tap pairs vary fastest
    const OodTap tp = taps[pair];
    const unsigned long long mask = (1ull << log_n) - 1;
    const Fp *col = cols + (unsigned long long)tp.col * stride;
    const unsigned long long base = (unsigned long long)chunk * OOD_CHUNK + threadIdx.x;
    fp::WideAcc a0, a1;
    fp::acc_init(a0);
    fp::acc_init(a1);
    if (tp.n_taps == 2) {
#pragma unroll 1
        for (int k = 0; k < OOD_ROWS; ++k) {
            const unsigned long long t = base + (unsigned long long)k * OOD_THREADS;
            if (t < count) {
                const unsigned long long j = row_begin + t;
                const Fp w = ld_fp(weights + t);
                fp::acc_mac(a0, ld_fp(col + ((j + tp.off0) & mask)), w);
                fp::acc_mac(a1, ld_fp(col + ((j + tp.off1) & mask)), w);
            }
        }
    } else {
#pragma unroll 1
        for (int k = 0; k < OOD_ROWS; ++k) {
            const unsigned long long t = base + (unsigned long long)k * OOD_THREADS;
            if (t < count) fp::acc_mac(a0, ld_fp(col + ((row_begin + t + tp.off0) & mask)), ld_fp(weights + t));
        }
    }
    // Montgomery-reduce the thread sums (linear: R^-1 * sum), then a shared-memory tree of modular additions
    Fp r0 = fp::acc_reduce(a0), r1 = fp::acc_reduce(a1);
    fp::cond_sub_4p(r0); fp::cond_sub_2p(r0);
    fp::cond_sub_4p(r1); fp::cond_sub_2p(r1);
    sm[0][threadIdx.x] = r0;
    sm[1][threadIdx.x] = r1;
    __syncthreads();
    for (int active = OOD_THREADS / 2; active >= 1; active >>= 1) {
        if ((int)threadIdx.x < active) {
            sm[0][threadIdx.x] = fp::add(sm[0][threadIdx.x], sm[0][threadIdx.x + active]);
            sm[1][threadIdx.x] = fp::add(sm[1][threadIdx.x], sm[1][threadIdx.x + active]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        Fp *dst = partials + ((unsigned long long)pair * n_chunks + chunk) * 2;
        st_fp(dst, sm[0][0]);
        st_fp(dst + 1, sm[1][0]);
    }
}

// ood_finish_kernel: block e sums the partials of evaluation e over the chunks and scales by (z^n - 1)/n
__global__ void __launch_bounds__(256) ood_finish_kernel(const Fp *__restrict__ partials, unsigned int n_chunks, const int2 *__restrict__ where,
                                                          Fp scale, Fp *__restrict__ out) {
    __shared__ Fp sm[256];
    const int2 w = where[blockIdx.x];                      // (tap pair, slot 0/1)
    const Fp *src = partials + (unsigned long long)w.x * n_chunks * 2 + w.y;
    Fp acc = fp::zero();
    for (unsigned int c = threadIdx.x; c < n_chunks; c += 256) acc = fp::add(acc, ld_fp(src + 2ull * c));
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int active = 128; active >= 1; active >>= 1) {
        if ((int)threadIdx.x < active) sm[threadIdx.x] = fp::add(sm[threadIdx.x], sm[threadIdx.x + active]);
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fp(out + blockIdx.x, fp::canon(fp::mul(sm[0], scale)));
}

void fill_x_lo(Fp *dst, size_t n, int log_n, int h) {
    Fp w = fp::one();
    {   // w_N
        uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0x00000011u, 0x08000000u};
        for (int s = 0; s < log_n; ++s)
            for (int i = 0; i < 8; ++i) { e[i] >>= 1; if (i < 7) e[i] |= e[i + 1] << 31; }
        w = fp::canon(fp::pow_limbs(fp::from_u32(3), e, 8));
    }
    Fp c = fp::from_u32((uint32_t)h);
    for (size_t i = 0; i < n; ++i) { dst[i] = fp::canon(c); c = fp::mul(c, w); }
}
void fill_x_hi(Fp *dst, size_t n, int log_n, int) {
    uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0x00000011u, 0x08000000u};
    for (int s = 0; s < log_n; ++s)
        for (int i = 0; i < 8; ++i) { e[i] >>= 1; if (i < 7) e[i] |= e[i + 1] << 31; }
    const Fp w = fp::pow_u64(fp::canon(fp::pow_limbs(fp::from_u32(3), e, 8)), 4096);
    Fp c = fp::one();
    for (size_t i = 0; i < n; ++i) { dst[i] = fp::canon(c); c = fp::mul(c, w); }
}

Fp host_root_of_unity(int log_n, bool inverse) {
    uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0x00000011u, 0x08000000u};
    for (int s = 0; s < log_n; ++s)
        for (int i = 0; i < 8; ++i) { e[i] >>= 1; if (i < 7) e[i] |= e[i + 1] << 31; }
    Fp w = fp::pow_limbs(fp::from_u32(3), e, 8);
    if (inverse) w = fp::inv(w);
    return fp::canon(w);
}
void fill_inv_lo(Fp *dst, size_t n, int log_n, int) {
    const Fp w = host_root_of_unity(log_n, true);
    Fp c = fp::one();
    for (size_t i = 0; i < n; ++i) { dst[i] = fp::canon(c); c = fp::mul(c, w); }
}
void fill_inv_hi(Fp *dst, size_t n, int log_n, int) {
    const Fp w = fp::pow_u64(host_root_of_unity(log_n, true), 4096);
    Fp c = fp::one();
    for (size_t i = 0; i < n; ++i) { dst[i] = fp::canon(c); c = fp::mul(c, w); }
}

Fp load_host(const void *p) { Fp v; memcpy(v.l, p, 32); return v; }

}  // namespace

extern "C" {

ss_status ss_fri_fold(ss_ctx *ctx, ss_field field, const void *d_evals, int log_n, int log_fold, const void *h_alpha,
                      const void *h_domain_offset, int flags, uint64_t out_begin, uint64_t out_count, void *d_out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_fri_fold: field %d not built", (int)field);
    if (!d_evals || !d_out || !h_alpha || !h_domain_offset || log_fold < 1 || log_fold > 4 || log_n < log_fold || log_n > 40)
        return fail(ctx, SS_ERR_INVALID, "ss_fri_fold: bad arguments (log_n=%d log_fold=%d)", log_n, log_fold);
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const int F = 1 << log_fold;
    FoldParams P;
    const Fp wf_inv = host_root_of_unity(log_fold, true);
    Fp c = fp::one();
    for (int j = 0; j < 8; ++j) { P.winv[j] = fp::canon(c); if (j < F / 2) c = fp::mul(c, wf_inv); }
    P.alpha_hinv = fp::canon(fp::mul(load_host(h_alpha), fp::inv(load_host(h_domain_offset))));
    P.scale = (flags & 1) ? fp::one() : fp::canon(fp::inv(fp::from_u32((uint32_t)F)));
    const size_t n = (size_t)1 << log_n;
    Fp *lo, *hi;
    ss_status rc;
    // tables of w_N^-i : same contents as the inverse-NTT twiddle tables (keys shared with ntt_host.cu)
    if ((rc = cached_table(ctx, {2 /*T_LO*/, log_n, 1}, n < 4096 ? n : 4096, fill_inv_lo, &lo))) return rc;
    if ((rc = cached_table(ctx, {3 /*T_HI*/, log_n, 1}, n <= 4096 ? 1 : n / 4096, fill_inv_hi, &hi))) return rc;
    P.xinv_lo = lo; P.xinv_hi = hi;
    const unsigned long long m = 1ull << (log_n - log_fold);
    if (out_count == 0) { out_begin = 0; out_count = m; }                 // 0 = every output
    if (out_begin + out_count > m) return fail(ctx, SS_ERR_INVALID, "ss_fri_fold: output range outside the folded domain");
    const unsigned grid = (unsigned)((out_count + 255) / 256);
    cudaStream_t st = pick_stream(ctx, stream);
    const Fp *in = static_cast<const Fp *>(d_evals);
    Fp *out = static_cast<Fp *>(d_out);
    switch (log_fold) {
    case 1: fri_fold_kernel<1><<<grid, 256, 0, st>>>(in, out, log_n, P, out_begin, out_count); break;
    case 2: fri_fold_kernel<2><<<grid, 256, 0, st>>>(in, out, log_n, P, out_begin, out_count); break;
    case 3: fri_fold_kernel<3><<<grid, 256, 0, st>>>(in, out, log_n, P, out_begin, out_count); break;
    default: fri_fold_kernel<4><<<grid, 256, 0, st>>>(in, out, log_n, P, out_begin, out_count); break;
    }
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

ss_status ss_inv_x_minus_c(ss_ctx *ctx, ss_field field, int log_n, int log_row_step, uint64_t row_begin, uint64_t row_count,
                           const void *h_c, void *d_out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_inv_x_minus_c: field %d not built", (int)field);
    if (!h_c || !d_out || log_n < 0 || log_n > 40 || log_row_step < 0 || log_row_step > log_n) return fail(ctx, SS_ERR_INVALID, "ss_inv_x_minus_c: bad arguments");
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << log_n;
    Fp *lo, *hi;
    ss_status rc;
    // x_i = 3 * w_N^i: same tables (and cache keys) as the constraint evaluator's OP_X
    if ((rc = cached_table(ctx, {20, log_n, 0}, n < 4096 ? n : 4096, [](Fp *d, size_t m, int ln, int) { fill_x_lo(d, m, ln, 3); }, &lo))) return rc;
    if ((rc = cached_table(ctx, {21, log_n, 0}, n <= 4096 ? 1 : n / 4096, fill_x_hi, &hi))) return rc;
    // rows (row_begin + t) << step, t < row_count, wrapping mod N (a rank's row range plus the halo its shifted reads need)
    if (row_count == 0) { row_begin = 0; row_count = n >> log_row_step; }
    if (row_count > (n >> log_row_step)) return fail(ctx, SS_ERR_INVALID, "ss_inv_x_minus_c: row range larger than the domain");
    const int rows = inv_rows_for(row_count);
    const unsigned long long chunk = (unsigned long long)INV_THREADS * rows;
    inv_x_minus_c_kernel<<<(unsigned)((row_count + chunk - 1) / chunk), INV_THREADS, 0, pick_stream(ctx, stream)>>>(
        static_cast<Fp *>(d_out), log_n, log_row_step, row_begin, row_count, rows, fp::canon(load_host(h_c)), lo, hi);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

ss_status ss_ood_eval(ss_ctx *ctx, ss_field field, const void *d_trace_cols, uint64_t col_stride, int log_n,
                      const int32_t *h_cols, const uint64_t *h_offsets, size_t n_evals, const void *h_z,
                      uint64_t row_begin, uint64_t row_count, void *h_out) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_ood_eval: field %d not built", (int)field);
    if (!d_trace_cols || log_n < 0 || log_n > 32 || !h_z || (n_evals && (!h_cols || !h_offsets || !h_out)) || col_stride < (1ull << log_n))
        return fail(ctx, SS_ERR_INVALID, "ss_ood_eval: bad arguments");
    if (n_evals == 0) return SS_OK;
    const unsigned long long n = 1ull << log_n;
    if (row_count == 0) { row_begin = 0; row_count = n; }
    if (row_begin + row_count > n) return fail(ctx, SS_ERR_INVALID, "ss_ood_eval: row range outside the trace domain");
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    // pair up the taps of each column (the weight vector is read once per pair)
    std::vector<size_t> order(n_evals);
    for (size_t e = 0; e < n_evals; ++e) {
        if (h_cols[e] < 0 || h_offsets[e] >= n) return fail(ctx, SS_ERR_INVALID, "ss_ood_eval: bad tap %zu", e);
        order[e] = e;
    }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return h_cols[a] < h_cols[b]; });
    std::vector<OodTap> taps;
    std::vector<int2> where(n_evals);
    for (size_t k = 0; k < n_evals;) {
        const size_t e0 = order[k];
        OodTap t{h_cols[e0], (unsigned int)h_offsets[e0], 0u, 1};
        where[e0] = make_int2((int)taps.size(), 0);
        if (k + 1 < n_evals && h_cols[order[k + 1]] == h_cols[e0]) {
            const size_t e1 = order[k + 1];
            t.off1 = (unsigned int)h_offsets[e1];
            t.n_taps = 2;
            where[e1] = make_int2((int)taps.size(), 1);
            k += 2;
        } else {
            k += 1;
        }
        taps.push_back(t);
    }
    // (z^n - 1) / n on the host (log n squarings, one inversion of a small integer)
    const Fp z = fp::canon(load_host(h_z));
    Fp zn = z;
    for (int s2 = 0; s2 < log_n; ++s2) zn = fp::sqr(zn);
    Fp n_field = fp::one();
    for (int s2 = 0; s2 < log_n; ++s2) n_field = fp::add(n_field, n_field);
    const Fp scale = fp::canon(fp::mul(fp::sub(zn, fp::one()), fp::inv(n_field)));
    // powers of g^-1 (g = generator of the trace domain): lo[j] = g^-j, j < 4096; hi[j] = g^-(4096 j)
    Fp *lo, *hi;
    ss_status rc;
    if ((rc = cached_table(ctx, {24, log_n, 0}, n < 4096 ? n : 4096,
                           [](Fp *d, size_t m, int ln, int) { const Fp w = host_root_of_unity(ln, true); Fp c = fp::one(); for (size_t i = 0; i < m; ++i) { d[i] = fp::canon(c); c = fp::mul(c, w); } }, &lo))) return rc;
    if ((rc = cached_table(ctx, {25, log_n, 0}, n <= 4096 ? 1 : n / 4096,
                           [](Fp *d, size_t m, int ln, int) { const Fp w = fp::pow_u64(host_root_of_unity(ln, true), 4096); Fp c = fp::one(); for (size_t i = 0; i < m; ++i) { d[i] = fp::canon(c); c = fp::mul(c, w); } }, &hi))) return rc;
    const unsigned int n_chunks = (unsigned int)((row_count + OOD_CHUNK - 1) / OOD_CHUNK);
    Fp *d_w = nullptr, *d_part = nullptr, *d_out = nullptr;
    OodTap *d_taps = nullptr;
    int2 *d_where = nullptr;
    cudaError_t ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_w), row_count * sizeof(Fp));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_part), taps.size() * (size_t)n_chunks * 2 * sizeof(Fp));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_out), n_evals * sizeof(Fp));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_taps), taps.size() * sizeof(OodTap));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_where), n_evals * sizeof(int2));
    auto cleanup = [&] { dev_free(ctx, d_w); dev_free(ctx, d_part); dev_free(ctx, d_out); dev_free(ctx, d_taps); dev_free(ctx, d_where); };
    if (ce != cudaSuccess) { cleanup(); return fail(ctx, SS_ERR_OOM, "ss_ood_eval: %s", cudaGetErrorString(ce)); }
    cudaMemcpy(d_taps, taps.data(), taps.size() * sizeof(OodTap), cudaMemcpyHostToDevice);
    cudaMemcpy(d_where, where.data(), n_evals * sizeof(int2), cudaMemcpyHostToDevice);
    const int wrows = inv_rows_for(row_count);
    const unsigned long long wchunk = (unsigned long long)INV_THREADS * wrows;
    bary_weights_kernel<<<(unsigned)((row_count + wchunk - 1) / wchunk), INV_THREADS>>>(d_w, row_begin, row_count, wrows, z, lo, hi);
    if ((unsigned long long)taps.size() * n_chunks > 0x7fffffffull) { cleanup(); return fail(ctx, SS_ERR_UNSUPPORTED, "ss_ood_eval: grid too large"); }
    ood_dot_kernel<<<(unsigned)(taps.size() * n_chunks), OOD_THREADS>>>(static_cast<const Fp *>(d_trace_cols), col_stride, log_n, d_w, row_begin,
                                                                           row_count, d_taps, (unsigned)taps.size(), n_chunks, d_part);
    ood_finish_kernel<<<(unsigned)n_evals, 256>>>(d_part, n_chunks, d_where, scale, d_out);
    ctx->launches += 3;
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaMemcpy(h_out, d_out, n_evals * sizeof(Fp), cudaMemcpyDeviceToHost);
    cleanup();
    if (ce != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_ood_eval: %s", cudaGetErrorString(ce));
    return SS_OK;
}

ss_status ss_poly_eval(ss_ctx *ctx, ss_field field, const void *d_coeffs, uint64_t coeff_stride, int log_n, int natural_order,
                       const int32_t *h_cols, const void *h_points, size_t n_evals, void *h_out) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_poly_eval: field %d not built", (int)field);
    if (!d_coeffs || log_n < 0 || log_n > 40 || (n_evals && (!h_cols || !h_points || !h_out)) || coeff_stride < (1ull << log_n))
        return fail(ctx, SS_ERR_INVALID, "ss_poly_eval: bad arguments");
    if (n_evals == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    const unsigned long long n = 1ull << log_n;
    const Fp *coeffs = static_cast<const Fp *>(d_coeffs);
    const Fp ginv = fp::inv(fp::from_u32(3));
    // multipliers per job.  ss_lde format (coefficient k scaled by 3^k at position brev(k)): level l
    // (0 = adjacent pairs) uses w^(n / 2^(l+1)) with w = z / 3.  Natural order, plain coefficients:
    // adjacent pairs are (a[2m], a[2m+1]) -> a[2m] + z a[2m+1], so level l uses z^(2^l).
    std::vector<Fp> mults(n_evals * (size_t)(log_n ? log_n : 1));
    for (size_t e = 0; e < n_evals; ++e) {
        const Fp z = load_host(static_cast<const uint8_t *>(h_points) + 32 * e);
        Fp pw = fp::canon(natural_order ? z : fp::mul(z, ginv));
        for (int s2 = 0; s2 < log_n; ++s2) {                  // pw = w^(2^s2)
            mults[e * log_n + (natural_order ? s2 : log_n - 1 - s2)] = pw;
            pw = fp::canon(fp::sqr(pw));
        }
    }
    // stage buffers: ping-pong of ceil(n/2048) partials per job
    const unsigned long long part0 = (n + 2047) / 2048;
    Fp *d_mult = nullptr, *d_a = nullptr, *d_b = nullptr;
    EvalJob *d_jobs = nullptr;
    cudaError_t ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_mult), mults.size() * sizeof(Fp));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_a), n_evals * part0 * sizeof(Fp));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_b), n_evals * ((part0 + 2047) / 2048) * sizeof(Fp));
    if (ce == cudaSuccess) ce = dev_alloc(ctx, reinterpret_cast<void **>(&d_jobs), n_evals * sizeof(EvalJob));
    if (ce != cudaSuccess) {
        dev_free(ctx, d_mult); dev_free(ctx, d_a); dev_free(ctx, d_b); dev_free(ctx, d_jobs);
        return fail(ctx, SS_ERR_OOM, "ss_poly_eval: %s", cudaGetErrorString(ce));
    }
    cudaMemcpy(d_mult, mults.data(), mults.size() * sizeof(Fp), cudaMemcpyHostToDevice);
    std::vector<EvalJob> jobs(n_evals);
    unsigned long long n_in = n;
    int level = 0;
    Fp *cur_out = d_a, *other = d_b;
    const Fp *final_src = nullptr;
    bool first = true;
    while (true) {
        const unsigned long long blocks = (n_in + 2047) / 2048;
        const int levels = (log_n - level) < 11 ? (log_n - level) : 11;
        for (size_t e = 0; e < n_evals; ++e) {
            jobs[e].src = first ? coeffs + (unsigned long long)h_cols[e] * coeff_stride : (cur_out == d_a ? d_b : d_a) + e * n_in;
            jobs[e].dst = cur_out + e * blocks;
            jobs[e].mult = d_mult + e * (size_t)(log_n ? log_n : 1) + level;
        }
        cudaMemcpy(d_jobs, jobs.data(), n_evals * sizeof(EvalJob), cudaMemcpyHostToDevice);
        dim3 grid((unsigned)blocks, (unsigned)n_evals, 1);
        poly_fold_kernel<<<grid, 256>>>(d_jobs, n_in, levels);
        ctx->launches++;
        level += levels;
        final_src = cur_out;
        n_in = blocks;
        first = false;
        if (level >= log_n) break;
        Fp *t = cur_out; cur_out = other; other = t;
    }
    // n_in == 1 per job now (blocks of the last stage == 1)
    std::vector<Fp> res(n_evals);
    ce = cudaMemcpy(res.data(), final_src, n_evals * sizeof(Fp), cudaMemcpyDeviceToHost);
    dev_free(ctx, d_mult); dev_free(ctx, d_a); dev_free(ctx, d_b); dev_free(ctx, d_jobs);
    if (ce != cudaSuccess) return fail(ctx, SS_ERR_CUDA, "ss_poly_eval: %s", cudaGetErrorString(ce));
    memcpy(h_out, res.data(), n_evals * sizeof(Fp));
    return SS_OK;
}

}  // extern "C"
