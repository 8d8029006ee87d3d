// Goldilocks NTT / LDE (SURVEY.md §8 a2/a3 for the single-limb field of BASELINE config 4).
//
// A transform of size 2^L runs as 1-3 launches of gl_pass_kernel.  One pass performs the radix-2 stages of S
// consecutive index bits [b0, b0 + S) on tiles staged in shared memory: a tile holds the 2^S elements that differ
// only in those bits, for W = 2^c adjacent values of the lower bits, so global accesses are runs of 8 W bytes
// (contiguous tiles when b0 = 0).  DIF (Gentleman-Sande) maps natural -> bit-reversed order and walks the bit groups
// from the top; DIT (Cooley-Tukey) maps bit-reversed -> natural and walks them from the bottom — the same split as
// the Fp252 kernels (ntt_fp252.cuh), so ss_lde = inverse DIF, scale, zero-padded forward DIT with no permutation.
// Twiddles w_N^e come from two cached tables (e mod 4096, e div 4096): one extra multiplication instead of an
// N/2-entry table.  Algorithmic bytes: 2 * 8 B per element per pass.
#include "ctx.h"
#include "goldilocks.cuh"

using namespace ss;

namespace {

constexpr int GL_THREADS = 256;
constexpr int GL_LOG_TILE = 12;                 // 4096 elements = 32 KB of shared memory per block
constexpr int GL_LOG_W = 3;                     // strided passes move runs of 8 elements (64 bytes)
constexpr int T_GL_LO = 40, T_GL_HI = 41, T_GL_PLO = 42, T_GL_PHI = 43;

struct GlPassArgs {
    uint64_t *data;                             // column 0
    unsigned long long stride;                  // elements between columns
    int log_n, b0, S, c;                        // bit group [b0, b0 + S), W = 2^c
    const uint64_t *tw_lo, *tw_hi;              // w^e (e < 4096), w^(4096 e)
};

__device__ __forceinline__ uint64_t twiddle(const GlPassArgs &A, unsigned long long e) {
    uint64_t w = __ldg(A.tw_lo + (e & 4095ull));
    if (e >> 12) w = gl::mul(w, __ldg(A.tw_hi + (e >> 12)));
    return w;
}

template <bool DIT>
__global__ void __launch_bounds__(GL_THREADS) gl_pass_kernel(const GlPassArgs A) {
    extern __shared__ uint64_t sm[];
    const int S = A.S, c = A.c, b0 = A.b0, L = A.log_n;
    const unsigned int W = 1u << c, T = 1u << (S + c);
    uint64_t *col = A.data + (unsigned long long)blockIdx.y * A.stride;
    const unsigned long long tile = blockIdx.x;
    const unsigned long long lo_blocks = 1ull << (b0 - c);
    const unsigned long long hi = tile / lo_blocks, lo_base = (tile % lo_blocks) << c;
    const unsigned long long base = (hi << (b0 + S)) | lo_base;
    for (unsigned int e = threadIdx.x; e < T; e += GL_THREADS) {
        const unsigned int j = e >> c, l = e & (W - 1);
        sm[e] = col[base | ((unsigned long long)j << b0) | l];
    }
    __syncthreads();
    for (int step = 0; step < S; ++step) {
        const int sl = DIT ? step : S - 1 - step;            // local stage: pairs (j, j + 2^sl)
        const unsigned int h = 1u << sl;
        const int s = b0 + sl;                                // global bit of the stage
        for (unsigned int q = threadIdx.x; q < T / 2; q += GL_THREADS) {
            const unsigned int l = q & (W - 1), r = q >> c;
            const unsigned int j = ((r >> sl) << (sl + 1)) | (r & (h - 1));
            const unsigned long long low = ((unsigned long long)(j & (h - 1)) << b0) | lo_base | l;     // i mod 2^s
            const uint64_t w = twiddle(A, low << (L - 1 - s));
            const unsigned int ia = (j << c) | l, ib = ((j + h) << c) | l;
            const uint64_t a = sm[ia], b = sm[ib];
            if (DIT) {
                const uint64_t t = gl::mul(b, w);
                sm[ia] = gl::add(a, t);
                sm[ib] = gl::sub(a, t);
            } else {
                sm[ia] = gl::add(a, b);
                sm[ib] = gl::mul(gl::sub(a, b), w);
            }
        }
        __syncthreads();
    }
    for (unsigned int e = threadIdx.x; e < T; e += GL_THREADS) {
        const unsigned int j = e >> c, l = e & (W - 1);
        col[base | ((unsigned long long)j << b0) | l] = sm[e];
    }
}

// data[pos] *= k * h^idx, idx = pos or brev(pos);  h^idx from two tables
__global__ void gl_scale_kernel(uint64_t *data, unsigned long long stride, int log_n, int brev_idx, uint64_t k,
                                const uint64_t *p_lo, const uint64_t *p_hi) {
    const unsigned long long pos = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (pos >= (1ull << log_n)) return;
    uint64_t *col = data + (unsigned long long)blockIdx.y * stride;
    uint64_t f = k;
    if (p_lo) {
        const unsigned long long idx = brev_idx ? (__brevll(pos) >> (64 - log_n)) : pos;
        f = gl::mul(f, __ldg(p_lo + (idx & 4095ull)));
        if (idx >> 12) f = gl::mul(f, __ldg(p_hi + (idx >> 12)));
    }
    col[pos] = gl::mul(col[pos], f);
}

__global__ void gl_bitrev_kernel(uint64_t *data, unsigned long long stride, int log_n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= (1ull << log_n)) return;
    const unsigned long long j = log_n ? (__brevll(i) >> (64 - log_n)) : 0;
    if (i < j) {
        uint64_t *col = data + (unsigned long long)blockIdx.y * stride;
        const uint64_t t = col[i];
        col[i] = col[j];
        col[j] = t;
    }
}

// dst[pos << log_blowup] = src[pos], zero elsewhere (bit-reversed zero padding of a coefficient vector)
__global__ void gl_expand_kernel(const uint64_t *src, unsigned long long src_stride, uint64_t *dst, unsigned long long dst_stride,
                                 int log_N, int log_blowup) {
    const unsigned long long pos = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (pos >= (1ull << log_N)) return;
    const unsigned long long m = (1ull << log_blowup) - 1;
    dst[(unsigned long long)blockIdx.y * dst_stride + pos] = (pos & m) ? 0ull : src[(unsigned long long)blockIdx.y * src_stride + (pos >> log_blowup)];
}

// cached u64 tables: base^e for e < count_lo and base^(4096 e)
ss_status power_tables(ss_ctx *ctx, int key_lo, int key_hi, int log_n, int variant, uint64_t base, size_t exps, const uint64_t **lo, const uint64_t **hi) {
    const size_t n_lo = exps < 4096 ? (exps ? exps : 1) : 4096, n_hi = exps <= 4096 ? 1 : (exps + 4095) / 4096;
    auto get = [&](int key, size_t count, uint64_t step, const uint64_t **out) -> ss_status {
        auto it = ctx->tables.find({key, log_n, variant});
        if (it != ctx->tables.end()) { *out = static_cast<const uint64_t *>(it->second); return SS_OK; }
        std::vector<uint64_t> h(count);
        uint64_t cur = 1;
        for (size_t i = 0; i < count; ++i) { h[i] = cur; cur = gl::mul(cur, step); }
        void *d = nullptr;
        SS_CUDA_CHECK(ctx, cudaMalloc(&d, count * 8));
        SS_CUDA_CHECK(ctx, cudaMemcpy(d, h.data(), count * 8, cudaMemcpyHostToDevice));
        ctx->tables[{key, log_n, variant}] = d;
        *out = static_cast<const uint64_t *>(d);
        return SS_OK;
    };
    ss_status rc = get(key_lo, n_lo, base, lo);
    if (rc) return rc;
    return get(key_hi, n_hi, gl::pow(base, 4096), hi);
}

// the radix-2 stages of one transform, natural -> bit-reversed (DIF) or bit-reversed -> natural (DIT)
ss_status run_passes(ss_ctx *ctx, uint64_t *data, unsigned long long stride, int n_cols, int log_n, bool inverse, bool dit, cudaStream_t st) {
    if (log_n == 0) return SS_OK;
    uint64_t w = gl::root_of_unity(log_n);
    if (inverse) w = gl::inv(w);
    const uint64_t *lo, *hi;
    ss_status rc = power_tables(ctx, T_GL_LO, T_GL_HI, log_n, inverse ? 1 : 0, w, (size_t)1 << (log_n - 1), &lo, &hi);
    if (rc) return rc;
    // bit groups, lowest first: one contiguous group of up to 12 bits at b0 = 0, strided groups of up to 9 bits above it
    struct Group { int b0, S, c; };
    std::vector<Group> groups;
    const int first = log_n < GL_LOG_TILE ? log_n : GL_LOG_TILE;
    groups.push_back({0, first, 0});
    const int rest = log_n - first, max_s = GL_LOG_TILE - GL_LOG_W;
    if (rest > 0) {
        const int n_g = (rest + max_s - 1) / max_s;
        int b = first;
        for (int g = 0; g < n_g; ++g) {
            const int S = rest / n_g + (g < rest % n_g ? 1 : 0);
            groups.push_back({b, S, GL_LOG_W});
            b += S;
        }
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gl_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << GL_LOG_TILE);
        cudaFuncSetAttribute(gl_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 << GL_LOG_TILE);
        attr_set = true;
    }
    for (size_t k = 0; k < groups.size(); ++k) {
        const Group g = dit ? groups[k] : groups[groups.size() - 1 - k];       // DIT: low bits first; DIF: high bits first
        GlPassArgs A{data, stride, log_n, g.b0, g.S, g.c, lo, hi};
        const dim3 grid((unsigned)(1ull << (log_n - g.S - g.c)), (unsigned)n_cols, 1);
        const size_t smem = (size_t)8 << (g.S + g.c);
        if (dit) gl_pass_kernel<true><<<grid, GL_THREADS, smem, st>>>(A);
        else gl_pass_kernel<false><<<grid, GL_THREADS, smem, st>>>(A);
        ctx->launches++;
    }
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

ss_status scale(ss_ctx *ctx, uint64_t *data, unsigned long long stride, int n_cols, int log_n, bool brev_idx, uint64_t k, bool with_powers,
                bool inverse_powers, cudaStream_t st) {
    const uint64_t *lo = nullptr, *hi = nullptr;
    if (with_powers) {
        const uint64_t h = inverse_powers ? gl::inv(gl::GENERATOR) : gl::GENERATOR;
        ss_status rc = power_tables(ctx, T_GL_PLO, T_GL_PHI, log_n, inverse_powers ? 1 : 0, h, (size_t)1 << log_n, &lo, &hi);
        if (rc) return rc;
    }
    const unsigned long long n = 1ull << log_n;
    gl_scale_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n_cols, 1), 256, 0, st>>>(data, stride, log_n, brev_idx ? 1 : 0, k, lo, hi);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

}  // namespace

namespace ss {

ss_status gl_ntt(ss_ctx *ctx, void *d_cols, uint64_t col_stride, int n_cols, int log_n, int inverse, int coset, ss_order in_order,
                 ss_order out_order, cudaStream_t st) {
    uint64_t *data = static_cast<uint64_t *>(d_cols);
    const bool dit = in_order == SS_ORDER_BITREV;         // DIF eats natural order, DIT eats bit-reversed
    ss_status rc;
    if (!inverse && coset && (rc = scale(ctx, data, col_stride, n_cols, log_n, dit, 1, true, false, st))) return rc;
    if ((rc = run_passes(ctx, data, col_stride, n_cols, log_n, inverse != 0, dit, st))) return rc;
    if (inverse) {
        const uint64_t ninv = gl::inv((1ull << log_n) % gl::P);
        if ((rc = scale(ctx, data, col_stride, n_cols, log_n, !dit, ninv, coset != 0, true, st))) return rc;
    }
    const ss_order produced = dit ? SS_ORDER_NATURAL : SS_ORDER_BITREV;
    if (produced != out_order) {
        const unsigned long long n = 1ull << log_n;
        gl_bitrev_kernel<<<dim3((unsigned)((n + 255) / 256), (unsigned)n_cols, 1), 256, 0, st>>>(data, col_stride, log_n);
        ctx->launches++;
        SS_CUDA_CHECK(ctx, cudaGetLastError());
    }
    return SS_OK;
}

ss_status gl_lde(ss_ctx *ctx, const void *d_trace, uint64_t trace_stride, int n_cols, int log_n, int log_blowup, void *d_lde,
                 uint64_t lde_stride, void *d_coeffs, uint64_t coeff_stride, ss_order out_order, cudaStream_t st) {
    const unsigned long long n = 1ull << log_n, N = n << log_blowup;
    uint64_t *coeffs = static_cast<uint64_t *>(d_coeffs);
    uint64_t cstride = coeff_stride;
    uint64_t *scratch = nullptr;
    if (!coeffs) {
        SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&scratch), (size_t)n_cols * n * 8));
        coeffs = scratch;
        cstride = n;
    }
    SS_CUDA_CHECK(ctx, cudaMemcpy2DAsync(coeffs, cstride * 8, d_trace, trace_stride * 8, n * 8, n_cols, cudaMemcpyDeviceToDevice, st));
    // evaluations (natural) -> coefficients (bit-reversed) * n^-1 * 7^k
    ss_status rc = run_passes(ctx, coeffs, cstride, n_cols, log_n, true, false, st);
    if (!rc) rc = scale(ctx, coeffs, cstride, n_cols, log_n, true, gl::inv(n % gl::P), true, false, st);
    if (!rc) {
        gl_expand_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)n_cols, 1), 256, 0, st>>>(coeffs, cstride, static_cast<uint64_t *>(d_lde), lde_stride,
                                                                                                 log_n + log_blowup, log_blowup);
        ctx->launches++;
        rc = run_passes(ctx, static_cast<uint64_t *>(d_lde), lde_stride, n_cols, log_n + log_blowup, false, true, st);
    }
    if (!rc && out_order == SS_ORDER_BITREV) {
        gl_bitrev_kernel<<<dim3((unsigned)((N + 255) / 256), (unsigned)n_cols, 1), 256, 0, st>>>(static_cast<uint64_t *>(d_lde), lde_stride, log_n + log_blowup);
        ctx->launches++;
    }
    if (scratch) dev_free(ctx, scratch);
    if (rc) return rc;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

}  // namespace ss
