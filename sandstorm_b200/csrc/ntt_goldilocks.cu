// Goldilocks NTT / LDE (SURVEY.md §8 a2/a3 for the single-limb field of BASELINE config 4).
//
// A transform of size 2^L runs as 1-3 launches of gl_pass_kernel.  One pass performs the radix-2 stages of S
// consecutive index bits [b0, b0 + S) on tiles staged in shared memory: a tile holds the 2^S elements that differ
// only in those bits, for W = 2^c adjacent values of the lower bits, so global accesses are runs of 8 W bytes
// (contiguous tiles when b0 = 0).  DIF (Gentleman-Sande) maps natural -> bit-reversed order and walks the bit groups
// from the top; DIT (Cooley-Tukey) maps bit-reversed -> natural and walks them from the bottom — the same split as
// the Fp252 kernels (ntt_fp252.cuh), so ss_lde = inverse DIF, scale, zero-padded forward DIT with no permutation.
// Four-step form: the stages of a pass take their twiddles from ONE 4096-entry table (w_8192^k, L1-resident), and a strided
// pass multiplies every element once by w_N^e between the passes (two cached tables: e mod 4096, e div 4096).  2^24 points =
// 13 + 11 bits: two passes.  Algorithmic bytes: 2 * 8 B per element per pass; the kernel is bound by instruction issue, not
// HBM.
#include "ctx.h"
#include "goldilocks.cuh"

using namespace ss;

namespace {

constexpr int GL_THREADS = 512;
constexpr int GL_LOG_TILE = 13;                 // 8192 elements = 64 KB of shared memory per block
constexpr int GL_LOG_W = 2;                     // strided passes move runs of 4 elements (one 32-byte sector)
constexpr int T_GL_LO = 40, T_GL_HI = 41, T_GL_PLO = 42, T_GL_PHI = 43, T_GL_LOCAL = 44;

struct GlPassArgs {
    uint64_t *data;                             // column 0
    unsigned long long stride;                  // elements between columns
    int log_n, b0, S, c;                        // bit group [b0, b0 + S), W = 2^c
    const uint64_t *tw_local;                   // w_(2^GL_LOG_TILE)^k, k < 2^(GL_LOG_TILE - 1): every in-tile twiddle
    const uint64_t *tw_lo, *tw_hi;              // w_N^e (e < 4096), w_N^(4096 e): the twiddles between the passes
};

// Four-step form: a pass is a plain size-2^S transform of the tile's index bits (twiddles from ONE small table, no
// multiplication to build them) and, for a strided pass, one multiplication per element by w_M^(lo * brev_S(l)), M = 2^(b0+S),
// lo = the index bits below the group, l = the element's (bit-reversed) position inside the group — after the stages for DIF,
// before them for DIT (the transposed factorisation).
__device__ __forceinline__ uint64_t interpass(const GlPassArgs &A, unsigned int j, unsigned long long lo) {
    const unsigned long long r = (unsigned long long)(__brev(j) >> (32 - A.S));
    const unsigned long long e = (lo * r) << (A.log_n - A.b0 - A.S);
    uint64_t w = __ldg(A.tw_lo + (e & 4095ull));
    if (e >> 12) w = gl::mul(w, __ldg(A.tw_hi + (e >> 12)));
    return w;
}

template <bool DIT>
__global__ void __launch_bounds__(GL_THREADS) gl_pass_kernel(const GlPassArgs A) {
    extern __shared__ uint64_t sm[];
    const int S = A.S, c = A.c, b0 = A.b0;
    const unsigned int W = 1u << c, T = 1u << (S + c);
    uint64_t *col = A.data + (unsigned long long)blockIdx.y * A.stride;
    const unsigned long long tile = blockIdx.x;
    const unsigned long long lo_blocks = 1ull << (b0 - c);
    const unsigned long long hi = tile / lo_blocks, lo_base = (tile % lo_blocks) << c;
    const unsigned long long base = (hi << (b0 + S)) | lo_base;
    for (unsigned int e = threadIdx.x; e < T; e += GL_THREADS) {
        const unsigned int j = e >> c, l = e & (W - 1);
        uint64_t v = col[base | ((unsigned long long)j << b0) | l];
        if (DIT && b0 > 0) v = gl::mul(v, interpass(A, j, lo_base | l));
        sm[e] = v;
    }
    __syncthreads();
    // two stages per round in registers (four elements per thread and round: half the shared-memory traffic and barriers of a
    // radix-2 sweep); a last single stage when S is odd
    int done = 0;
    while (done < S) {
        if (S - done >= 2) {
            const int s0 = DIT ? done : S - 2 - done, s1 = s0 + 1;           // the lower / upper stage of the round
            const unsigned int h0 = 1u << s0;
            const int sh0 = GL_LOG_TILE - 1 - s0, sh1 = GL_LOG_TILE - 1 - s1;
            for (unsigned int q = threadIdx.x; q < T / 4; q += GL_THREADS) {
                const unsigned int l = q & (W - 1), r = q >> c;
                const unsigned int k = r & (h0 - 1);
                const unsigned int j = ((r >> s0) << (s0 + 2)) | k;
                const unsigned int i0 = (j << c) | l, i1 = ((j + h0) << c) | l, i2 = ((j + 2 * h0) << c) | l, i3 = ((j + 3 * h0) << c) | l;
                uint64_t x0 = sm[i0], x1 = sm[i1], x2 = sm[i2], x3 = sm[i3];
                const uint64_t w1a = __ldg(A.tw_local + ((unsigned long long)k << sh1));
                const uint64_t w1b = __ldg(A.tw_local + ((unsigned long long)(k + h0) << sh1));
                if (DIT) {
                    if (s0 != 0) {
                        const uint64_t w0 = __ldg(A.tw_local + ((unsigned long long)k << sh0));
                        x1 = gl::mul(x1, w0);
                        x3 = gl::mul(x3, w0);
                    }
                    const uint64_t a = gl::add(x0, x1), bb = gl::sub(x0, x1), cc = gl::add(x2, x3), d = gl::sub(x2, x3);
                    const uint64_t tc = gl::mul(cc, w1a), td = gl::mul(d, w1b);
                    x0 = gl::add(a, tc); x2 = gl::sub(a, tc);
                    x1 = gl::add(bb, td); x3 = gl::sub(bb, td);
                } else {
                    const uint64_t a = gl::add(x0, x2), cc = gl::mul(gl::sub(x0, x2), w1a);
                    const uint64_t bb = gl::add(x1, x3), d = gl::mul(gl::sub(x1, x3), w1b);
                    x0 = gl::add(a, bb); x1 = gl::sub(a, bb);
                    x2 = gl::add(cc, d); x3 = gl::sub(cc, d);
                    if (s0 != 0) {
                        const uint64_t w0 = __ldg(A.tw_local + ((unsigned long long)k << sh0));
                        x1 = gl::mul(x1, w0);
                        x3 = gl::mul(x3, w0);
                    }
                }
                sm[i0] = x0; sm[i1] = x1; sm[i2] = x2; sm[i3] = x3;
            }
            done += 2;
        } else {
            const int sl = DIT ? done : 0;                                   // the one stage left: the top one (DIT) or stage 0 (DIF)
            const unsigned int h = 1u << sl;
            for (unsigned int q = threadIdx.x; q < T / 2; q += GL_THREADS) {
                const unsigned int l = q & (W - 1), r = q >> c;
                const unsigned int k = r & (h - 1);
                const unsigned int j = ((r >> sl) << (sl + 1)) | k;
                const unsigned int ia = (j << c) | l, ib = ((j + h) << c) | l;
                const uint64_t a = sm[ia], bv = sm[ib];
                if (sl == 0) {
                    sm[ia] = gl::add(a, bv);
                    sm[ib] = gl::sub(a, bv);
                } else {
                    const uint64_t w = __ldg(A.tw_local + ((unsigned long long)k << (GL_LOG_TILE - 1 - sl)));
                    if (DIT) {
                        const uint64_t t = gl::mul(bv, w);
                        sm[ia] = gl::add(a, t);
                        sm[ib] = gl::sub(a, t);
                    } else {
                        sm[ia] = gl::add(a, bv);
                        sm[ib] = gl::mul(gl::sub(a, bv), w);
                    }
                }
            }
            done += 1;
        }
        __syncthreads();
    }
    for (unsigned int e = threadIdx.x; e < T; e += GL_THREADS) {
        const unsigned int j = e >> c, l = e & (W - 1);
        uint64_t v = sm[e];
        if (!DIT && b0 > 0) v = gl::mul(v, interpass(A, j, lo_base | l));
        col[base | ((unsigned long long)j << b0) | l] = v;
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
    const uint64_t *lo, *hi, *local, *unused;
    ss_status rc = power_tables(ctx, T_GL_LO, T_GL_HI, log_n, inverse ? 1 : 0, w, (size_t)1 << log_n, &lo, &hi);
    if (rc) return rc;
    uint64_t wt = gl::root_of_unity(GL_LOG_TILE);
    if (inverse) wt = gl::inv(wt);
    // (keyed by the tile size: one table of 2^(GL_LOG_TILE-1) in-tile twiddles per direction, shared by every transform length)
    if ((rc = power_tables(ctx, T_GL_LOCAL, T_GL_LOCAL + 100, GL_LOG_TILE, inverse ? 1 : 0, wt, (size_t)1 << (GL_LOG_TILE - 1), &local, &unused))) return rc;
    // bit groups, lowest first: one contiguous group of up to GL_LOG_TILE bits at b0 = 0, strided groups of up to
    // GL_LOG_TILE - GL_LOG_W bits above it (2^24 = 13 + 11: two passes)
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
            const int c = GL_LOG_TILE - S < b ? GL_LOG_TILE - S : b;      // fill the tile: short groups move longer runs
            groups.push_back({b, S, c});
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
        GlPassArgs A{data, stride, log_n, g.b0, g.S, g.c, local, lo, hi};
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
