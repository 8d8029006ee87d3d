// Host side of the Fp252 NTT / LDE: twiddle + scale tables, pass planning, kernel launches,
// and the ss_ntt / ss_lde entry points (include/sandstorm_b200.h).
#include "ctx.h"
#include "ntt_fp252.cuh"
#include <cstdlib>
#include <cstring>

using namespace ss;

namespace {

enum TableKind : int { T_LOCAL = 1, T_LO = 2, T_HI = 3, T_SCALE_LO = 4, T_SCALE_HI = 5 };
enum ScaleVariant : int { SV_LDE = 0, SV_COSET_FWD = 1, SV_COSET_INV = 2, SV_NINV = 3 };

// w_n = 3^((p-1)/n), n = 2^log_n; inverse -> w_n^-1.  Canonical Montgomery form.
Fp root_of_unity(int log_n, bool inverse) {
    uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0x00000011u, 0x08000000u};   // p - 1
    for (int s = 0; s < log_n; ++s) {
        for (int i = 0; i < 8; ++i) {
            e[i] >>= 1;
            if (i < 7) e[i] |= e[i + 1] << 31;
        }
    }
    Fp w = fp::pow_limbs(fp::from_u32(3), e, 8);
    if (inverse) w = fp::inv(w);
    return fp::canon(w);
}

// dst[i] = c * h^i
void geometric(Fp *dst, size_t n, Fp c, const Fp &h) {
    for (size_t i = 0; i < n; ++i) {
        dst[i] = fp::canon(c);
        c = fp::mul(c, h);
    }
}

// local twiddles are always powers of w_4096 (2048 entries), whatever the tile size
void fill_local(Fp *dst, size_t n, int, int inverse) { geometric(dst, n, fp::one(), root_of_unity(12, inverse != 0)); }
void fill_lo(Fp *dst, size_t n, int log_n, int inverse) { geometric(dst, n, fp::one(), root_of_unity(log_n, inverse != 0)); }
void fill_hi(Fp *dst, size_t n, int log_n, int inverse) {
    const Fp w = root_of_unity(log_n, inverse != 0);
    geometric(dst, n, fp::one(), fp::pow_u64(w, 4096));
}

void scale_params(int log_n, int variant, Fp &c, Fp &h) {
    Fp ninv = fp::inv(fp::pow_u64(fp::from_u32(2), (uint64_t)log_n));
    const Fp g = fp::from_u32(3);          // Fp::GENERATOR, the LDE coset offset
    switch (variant) {
    case SV_LDE: c = ninv; h = g; break;
    case SV_COSET_FWD: c = fp::one(); h = g; break;
    case SV_COSET_INV: c = ninv; h = fp::inv(g); break;
    default: c = ninv; h = fp::one(); break;
    }
}
void fill_scale_lo(Fp *dst, size_t n, int log_n, int variant) {
    Fp c, h;
    scale_params(log_n, variant, c, h);
    geometric(dst, n, c, h);
}
void fill_scale_hi(Fp *dst, size_t n, int log_n, int variant) {
    Fp c, h;
    scale_params(log_n, variant, c, h);
    geometric(dst, n, fp::one(), fp::pow_u64(h, 4096));
}

std::vector<int> plan_passes(int log_n) {
    if (log_n <= NTT_LOG_TILE) return {log_n};
    const int k = (log_n + NTT_LOG_TILE - 1) / NTT_LOG_TILE;
    std::vector<int> out;
    for (int i = 0; i < k; ++i) out.push_back(log_n / k + (i < log_n % k ? 1 : 0));
    return out;
}

struct NttJob {
    Fp *dst;
    const Fp *src;
    uint64_t dst_stride, src_stride;
    int n_cols, log_n;
    bool inverse, dit;
    int expand_log;            // DIT only: first pass reads src[p >> expand_log]
    int pre_scale, post_scale; // NttScaleMode
    int scale_variant;
    bool canon_out;
    const Fp *sc_lo = nullptr, *sc_hi = nullptr;   // explicit two-level scale table (custom_scale) instead of a variant
};

// lo[i] = c * h^i (i < 4096), hi[i] = h^(4096 i) for an arbitrary geometric scale, cached by value
// (one device allocation per table pair: .first owns it, .second points into it)
ss_status custom_scale(ss_ctx *ctx, int log_len, const Fp &c, const Fp &h, const Fp **lo, const Fp **hi) {
    std::vector<uint32_t> key(17);
    key[0] = (uint32_t)log_len;
    for (int i = 0; i < 8; ++i) { key[1 + i] = c.l[i]; key[9 + i] = h.l[i]; }
    const size_t n = (size_t)1 << log_len, n_lo = n < 4096 ? n : 4096, n_hi = n <= 4096 ? 1 : n / 4096;
    auto it = ctx->scale_tables.find(key);
    if (it == ctx->scale_tables.end()) {
        if (ctx->scale_tables.size() >= 48) {       // per-proof scales (ss_coset_eval at z/3) would pile up: start over
            for (auto &kv : ctx->scale_tables) cudaFree(kv.second.first);      // (cudaFree waits for kernels still reading them)
            ctx->scale_tables.clear();
        }
        std::vector<Fp> host(n_lo + n_hi);
        geometric(host.data(), n_lo, c, h);
        geometric(host.data() + n_lo, n_hi, fp::one(), fp::pow_u64(h, 4096));
        void *d = nullptr;
        SS_CUDA_CHECK(ctx, cudaMalloc(&d, host.size() * sizeof(Fp)));
        SS_CUDA_CHECK(ctx, cudaMemcpy(d, host.data(), host.size() * sizeof(Fp), cudaMemcpyHostToDevice));
        it = ctx->scale_tables.emplace(key, std::make_pair(d, static_cast<void *>(nullptr))).first;
    }
    *lo = static_cast<const Fp *>(it->second.first);
    *hi = *lo + n_lo;
    return SS_OK;
}

template <bool DIT>
ss_status launch_pass(ss_ctx *ctx, const NttPass &p, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        SS_CUDA_CHECK(ctx, cudaFuncSetAttribute(ntt_pass_kernel<DIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, NTT_SMEM_BYTES));
        configured = true;
    }
    dim3 grid;
    if (p.log_n < NTT_LOG_TILE) {
        const int per_tile = 1 << (NTT_LOG_TILE - p.log_n);
        grid = dim3((p.n_cols + per_tile - 1) / per_tile, 1, 1);
    } else {
        grid = dim3(1u << (p.log_n - NTT_LOG_TILE), p.n_cols, 1);
    }
    ntt_pass_kernel<DIT><<<grid, NTT_THREADS, NTT_SMEM_BYTES, st>>>(p);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

ss_status run_ntt(ss_ctx *ctx, const NttJob &job, cudaStream_t st) {
    if (job.log_n == 0) {
        if (job.dst != job.src)
            for (int c = 0; c < job.n_cols; ++c)
                SS_CUDA_CHECK(ctx, cudaMemcpyAsync(job.dst + c * job.dst_stride, job.src + c * job.src_stride, sizeof(Fp), cudaMemcpyDeviceToDevice, st));
        return SS_OK;
    }
    const int inv = job.inverse ? 1 : 0;
    Fp *tw_local, *tw_lo, *tw_hi, *sc_lo = nullptr, *sc_hi = nullptr;
    ss_status rc;
    const size_t n = (size_t)1 << job.log_n;
    if ((rc = cached_table(ctx, {T_LOCAL, 12, inv}, 2048, fill_local, &tw_local))) return rc;
    if ((rc = cached_table(ctx, {T_LO, job.log_n, inv}, n < 4096 ? n : 4096, fill_lo, &tw_lo))) return rc;
    if ((rc = cached_table(ctx, {T_HI, job.log_n, inv}, n <= 4096 ? 1 : n / 4096, fill_hi, &tw_hi))) return rc;
    if (job.sc_lo) {
        sc_lo = const_cast<Fp *>(job.sc_lo);
        sc_hi = const_cast<Fp *>(job.sc_hi);
    } else if (job.pre_scale != SCALE_NONE || job.post_scale != SCALE_NONE) {
        const bool is_const = job.pre_scale == SCALE_CONST || job.post_scale == SCALE_CONST;
        if ((rc = cached_table(ctx, {T_SCALE_LO, job.log_n, job.scale_variant}, is_const ? 1 : (n < 4096 ? n : 4096), fill_scale_lo, &sc_lo))) return rc;
        if (!is_const)
            if ((rc = cached_table(ctx, {T_SCALE_HI, job.log_n, job.scale_variant}, n <= 4096 ? 1 : n / 4096, fill_scale_hi, &sc_hi))) return rc;
    }
    const std::vector<int> passes = plan_passes(job.log_n);
    const int np = (int)passes.size();
    int log_b = job.dit ? 0 : job.log_n;
    for (int i = 0; i < np; ++i) {
        const int L = job.dit ? passes[np - 1 - i] : passes[i];
        if (job.dit) log_b += L;
        NttPass p{};
        const bool first = i == 0, last = i == np - 1;
        p.dst = job.dst;
        p.src = first ? job.src : job.dst;
        p.dst_col_stride = job.dst_stride;
        p.src_col_stride = first ? job.src_stride : job.dst_stride;
        p.n_cols = job.n_cols;
        p.log_n = job.log_n;
        p.log_block = log_b;
        p.L = L;
        p.contiguous = (log_b == L) ? 1 : 0;
        p.expand_log = first ? job.expand_log : 0;
        p.tw_local = tw_local;
        p.tw_lo = tw_lo;
        p.tw_hi = tw_hi;
        p.pre_scale = first ? job.pre_scale : SCALE_NONE;
        p.post_scale = last ? job.post_scale : SCALE_NONE;
        p.scale_lo = sc_lo;
        p.scale_hi = sc_hi;
        p.canon_out = (last && job.canon_out) ? 1 : 0;
        rc = job.dit ? launch_pass<true>(ctx, p, st) : launch_pass<false>(ctx, p, st);
        if (rc) return rc;
        if (!job.dit) log_b -= L;
    }
    return SS_OK;
}

// in-place bit-reversal permutation of every column
__global__ void bitrev_permute_kernel(Fp *cols, unsigned long long stride, int log_n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    const unsigned long long j = __brevll(i) >> (64 - log_n);
    if (i < j) {
        uint4 *a = reinterpret_cast<uint4 *>(cols + blockIdx.y * stride + i);
        uint4 *b = reinterpret_cast<uint4 *>(cols + blockIdx.y * stride + j);
        const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
        a[0] = b0; a[1] = b1; b[0] = a0; b[1] = a1;
    }
}

ss_status bitrev_permute(ss_ctx *ctx, Fp *cols, uint64_t stride, int n_cols, int log_n, cudaStream_t st) {
    if (log_n < 2) return SS_OK;
    const unsigned long long n = 1ull << log_n;
    dim3 grid((unsigned)((n + 255) / 256), n_cols, 1);
    bitrev_permute_kernel<<<grid, 256, 0, st>>>(cols, stride, log_n);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

// ---- the size-W transform ACROSS ranks of a sharded NTT (W = 2, 4, 8) -------------------------------------------------
// out[k1][i] = tw(i)^k1 * sum_{j1 < W} in[j1][i] * wW^(j1 k1),   i < count,   tw(i) = base^(tw_offset + i)  (or 1)
// in / out are W slabs each (base pointer + slab stride).  One thread per i: W loads, a radix-2 network in registers
// ((W/2) log2 W - trivial multiplications), at most 2 (W - 1) twiddle multiplications, W stores.  HBM-bound:
// 2 * W * 32 B per thread.
struct ShardDftArgs {
    const Fp *in; Fp *out;
    unsigned long long in_stride, out_stride, count, tw_offset;
    Fp wpow[4];                     // wW^e, e < W/2 (direction applied)
    const Fp *tw_lo, *tw_hi;        // base^i (i < 4096), base^(4096 i); nullptr = no twiddle
};

template <int LOG_W>
__global__ void __launch_bounds__(256) shard_dft_kernel(const ShardDftArgs A) {
    constexpr int W = 1 << LOG_W;
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= A.count) return;
    Fp x[W];
#pragma unroll
    for (int j = 0; j < W; ++j) x[j] = nttk::ld_stream(A.in + (unsigned long long)j * A.in_stride + i);
    // DIF network: natural in, bit-reversed out
#pragma unroll
    for (int len = W / 2; len >= 1; len >>= 1) {
#pragma unroll
        for (int b = 0; b < W; b += 2 * len) {
#pragma unroll
            for (int k = 0; k < len; ++k) {
                const Fp a = x[b + k], c = x[b + k + len];
                x[b + k] = fp::add(a, c);
                Fp d = fp::sub(a, c);
                const int e = k * (W / (2 * len));
                if (e) d = fp::mul(d, A.wpow[e]);
                x[b + k + len] = d;
            }
        }
    }
    Fp t, tk;
    if (A.tw_lo) { t = nttk::two_level_tw(A.tw_lo, A.tw_hi, A.tw_offset + i); tk = t; }
#pragma unroll
    for (int k1 = 0; k1 < W; ++k1) {
        int pos = 0;                                           // x[brev(k1)] holds output k1
#pragma unroll
        for (int bit = 0; bit < LOG_W; ++bit) pos |= ((k1 >> bit) & 1) << (LOG_W - 1 - bit);
        Fp v = x[pos];
        if (A.tw_lo && k1 >= 1) {
            v = fp::mul(v, tk);
            if (k1 + 1 < W) tk = fp::mul(tk, t);
        }
        nttk::st_stream(A.out + (unsigned long long)k1 * A.out_stride + i, fp::canon(v));
    }
}

}  // namespace

extern "C" {

ss_status ss_ntt(ss_ctx *ctx, ss_field field, void *d_cols, uint64_t col_stride, int n_cols, int log_n,
                 int inverse, int coset, ss_order in_order, ss_order out_order, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252 && field != SS_FIELD_GOLDILOCKS) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_ntt: field %d not built", (int)field);
    if (!d_cols || n_cols < 0 || log_n < 0 || log_n > 40 || col_stride < (1ull << log_n))
        return fail(ctx, SS_ERR_INVALID, "ss_ntt: bad arguments (n_cols=%d log_n=%d stride=%llu)", n_cols, log_n, (unsigned long long)col_stride);
    if (n_cols == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    if (field == SS_FIELD_GOLDILOCKS) {
        if (log_n > 32) return fail(ctx, SS_ERR_INVALID, "ss_ntt: Goldilocks has 2-adicity 32");
        return gl_ntt(ctx, d_cols, col_stride, n_cols, log_n, inverse, coset, in_order, out_order, st);
    }
    NttJob job{};
    job.dst = static_cast<Fp *>(d_cols);
    job.src = job.dst;
    job.dst_stride = job.src_stride = col_stride;
    job.n_cols = n_cols;
    job.log_n = log_n;
    job.inverse = inverse != 0;
    job.dit = in_order == SS_ORDER_BITREV;          // DIF eats natural order, DIT eats bit-reversed
    job.canon_out = true;
    // position -> coefficient/evaluation index is brev(pos) on the bit-reversed side of the transform
    if (!inverse) {
        if (coset) { job.pre_scale = job.dit ? SCALE_TABLE_BREV : SCALE_TABLE_NAT; job.scale_variant = SV_COSET_FWD; }
    } else {
        if (coset) { job.post_scale = job.dit ? SCALE_TABLE_NAT : SCALE_TABLE_BREV; job.scale_variant = SV_COSET_INV; }
        else       { job.post_scale = SCALE_CONST; job.scale_variant = SV_NINV; }
    }
    ss_status rc = run_ntt(ctx, job, st);
    if (rc) return rc;
    const ss_order produced = job.dit ? SS_ORDER_NATURAL : SS_ORDER_BITREV;
    if (produced != out_order) return bitrev_permute(ctx, job.dst, col_stride, n_cols, log_n, st);
    return SS_OK;
}

ss_status ss_lde(ss_ctx *ctx, ss_field field, const void *d_trace, uint64_t trace_stride, int n_cols,
                 int log_n, int log_blowup, void *d_lde, uint64_t lde_stride, void *d_coeffs,
                 uint64_t coeff_stride, ss_order out_order, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252 && field != SS_FIELD_GOLDILOCKS) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_lde: field %d not built", (int)field);
    if (!d_trace || !d_lde || n_cols < 0 || log_n < 0 || log_blowup < 0 || log_n + log_blowup > 40 ||
        trace_stride < (1ull << log_n) || lde_stride < (1ull << (log_n + log_blowup)) ||
        (d_coeffs && coeff_stride < (1ull << log_n)))
        return fail(ctx, SS_ERR_INVALID, "ss_lde: bad arguments");
    if (n_cols == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    if (field == SS_FIELD_GOLDILOCKS) {
        if (log_n + log_blowup > 32) return fail(ctx, SS_ERR_INVALID, "ss_lde: Goldilocks has 2-adicity 32");
        return gl_lde(ctx, d_trace, trace_stride, n_cols, log_n, log_blowup, d_lde, lde_stride, d_coeffs, coeff_stride, out_order, st);
    }
    const size_t n = (size_t)1 << log_n;
    Fp *coeffs = static_cast<Fp *>(d_coeffs);
    uint64_t cstride = coeff_stride;
    if (!coeffs) {
        void *scratch;
        ss_status rc = scratch_reserve(ctx, (size_t)n_cols * n * sizeof(Fp), &scratch);
        if (rc) return rc;
        coeffs = static_cast<Fp *>(scratch);
        cstride = n;
    }
    // 1) inverse DIF: evaluations (natural) -> coefficients (bit-reversed) * n^-1 * 3^k
    NttJob inv{};
    inv.dst = coeffs; inv.src = static_cast<const Fp *>(d_trace);
    inv.dst_stride = cstride; inv.src_stride = trace_stride;
    inv.n_cols = n_cols; inv.log_n = log_n; inv.inverse = true; inv.dit = false;
    inv.post_scale = SCALE_TABLE_BREV; inv.scale_variant = SV_LDE; inv.canon_out = true;
    ss_status rc = run_ntt(ctx, inv, st);
    if (rc) return rc;
    // 2) forward DIT of size N; bit-reversed zero padding = coefficient p sits at position p << log_blowup
    NttJob fwd{};
    fwd.dst = static_cast<Fp *>(d_lde); fwd.src = coeffs;
    fwd.dst_stride = lde_stride; fwd.src_stride = cstride;
    fwd.n_cols = n_cols; fwd.log_n = log_n + log_blowup; fwd.inverse = false; fwd.dit = true;
    fwd.expand_log = log_blowup; fwd.canon_out = true;
    if (log_blowup == 0 && log_n == 0) {
        // degenerate: single element, DIT job with log_n 0 copies src -> dst
    }
    rc = run_ntt(ctx, fwd, st);
    if (rc) return rc;
    if (out_order == SS_ORDER_BITREV) return bitrev_permute(ctx, fwd.dst, lde_stride, n_cols, log_n + log_blowup, st);
    return SS_OK;
}

/* Local part of a transform that is sharded over W ranks (sandstorm_b200/parallel.py): the size-m
 * transforms between the two all-to-all exchanges, with the geometric scale factors the decomposition needs.
 *   stages & 1: inverse DIF of size m = 2^log_m, natural -> bit-reversed, output k scaled by c0 * h0^k
 *               (1/M, the coset shift of this rank's residue class, ... all folded into (c0, h0))
 *   stages & 2: forward DIT of size m << log_expand from bit-reversed coefficients (zero-padded by interleaving),
 *               natural output k2 scaled by tw^k2 (the inter-rank twiddle w_M^(j1 k2); NULL = none)
 * stages = 3 runs both through the context's scratch buffer (the local LDE).  d_dst may equal d_src for stages 1 / 2
 * without expansion. */
ss_status ss_ntt_shard(ss_ctx *ctx, ss_field field, const void *d_src, uint64_t src_stride, int n_cols, int log_m, int stages,
                       int log_expand, const void *h_c0, const void *h_h0, const void *h_tw, void *d_dst, uint64_t dst_stride,
                       void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_ntt_shard: field %d not built", (int)field);
    if (!d_src || !d_dst || n_cols < 0 || log_m < 0 || log_expand < 0 || log_m + log_expand > 40 || stages < 1 || stages > 3 ||
        ((stages & 1) && (!h_c0 || !h_h0)) || (!(stages & 2) && log_expand) || src_stride < (1ull << log_m) ||
        dst_stride < (1ull << (log_m + log_expand)))
        return fail(ctx, SS_ERR_INVALID, "ss_ntt_shard: bad arguments");
    if (n_cols == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    auto load = [](const void *p) { Fp v; memcpy(v.l, p, 32); return fp::canon(v); };
    const size_t m = (size_t)1 << log_m;
    const Fp *coeffs = static_cast<const Fp *>(d_src);
    uint64_t cstride = src_stride;
    ss_status rc;
    if (stages & 1) {
        NttJob inv{};
        inv.src = static_cast<const Fp *>(d_src); inv.src_stride = src_stride;
        if (stages & 2) {
            void *scratch;
            if ((rc = scratch_reserve(ctx, (size_t)n_cols * m * sizeof(Fp), &scratch))) return rc;
            inv.dst = static_cast<Fp *>(scratch); inv.dst_stride = m;
        } else {
            inv.dst = static_cast<Fp *>(d_dst); inv.dst_stride = dst_stride;
        }
        inv.n_cols = n_cols; inv.log_n = log_m; inv.inverse = true; inv.dit = false;
        inv.post_scale = SCALE_TABLE_BREV; inv.canon_out = true;
        if ((rc = custom_scale(ctx, log_m, load(h_c0), load(h_h0), &inv.sc_lo, &inv.sc_hi))) return rc;
        if (log_m == 0) {
            // a single element: run_ntt copies; apply the scale c0 through a size-1 "table" is not wired -> handle on the host side
            return fail(ctx, SS_ERR_UNSUPPORTED, "ss_ntt_shard: log_m must be >= 1");
        }
        if ((rc = run_ntt(ctx, inv, st))) return rc;
        coeffs = inv.dst; cstride = inv.dst_stride;
    }
    if (stages & 2) {
        NttJob fwd{};
        fwd.dst = static_cast<Fp *>(d_dst); fwd.src = coeffs;
        fwd.dst_stride = dst_stride; fwd.src_stride = cstride;
        fwd.n_cols = n_cols; fwd.log_n = log_m + log_expand; fwd.inverse = false; fwd.dit = true;
        fwd.expand_log = log_expand; fwd.canon_out = true;
        if (h_tw) {
            fwd.post_scale = SCALE_TABLE_NAT;
            if ((rc = custom_scale(ctx, log_m + log_expand, fp::one(), load(h_tw), &fwd.sc_lo, &fwd.sc_hi))) return rc;
        }
        if ((rc = run_ntt(ctx, fwd, st))) return rc;
    }
    return SS_OK;
}

/* Evaluations of polynomials on an arbitrary coset h<w_n> from the coefficient vectors ss_lde leaves behind (d_coeffs:
 * coefficient k, possibly pre-scaled, at position brev(k)):  dst[j] = sum_k coeff_k * h^k * w_n^(j k), natural order —
 * one forward transform with the scale fused into its first pass.  With ss_lde's coset-scaled coefficients (c_k 3^k) and
 * h = z / 3 this is T(z g^j) for EVERY j at once: the out-of-domain mask values of a column with many taps
 * (air.trace_arguments(): up to 105 offsets of one column in the starknet layout) for the price of one NTT instead of one
 * n-term sum per tap.  d_dst may equal d_coeffs. */
ss_status ss_coset_eval(ss_ctx *ctx, ss_field field, const void *d_coeffs, uint64_t coeff_stride, int n_cols, int log_n, const void *h_h,
                        void *d_dst, uint64_t dst_stride, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_coset_eval: field %d not built", (int)field);
    if (!d_coeffs || !d_dst || !h_h || n_cols < 0 || log_n < 1 || log_n > 40 || coeff_stride < (1ull << log_n) || dst_stride < (1ull << log_n))
        return fail(ctx, SS_ERR_INVALID, "ss_coset_eval: bad arguments");
    if (n_cols == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    Fp h;
    memcpy(h.l, h_h, 32);
    NttJob fwd{};
    fwd.dst = static_cast<Fp *>(d_dst); fwd.src = static_cast<const Fp *>(d_coeffs);
    fwd.dst_stride = dst_stride; fwd.src_stride = coeff_stride;
    fwd.n_cols = n_cols; fwd.log_n = log_n; fwd.inverse = false; fwd.dit = true;
    fwd.canon_out = true;
    fwd.pre_scale = SCALE_TABLE_BREV;
    ss_status rc = custom_scale(ctx, log_n, fp::one(), fp::canon(h), &fwd.sc_lo, &fwd.sc_hi);
    if (rc) return rc;
    return run_ntt(ctx, fwd, pick_stream(ctx, stream));
}

/* The size-W transform across the W ranks of a sharded NTT (W = 2^log_w <= 8):
 *   out[k1 * out_stride + i] = tw(i)^k1 * sum_{j1 < W} in[j1 * in_stride + i] * w_W^(+-j1 k1),    i < count,
 * with tw(i) = w_M^(+-(tw_offset + i)), M = 2^tw_log_m (tw_log_m < 0: no twiddle); inverse selects the sign of both roots. */
ss_status ss_shard_dft(ss_ctx *ctx, ss_field field, const void *d_in, uint64_t in_stride, void *d_out, uint64_t out_stride,
                       uint64_t count, int log_w, int inverse, int tw_log_m, uint64_t tw_offset, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_shard_dft: field %d not built", (int)field);
    if (!d_in || !d_out || log_w < 1 || log_w > 3 || tw_log_m > 40 || (tw_log_m >= 0 && tw_offset + count > (1ull << tw_log_m)))
        return fail(ctx, SS_ERR_INVALID, "ss_shard_dft: bad arguments");
    if (count == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    ShardDftArgs A{};
    A.in = static_cast<const Fp *>(d_in); A.out = static_cast<Fp *>(d_out);
    A.in_stride = in_stride; A.out_stride = out_stride; A.count = count; A.tw_offset = tw_offset;
    const Fp w = root_of_unity(log_w, inverse != 0);
    Fp acc = fp::one();
    for (int e = 0; e < 4; ++e) { A.wpow[e] = fp::canon(acc); acc = fp::mul(acc, w); }
    if (tw_log_m >= 0) {
        Fp *lo, *hi;
        ss_status rc;
        const size_t M = (size_t)1 << tw_log_m;
        const int inv = inverse ? 1 : 0;
        if ((rc = cached_table(ctx, {T_LO, tw_log_m, inv}, M < 4096 ? M : 4096, fill_lo, &lo))) return rc;
        if ((rc = cached_table(ctx, {T_HI, tw_log_m, inv}, M <= 4096 ? 1 : M / 4096, fill_hi, &hi))) return rc;
        A.tw_lo = lo; A.tw_hi = hi;
    }
    cudaStream_t st = pick_stream(ctx, stream);
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (log_w == 1) shard_dft_kernel<1><<<grid, 256, 0, st>>>(A);
    else if (log_w == 2) shard_dft_kernel<2><<<grid, 256, 0, st>>>(A);
    else shard_dft_kernel<3><<<grid, 256, 0, st>>>(A);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    return SS_OK;
}

}  // extern "C"
