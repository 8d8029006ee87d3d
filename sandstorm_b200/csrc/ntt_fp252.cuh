// Batched Fp252 NTT for sm_100a: multi-pass "four-step" decomposition, each pass = one kernel that
// stages a 2^NTT_LOG_TILE-element tile (2048 elements = 64 KB, two CTAs per SM) in shared memory and
// runs up to NTT_LOG_TILE radix-2 stages on it with register radix-8 rounds (8 elements per thread,
// conflict-free swizzled smem).
//
//   DIF (natural -> bit-reversed):  top bits first; strided tiles; post-multiply by the inter-pass
//        twiddle w_B^(lo * brev_L(m)); last pass contiguous tiles.
//   DIT (bit-reversed -> natural):  mirror image (contiguous pass first, pre-twiddle, strided last).
//
// The algorithm (plan, tile geometry, twiddle exponents, fused LDE) is the one modelled and checked
// against a naive DFT in tools/ntt_model.py.  Semantics = ark-poly Radix2EvaluationDomain fft/ifft
// (SURVEY.md §8 a2/a3); algorithmic HBM bytes per pass = 2 * 32 B per element.
#pragma once
#include "fp252.cuh"

namespace ss {

// Tile geometry (A/B measured on B200, profiles/r01_ntt_tile_ab.md): 2048-element tiles with two
// resident CTAs per SM beat 4096 x 1, 1024 x 4 and 512 x 8.
#ifndef SS_NTT_LOG_TILE
#define SS_NTT_LOG_TILE 11
#endif
#ifndef SS_NTT_MIN_CTAS
#define SS_NTT_MIN_CTAS 2
#endif
constexpr int NTT_LOG_TILE = SS_NTT_LOG_TILE;
constexpr int NTT_TILE = 1 << NTT_LOG_TILE;
constexpr int NTT_THREADS = NTT_TILE / 8;
constexpr int NTT_SMEM_BYTES = NTT_TILE * 32;

enum NttScaleMode : int {
    SCALE_NONE = 0,
    SCALE_CONST = 1,        // x *= scale_lo[0]
    SCALE_TABLE_NAT = 2,    // x *= scale_lo[k & 4095] * scale_hi[k >> 12],  k = position
    SCALE_TABLE_BREV = 3    // same with k = brev_{log_n}(position)
};

struct NttPass {
    Fp *dst;                    // column 0 of the destination matrix
    const Fp *src;              // column 0 of the source matrix (== dst for in-place passes)
    unsigned long long dst_col_stride, src_col_stride;   // elements between columns
    int n_cols;
    int log_n;                  // transform length of the destination
    int log_block;              // log2 of the sub-problem size this pass works on
    int L;                      // radix-2 stages done by this pass
    int contiguous;             // 1: tile = 4096 consecutive elements (stride-1 sub-problems)
    int expand_log;             // DIT first pass of an LDE: dst position p reads src[p >> e] if low bits zero else 0
    const Fp *tw_local;         // w_4096^i, i < 2048         (direction already applied)
    const Fp *tw_lo;            // w_N^i, i < 4096
    const Fp *tw_hi;            // w_N^(4096 i), i < max(1, N/4096)
    int pre_scale, post_scale;  // NttScaleMode
    const Fp *scale_lo, *scale_hi;
    int canon_out;              // canonicalise on store (last pass)
};

namespace nttk {

struct Smem {
    uint4 *lo, *hi;
    __device__ __forceinline__ static int slot(int t) { return t ^ ((t >> 3) & 7); }
    __device__ __forceinline__ void put(int t, const Fp &v) const {
        const int s = slot(t);
        lo[s] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[s] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
    }
    __device__ __forceinline__ Fp get(int t) const {
        const int s = slot(t);
        const uint4 a = lo[s], b = hi[s];
        Fp v;
        v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
        v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
        return v;
    }
};

__device__ __forceinline__ Fp ldg_fp(const Fp *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
    v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}
// streaming load / store of matrix data (read once, written once per pass)
__device__ __forceinline__ Fp ld_stream(const Fp *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldcs(q), b = __ldcs(q + 1);
    Fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
    v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}
__device__ __forceinline__ void st_stream(Fp *p, const Fp &v) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
    __stcs(q, make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]));
    __stcs(q + 1, make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]));
}

// lo[i] = c * h^i (i < 4096), hi[i] = h^(4096 i):  c * h^k
__device__ __forceinline__ Fp two_level(const Fp *lo, const Fp *hi, unsigned long long k) {
    Fp v = ldg_fp(lo + (k & 4095ull));
    if (k >> 12) v = fp::mul(v, ldg_fp(hi + (k >> 12)));
    return v;
}
// twiddle tables have c = 1, so a multiple of 4096 needs no multiplication
__device__ __forceinline__ Fp two_level_tw(const Fp *lo, const Fp *hi, unsigned long long k) {
    if ((k & 4095ull) == 0) return ldg_fp(hi + (k >> 12));
    return two_level(lo, hi, k);
}

__device__ __forceinline__ Fp apply_scale(const Fp &x, int mode, const Fp *lo, const Fp *hi,
                                          unsigned long long pos, int log_n) {
    if (mode == SCALE_NONE) return x;
    if (mode == SCALE_CONST) return fp::mul(x, ldg_fp(lo));
    unsigned long long k = pos;
    if (mode == SCALE_TABLE_BREV) k = __brevll(pos) >> (64 - log_n);
    return fp::mul(x, two_level(lo, hi, k));
}

// One radix-2 stage on the 4 register pairs that differ in register bit RB.
template <bool DIT, int RB>
__device__ __forceinline__ void stage(Fp (&x)[8], int base, int q, int q0, int L, const Fp *tw_local) {
    const int bl = q + RB - q0;                     // local (transform) bit of this stage
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        constexpr int LOWMASK = (1 << RB) - 1;
        const int i0 = ((j >> RB) << (RB + 1)) | (j & LOWMASK);
        const int i1 = i0 | (1 << RB);
        const unsigned int tau0 = (unsigned int)(base | (i0 << q));
        const unsigned int l0 = (tau0 >> q0) & ((1u << L) - 1u);
        const unsigned int tw_idx = (l0 & ((1u << bl) - 1u)) << (11 - bl);
        // w^0 = 1: no multiplication.  In the round that holds the lowest transform bits the index depends on register bits only
        // (the same for every thread of the warp), so half of the bl = 1 and a quarter of the bl = 2 butterflies really skip it.
        const bool unit = bl == 0 || tw_idx == 0;
        if (DIT) {
            Fp tm;
            if (unit) {
                tm = x[i1];
                if (bl != 0) { fp::cond_sub_4p(tm); fp::cond_sub_2p(tm); }     // (a product would have come back below 2p)
            } else {
                tm = fp::mul(x[i1], ldg_fp(tw_local + tw_idx));
            }
            x[i1] = fp::sub2p(x[i0], tm);
            x[i0] = fp::add_raw(x[i0], tm);
        } else {
            const Fp d = fp::sub4p(x[i0], x[i1]);
            x[i0] = fp::add_fast(x[i0], x[i1]);
            if (unit) { x[i1] = d; fp::cond_sub_4p(x[i1]); fp::cond_sub_2p(x[i1]); }
            else x[i1] = fp::mul(d, ldg_fp(tw_local + tw_idx));
        }
    }
}

}  // namespace nttk

// One pass.  grid = (tiles per column [or tiles over all columns when n < 4096], n_cols or 1).
template <bool DIT>
__global__ void __launch_bounds__(NTT_THREADS, SS_NTT_MIN_CTAS) ntt_pass_kernel(const NttPass P) {
    extern __shared__ uint4 smem_raw[];
    nttk::Smem sm{smem_raw, smem_raw + NTT_TILE};
    const int t = threadIdx.x;
    const int L = P.L;
    const int q0 = P.contiguous ? 0 : (NTT_LOG_TILE - L);          // lowest tile bit that is a transform bit
    const unsigned long long n = 1ull << P.log_n;
    const bool small = P.log_n < NTT_LOG_TILE;                       // several whole columns per tile
    const unsigned long long tile = blockIdx.x;
    const int col_base = small ? (int)(tile << (NTT_LOG_TILE - P.log_n)) : (int)blockIdx.y;

    // tile element tau -> (column, position within column)
    auto locate = [&](int tau, int &col, unsigned long long &pos) {
        if (small) {
            col = col_base + (tau >> P.log_n);
            pos = (unsigned long long)(tau & ((1 << P.log_n) - 1));
        } else if (P.contiguous) {
            col = col_base;
            pos = (tile << NTT_LOG_TILE) + (unsigned long long)tau;
        } else {
            const int lb = P.log_block;
            const unsigned long long blk = tile >> (lb - NTT_LOG_TILE);
            const unsigned long long lo0 = (tile & ((1ull << (lb - NTT_LOG_TILE)) - 1)) << (NTT_LOG_TILE - L);
            const unsigned long long l = (unsigned long long)(tau >> (NTT_LOG_TILE - L));
            const unsigned long long c = (unsigned long long)(tau & ((1 << (NTT_LOG_TILE - L)) - 1));
            col = col_base;
            pos = (blk << lb) + (l << (lb - L)) + lo0 + c;
        }
    };
    // inter-pass twiddle w_B^(lo * brev_L(l)) as a power of w_N (strided passes only)
    auto interpass = [&](int tau, unsigned long long pos) -> Fp {
        const int lb = P.log_block;
        const unsigned long long lo = pos & ((1ull << (lb - L)) - 1);
        const unsigned int l = (unsigned int)(tau >> (NTT_LOG_TILE - L));
        const unsigned long long r = (unsigned long long)(__brev(l) >> (32 - L));
        const unsigned long long e = (lo * r) << (P.log_n - lb);
        return nttk::two_level_tw(P.tw_lo, P.tw_hi, e);
    };

    // ---------------------------------------------------------------- load
    // fully unrolled: 16 independent LDG.128 in flight per thread before the first STS
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int tau = t + NTT_THREADS * i;
        int col; unsigned long long pos;
        locate(tau, col, pos);
        Fp v = fp::zero();
        if (col < P.n_cols) {
            if (P.expand_log) {
                if ((pos & ((1ull << P.expand_log) - 1)) == 0)
                    v = nttk::ld_stream(P.src + (unsigned long long)col * P.src_col_stride + (pos >> P.expand_log));
            } else {
                v = nttk::ld_stream(P.src + (unsigned long long)col * P.src_col_stride + pos);
            }
            if (P.pre_scale != SCALE_NONE)
                v = nttk::apply_scale(v, P.pre_scale, P.scale_lo, P.scale_hi, pos, P.log_n);
            if (DIT && !P.contiguous) v = fp::mul(v, interpass(tau, pos));
        }
        sm.put(tau, v);
    }
    __syncthreads();

    // -------------------------------------------------------------- rounds
    // transform bits are tile bits [q0, q0+L); DIF walks them top-down, DIT bottom-up, 3 per round
    const int n_rounds = (L + 2) / 3;
    for (int r = 0; r < n_rounds; ++r) {
        int g, w;                                    // this round handles tile bits [g, g+w)
        if (DIT) { g = q0 + 3 * r; w = min(3, q0 + L - g); }
        else     { const int top = q0 + L - 3 * r; w = min(3, top - q0); g = top - w; }
        const int q = min(g, NTT_LOG_TILE - 3);      // register bits sit at tile bits [q, q+3)
        const int base = ((t >> q) << (q + 3)) | (t & ((1 << q) - 1));
        Fp x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = sm.get(base | (i << q));
        // register bit rb <-> tile bit q + rb; active iff g <= q + rb < g + w.  rb is a template
        // argument so that x[] is indexed with compile-time constants (stays in registers).
        const int off = g - q;
        if (DIT) {
            if (0 >= off && 0 < off + w) nttk::stage<true, 0>(x, base, q, q0, L, P.tw_local);
            if (1 >= off && 1 < off + w) nttk::stage<true, 1>(x, base, q, q0, L, P.tw_local);
            if (2 >= off && 2 < off + w) nttk::stage<true, 2>(x, base, q, q0, L, P.tw_local);
        } else {
            if (2 >= off && 2 < off + w) nttk::stage<false, 2>(x, base, q, q0, L, P.tw_local);
            if (1 >= off && 1 < off + w) nttk::stage<false, 1>(x, base, q, q0, L, P.tw_local);
            if (0 >= off && 0 < off + w) nttk::stage<false, 0>(x, base, q, q0, L, P.tw_local);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (DIT) { fp::cond_sub_4p(x[i]); fp::cond_sub_2p(x[i]); }
            sm.put(base | (i << q), x[i]);
        }
        __syncthreads();
    }

    // --------------------------------------------------------------- store
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
        const int tau = t + NTT_THREADS * i;
        int col; unsigned long long pos;
        locate(tau, col, pos);
        if (col >= P.n_cols) continue;
        Fp v = sm.get(tau);
        if (!DIT && !P.contiguous) v = fp::mul(v, interpass(tau, pos));
        v = nttk::apply_scale(v, P.post_scale, P.scale_lo, P.scale_hi, pos, P.log_n);
        if (P.canon_out) v = fp::canon(v);
        nttk::st_stream(P.dst + (unsigned long long)col * P.dst_col_stride + pos, v);
    }
}

}  // namespace ss
