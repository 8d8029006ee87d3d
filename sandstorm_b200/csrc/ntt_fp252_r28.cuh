// Carry-free variant of the NTT pass kernel (see ntt_fp252.cuh for the algorithm and fp28.cuh for the
// arithmetic): tile elements live in shared memory as 9 x 28-bit limbs (36 bytes: two uint4 planes + one
// word plane), butterflies use limb-wise adds and 81 + 30 plain IMAD.WIDE per twiddle multiplication,
// twiddle / scale tables hold constants pre-scaled by 2^280 as 12-word entries.
#pragma once
#include "fp28.cuh"
#include "ntt_fp252.cuh"

namespace ss {

constexpr int NTT28_SMEM_BYTES = NTT_TILE * 36;
constexpr int F28_WORDS = 12;                 // table entry stride (48 bytes, 16-byte aligned)

struct NttPass28 {
    NttPass base;                             // geometry, pointers to data, modes (table pointers unused)
    const uint32_t *tw_local, *tw_lo, *tw_hi, *scale_lo, *scale_hi;   // F28 tables
    f28::MulK K;
};

namespace nttk28 {

struct Smem28 {
    uint4 *lo, *hi;
    uint32_t *top;
    __device__ __forceinline__ static int slot(int t) { return t ^ ((t >> 3) & 7); }
    __device__ __forceinline__ void put(int t, const F28 &v) const {
        const int s = slot(t);
        lo[s] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
        hi[s] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
        top[s] = v.l[8];
    }
    __device__ __forceinline__ F28 get(int t) const {
        const int s = slot(t);
        const uint4 a = lo[s], b = hi[s];
        F28 v;
        v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
        v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
        v.l[8] = top[s];
        return v;
    }
};

__device__ __forceinline__ F28 ldg_c(const uint32_t *table, unsigned long long idx) {
    const uint32_t *p = table + idx * F28_WORDS;
    const uint4 a = __ldg(reinterpret_cast<const uint4 *>(p)), b = __ldg(reinterpret_cast<const uint4 *>(p) + 1);
    F28 v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w;
    v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    v.l[8] = __ldg(p + 8);
    return v;
}

// lo[i] = c * h^i (i < 4096), hi[i] = h^(4096 i), both in table form: c * h^k
__device__ __forceinline__ F28 two_level(const uint32_t *lo, const uint32_t *hi, unsigned long long k, const f28::MulK &K) {
    F28 v = ldg_c(lo, k & 4095ull);
    if (k >> 12) v = f28::mulc(v, ldg_c(hi, k >> 12), K);
    return v;
}
__device__ __forceinline__ F28 two_level_tw(const uint32_t *lo, const uint32_t *hi, unsigned long long k, const f28::MulK &K) {
    if ((k & 4095ull) == 0) return ldg_c(hi, k >> 12);
    return two_level(lo, hi, k, K);
}

__device__ __forceinline__ F28 apply_scale(const F28 &x, int mode, const uint32_t *lo, const uint32_t *hi,
                                           unsigned long long pos, int log_n, const f28::MulK &K) {
    if (mode == SCALE_NONE) return x;
    if (mode == SCALE_CONST) return f28::mulc(x, ldg_c(lo, 0), K);
    unsigned long long k = pos;
    if (mode == SCALE_TABLE_BREV) k = __brevll(pos) >> (64 - log_n);
    return f28::mulc(x, two_level(lo, hi, k, K), K);
}

// One radix-2 stage on the 4 register pairs that differ in register bit RB (bounds: fp28.cuh).
template <bool DIT, int RB>
__device__ __forceinline__ void stage(F28 (&x)[8], int base, int q, int q0, int L, const uint32_t *tw_local, const f28::MulK &K) {
    const int bl = q + RB - q0;                     // local (transform) bit of this stage
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        constexpr int LOWMASK = (1 << RB) - 1;
        const int i0 = ((j >> RB) << (RB + 1)) | (j & LOWMASK);
        const int i1 = i0 | (1 << RB);
        const unsigned int tau0 = (unsigned int)(base | (i0 << q));
        const unsigned int l0 = (tau0 >> q0) & ((1u << L) - 1u);
        const unsigned int tw_idx = (l0 & ((1u << bl) - 1u)) << (11 - bl);
        if (DIT) {
            const F28 tm = (bl == 0) ? x[i1] : f28::mulc(x[i1], ldg_c(tw_local, tw_idx), K);
            x[i1] = f28::sub_dit(x[i0], tm);
            x[i0] = f28::add(x[i0], tm);
        } else {
            const F28 d = f28::sub_dif(x[i0], x[i1]);
            x[i0] = f28::add(x[i0], x[i1]);
            x[i1] = (bl == 0) ? d : f28::mulc(d, ldg_c(tw_local, tw_idx), K);
        }
    }
}

}  // namespace nttk28

template <bool DIT>
__global__ void __launch_bounds__(NTT_THREADS, SS_NTT_MIN_CTAS) ntt_pass_kernel_r28(const NttPass28 Q) {
    const NttPass &P = Q.base;
    const f28::MulK K = Q.K;
    extern __shared__ uint4 smem_raw[];
    nttk28::Smem28 sm{smem_raw, smem_raw + NTT_TILE, reinterpret_cast<uint32_t *>(smem_raw + 2 * NTT_TILE)};
    const int t = threadIdx.x;
    const int L = P.L;
    const int q0 = P.contiguous ? 0 : (NTT_LOG_TILE - L);
    const bool small = P.log_n < NTT_LOG_TILE;
    const unsigned long long tile = blockIdx.x;
    const int col_base = small ? (int)(tile << (NTT_LOG_TILE - P.log_n)) : (int)blockIdx.y;

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
    auto interpass = [&](int tau, unsigned long long pos) -> F28 {
        const int lb = P.log_block;
        const unsigned long long lo = pos & ((1ull << (lb - L)) - 1);
        const unsigned int l = (unsigned int)(tau >> (NTT_LOG_TILE - L));
        const unsigned long long r = (unsigned long long)(__brev(l) >> (32 - L));
        const unsigned long long e = (lo * r) << (P.log_n - lb);
        return nttk28::two_level_tw(Q.tw_lo, Q.tw_hi, e, K);
    };

    // ---------------------------------------------------------------- load
#pragma unroll 2
    for (int i = 0; i < 8; ++i) {
        const int tau = t + NTT_THREADS * i;
        int col; unsigned long long pos;
        locate(tau, col, pos);
        Fp v = fp::zero();
        bool live = false;
        if (col < P.n_cols) {
            if (P.expand_log) {
                if ((pos & ((1ull << P.expand_log) - 1)) == 0) {
                    v = nttk::ld_stream(P.src + (unsigned long long)col * P.src_col_stride + (pos >> P.expand_log));
                    live = true;
                }
            } else {
                v = nttk::ld_stream(P.src + (unsigned long long)col * P.src_col_stride + pos);
                live = true;
            }
        }
        F28 x = f28::from_fp(v);
        if (live) {
            if (P.pre_scale != SCALE_NONE) x = nttk28::apply_scale(x, P.pre_scale, Q.scale_lo, Q.scale_hi, pos, P.log_n, K);
            if (DIT && !P.contiguous) x = f28::mulc(x, interpass(tau, pos), K);
        }
        sm.put(tau, x);
    }
    __syncthreads();

    // -------------------------------------------------------------- rounds
    const int n_rounds = (L + 2) / 3;
    for (int r = 0; r < n_rounds; ++r) {
        int g, w;
        if (DIT) { g = q0 + 3 * r; w = min(3, q0 + L - g); }
        else     { const int top = q0 + L - 3 * r; w = min(3, top - q0); g = top - w; }
        const int q = min(g, NTT_LOG_TILE - 3);
        const int base = ((t >> q) << (q + 3)) | (t & ((1 << q) - 1));
        F28 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = sm.get(base | (i << q));
        const int off = g - q;
        if (DIT) {
            if (0 >= off && 0 < off + w) nttk28::stage<true, 0>(x, base, q, q0, L, Q.tw_local, K);
            if (1 >= off && 1 < off + w) nttk28::stage<true, 1>(x, base, q, q0, L, Q.tw_local, K);
            if (2 >= off && 2 < off + w) nttk28::stage<true, 2>(x, base, q, q0, L, Q.tw_local, K);
        } else {
            if (2 >= off && 2 < off + w) nttk28::stage<false, 2>(x, base, q, q0, L, Q.tw_local, K);
            if (1 >= off && 1 < off + w) nttk28::stage<false, 1>(x, base, q, q0, L, Q.tw_local, K);
            if (0 >= off && 0 < off + w) nttk28::stage<false, 0>(x, base, q, q0, L, Q.tw_local, K);
        }
        // round end: lazy values (limb-wise sums, a - b + Kp) back to normalised limbs, value < 2^252 + 9 * 2^224.
        // DIF: the outputs of the last stage with its register bit set are products (already normalised),
        // unless that stage had the trivial twiddle (local bit 0), where they are raw differences.
        const int rb_last = off;                      // lowest active register bit = last DIF stage
        const bool last_trivial = (g == q0);          // local bit 0 handled in this round
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool lazy = DIT || last_trivial || (((i >> rb_last) & 1) == 0);
            if (lazy) x[i] = f28::weak_reduce(x[i]);
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
        F28 x = sm.get(tau);
        if (!DIT && !P.contiguous) x = f28::mulc(x, interpass(tau, pos), K);
        x = nttk28::apply_scale(x, P.post_scale, Q.scale_lo, Q.scale_hi, pos, P.log_n, K);
        const Fp v = P.canon_out ? f28::to_canonical_fp(x) : f28::to_fp(x);
        nttk::st_stream(P.dst + (unsigned long long)col * P.dst_col_stride + pos, v);
    }
}

}  // namespace ss
