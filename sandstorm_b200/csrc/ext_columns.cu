// `Trace::build_extension_columns` on the device: the running-product columns of the Cairo
// layouts — memory, range-check and diluted-check permutation arguments and the diluted-check aggregation — which
// the reference computes with sequential, single-threaded loops plus batch inversions
// (layouts/src/recursive/trace.rs:699-814, "TODO: multithread" :700; layouts/src/starknet/trace.rs:997-1100).
//
//   ss_perm_product      out[i] = prod_{j<=i} (z - (alpha * num_v[j] + num_a[j])) / (z - (alpha * den_v[j] + den_a[j]))
//                        (value pointers NULL: z - a[j])                                       trace.rs:706-755
//   ss_diluted_aggregate out[0] = 1,  out[i] = out[i-1] * (1 + z u_i) + alpha u_i^2,  u_i = d[i] - d[i-1]   trace.rs:787-806
//
// Both are prefix scans under an associative operation on pairs of field elements:
//   permutation : (n, d)  with (n1, d1) . (n2, d2) = (n1 n2, d1 d2);   out = n / d  (one inversion per block, Montgomery's trick)
//   aggregation : affine maps x -> A x + B with composition;           out = A + B  (the map applied to the initial 1)
// done in three launches: per-block totals, an exclusive scan of the block totals by one block, per-block rescan
// with the carried-in prefix.  Elements are read with a caller-given stride (the columns interleave several virtual
// columns: memory pairs at stride 2, range-check cells at stride 4, ...) and written with their own stride.
//
// Algorithmic bytes per element: 2-4 inputs + 1 output of 32 B; about 4 multiplications per input element per pass.
#include "ctx.h"
#include "pedersen.cuh"   // ec::inv_chain
#include <cstring>

using namespace ss;

namespace {

constexpr int SC_THREADS = 128;
constexpr int SC_PER_THREAD = 4;
constexpr int SC_CHUNK = SC_THREADS * SC_PER_THREAD;

struct Pair { Fp a, b; };

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

struct PermArgs {
    const Fp *num_a, *num_v, *den_a, *den_v;
    unsigned long long stride, count;
    Fp z, alpha;
    Fp *out;
    unsigned long long out_stride;
};
struct AggArgs {
    const Fp *d;
    unsigned long long stride, count;
    Fp z, alpha;
    Fp *out;
    unsigned long long out_stride;
};

// ---- the two monoids ----------------------------------------------------------------------------------------------
struct PermOp {
    using Args = PermArgs;
    __device__ static Pair identity() { return {fp::one(), fp::one()}; }
    // `first` happens before `then`
    __device__ static Pair combine(const Pair &first, const Pair &then) { return {fp::mul(first.a, then.a), fp::mul(first.b, then.b)}; }
    __device__ static Pair element(const Args &A, unsigned long long j) {
        if (j >= A.count) return identity();
        Fp n = ld_fp(A.num_a + j * A.stride), d = ld_fp(A.den_a + j * A.stride);
        if (A.num_v) {
            n = fp::add(n, fp::mul(A.alpha, ld_fp(A.num_v + j * A.stride)));
            d = fp::add(d, fp::mul(A.alpha, ld_fp(A.den_v + j * A.stride)));
        }
        return {fp::sub(A.z, n), fp::sub(A.z, d)};
    }
};
struct AggOp {
    using Args = AggArgs;
    __device__ static Pair identity() { return {fp::one(), fp::zero()}; }
    // maps compose: then(first(x)) = then.a * (first.a x + first.b) + then.b
    __device__ static Pair combine(const Pair &first, const Pair &then) {
        return {fp::mul(first.a, then.a), fp::add(fp::mul(then.a, first.b), then.b)};
    }
    __device__ static Pair element(const Args &A, unsigned long long j) {
        if (j >= A.count || j == 0) return identity();                  // out[0] = 1: the first map is the identity
        const Fp u = fp::sub(ld_fp(A.d + j * A.stride), ld_fp(A.d + (j - 1) * A.stride));
        return {fp::add(fp::one(), fp::mul(A.z, u)), fp::mul(A.alpha, fp::mul(u, u))};
    }
};

// inclusive scan of one value per thread across the block (Hillis-Steele in shared memory); returns the inclusive
// prefix of this thread and leaves the block total in sm[SC_THREADS - 1]
template <class Op>
__device__ __forceinline__ Pair block_scan(Pair v, Pair *sm) {
    const int tid = threadIdx.x;
    sm[tid] = v;
    __syncthreads();
    for (int off = 1; off < SC_THREADS; off <<= 1) {
        Pair left;
        const bool has = tid >= off;
        if (has) left = sm[tid - off];
        __syncthreads();
        if (has) { v = Op::combine(left, v); sm[tid] = v; }
        __syncthreads();
    }
    return v;
}

// pass 1: total of every block's SC_CHUNK elements
template <class Op>
__global__ void __launch_bounds__(SC_THREADS) scan_totals_kernel(const typename Op::Args A, Pair *totals) {
    __shared__ Pair sm[SC_THREADS];
    const unsigned long long base = (unsigned long long)blockIdx.x * SC_CHUNK + (unsigned long long)threadIdx.x * SC_PER_THREAD;
    Pair acc = Op::element(A, base);
#pragma unroll 1
    for (int k = 1; k < SC_PER_THREAD; ++k) acc = Op::combine(acc, Op::element(A, base + k));
    block_scan<Op>(acc, sm);
    if (threadIdx.x == 0) {
        Pair t = sm[SC_THREADS - 1];
        t.a = fp::canon(t.a); t.b = fp::canon(t.b);
        totals[blockIdx.x] = t;
    }
}

// pass 2 (one block): exclusive scan of the block totals, in place
template <class Op>
__global__ void __launch_bounds__(SC_THREADS) scan_blocks_kernel(Pair *totals, unsigned int n_blocks) {
    __shared__ Pair sm[SC_THREADS];
    const unsigned int per = (n_blocks + SC_THREADS - 1) / SC_THREADS;
    const unsigned int lo = threadIdx.x * per, hi = min(lo + per, n_blocks);
    Pair acc = Op::identity();
    for (unsigned int k = lo; k < hi; ++k) acc = Op::combine(acc, totals[k]);
    const Pair incl = block_scan<Op>(acc, sm);
    // exclusive prefix of this thread = inclusive prefix of the previous thread
    Pair run = threadIdx.x ? sm[threadIdx.x - 1] : Op::identity();
    (void)incl;
    for (unsigned int k = lo; k < hi; ++k) {
        const Pair t = totals[k];
        Pair w = run;
        w.a = fp::canon(w.a); w.b = fp::canon(w.b);
        totals[k] = w;
        run = Op::combine(run, t);
    }
}

// Montgomery's trick across the block: every thread hands in the product of its denominators and gets its inverse
__device__ __forceinline__ Fp block_inverse(const Fp &v, Fp *sm) {
    const int tid = threadIdx.x;
    sm[SC_THREADS + tid] = v;
    __syncthreads();
    for (int w = SC_THREADS / 2; w >= 1; w >>= 1) {
        if (tid < w) sm[w + tid] = fp::mul(sm[2 * (w + tid)], sm[2 * (w + tid) + 1]);
        __syncthreads();
    }
    if (tid == 0) sm[1] = ec::inv_chain(sm[1]);
    __syncthreads();
    for (int w = 1; w < SC_THREADS; w <<= 1) {
        if (tid < w) {
            const int k = w + tid;
            const Fp ik = sm[k], l = sm[2 * k], r = sm[2 * k + 1];
            sm[2 * k] = fp::mul(ik, r);
            sm[2 * k + 1] = fp::mul(ik, l);
        }
        __syncthreads();
    }
    return sm[SC_THREADS + tid];
}

// pass 3, permutation: out = running numerator / running denominator
__global__ void __launch_bounds__(SC_THREADS) perm_finish_kernel(const PermArgs A, const Pair *prefix) {
    __shared__ Pair sm[SC_THREADS];
    __shared__ Fp inv_sm[2 * SC_THREADS];
    const unsigned long long base = (unsigned long long)blockIdx.x * SC_CHUNK + (unsigned long long)threadIdx.x * SC_PER_THREAD;
    Pair acc = PermOp::identity();
#pragma unroll
    for (int k = 0; k < SC_PER_THREAD; ++k) acc = PermOp::combine(acc, PermOp::element(A, base + k));
    block_scan<PermOp>(acc, sm);
    Pair run = PermOp::combine(prefix[blockIdx.x], threadIdx.x ? sm[threadIdx.x - 1] : PermOp::identity());
    // running pairs of this thread; denominators inverted together: pre[k] = d_0 .. d_{k-1}
    Fp num[SC_PER_THREAD], den[SC_PER_THREAD], pre[SC_PER_THREAD];
    Fp dprod = fp::one();
#pragma unroll
    for (int k = 0; k < SC_PER_THREAD; ++k) {
        run = PermOp::combine(run, PermOp::element(A, base + k));      // (recomputed: cheaper than keeping 2 x 8 registers alive)
        num[k] = run.a; den[k] = run.b;
        pre[k] = dprod;
        dprod = fp::mul(dprod, run.b);
    }
    Fp inv = block_inverse(dprod, inv_sm);
#pragma unroll
    for (int k = SC_PER_THREAD - 1; k >= 0; --k) {
        const Fp dinv = fp::mul(inv, pre[k]);
        inv = fp::mul(inv, den[k]);
        if (base + k < A.count) st_fp(A.out + (base + k) * A.out_stride, fp::canon(fp::mul(num[k], dinv)));
    }
}

// pass 3, aggregation: out = A + B (the composed map applied to the initial value 1)
__global__ void __launch_bounds__(SC_THREADS) agg_finish_kernel(const AggArgs A, const Pair *prefix) {
    __shared__ Pair sm[SC_THREADS];
    const unsigned long long base = (unsigned long long)blockIdx.x * SC_CHUNK + (unsigned long long)threadIdx.x * SC_PER_THREAD;
    Pair acc = AggOp::identity();
#pragma unroll
    for (int k = 0; k < SC_PER_THREAD; ++k) acc = AggOp::combine(acc, AggOp::element(A, base + k));
    block_scan<AggOp>(acc, sm);
    Pair run = AggOp::combine(prefix[blockIdx.x], threadIdx.x ? sm[threadIdx.x - 1] : AggOp::identity());
#pragma unroll
    for (int k = 0; k < SC_PER_THREAD; ++k) {
        run = AggOp::combine(run, AggOp::element(A, base + k));
        if (base + k < A.count) st_fp(A.out + (base + k) * A.out_stride, fp::canon(fp::add(run.a, run.b)));
    }
}

Fp load_host(const void *p) { Fp v; memcpy(v.l, p, 32); return fp::canon(v); }

template <class Op, class Finish>
ss_status run_scan(ss_ctx *ctx, const typename Op::Args &A, Finish finish, cudaStream_t st) {
    const unsigned int n_blocks = (unsigned int)((A.count + SC_CHUNK - 1) / SC_CHUNK);
    Pair *totals = nullptr;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&totals), (size_t)n_blocks * sizeof(Pair)));
    scan_totals_kernel<Op><<<n_blocks, SC_THREADS, 0, st>>>(A, totals);
    scan_blocks_kernel<Op><<<1, SC_THREADS, 0, st>>>(totals, n_blocks);
    finish<<<n_blocks, SC_THREADS, 0, st>>>(A, totals);
    ctx->launches += 3;
    const cudaError_t e = cudaGetLastError();
    dev_free(ctx, totals);
    SS_CUDA_CHECK(ctx, e);
    return SS_OK;
}

}  // namespace

extern "C" {

ss_status ss_perm_product(ss_ctx *ctx, ss_field field, const void *d_num_a, const void *d_num_v, const void *d_den_a,
                          const void *d_den_v, uint64_t stride, uint64_t count, const void *h_z, const void *h_alpha,
                          void *d_out, uint64_t out_stride, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_perm_product: field %d not built", (int)field);
    if (!d_num_a || !d_den_a || !d_out || !h_z || stride == 0 || out_stride == 0 || count >= (1ull << 40) ||
        ((d_num_v != nullptr) != (d_den_v != nullptr)) || (d_num_v && !h_alpha))
        return fail(ctx, SS_ERR_INVALID, "ss_perm_product: bad arguments");
    if (count == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    PermArgs A;
    A.num_a = static_cast<const Fp *>(d_num_a); A.num_v = static_cast<const Fp *>(d_num_v);
    A.den_a = static_cast<const Fp *>(d_den_a); A.den_v = static_cast<const Fp *>(d_den_v);
    A.stride = stride; A.count = count;
    A.z = load_host(h_z);
    A.alpha = h_alpha ? load_host(h_alpha) : fp::zero();
    A.out = static_cast<Fp *>(d_out); A.out_stride = out_stride;
    return run_scan<PermOp>(ctx, A, perm_finish_kernel, pick_stream(ctx, stream));
}

ss_status ss_diluted_aggregate(ss_ctx *ctx, ss_field field, const void *d_ordered, uint64_t stride, uint64_t count,
                               const void *h_z, const void *h_alpha, void *d_out, uint64_t out_stride, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (field != SS_FIELD_FP252) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_diluted_aggregate: field %d not built", (int)field);
    if (!d_ordered || !d_out || !h_z || !h_alpha || stride == 0 || out_stride == 0 || count >= (1ull << 40))
        return fail(ctx, SS_ERR_INVALID, "ss_diluted_aggregate: bad arguments");
    if (count == 0) return SS_OK;
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    AggArgs A;
    A.d = static_cast<const Fp *>(d_ordered);
    A.stride = stride; A.count = count;
    A.z = load_host(h_z); A.alpha = load_host(h_alpha);
    A.out = static_cast<Fp *>(d_out); A.out_stride = out_stride;
    return run_scan<AggOp>(ctx, A, agg_finish_kernel, pick_stream(ctx, stream));
}

}  // extern "C"
