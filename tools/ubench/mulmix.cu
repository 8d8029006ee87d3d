// Go / no-go microbenchmark for VERDICT r1 item 4(c): can Fp252 multiplications on the FP64 pipe (DFMA split products, 52-bit
// limbs) run BESIDE the carry-chained IMAD.WIDE multiplier and raise the per-SM multiplication rate?
//   mul_f64: a, b -> a*b/2^256 mod p through 25 limb products, each hi = fma_rz(a_i, b_j, 2^104), lo = fma_rz(a_i, b_j, 2^104 + 2^52 - hi)
//            (exact 104-bit product split at bit 52), integer accumulation of the bit patterns, re-alignment to 16 x u32 and the
//            same sparse Montgomery reduction as fp::mul.  Checked against fp::mul on random operands first.
//   modes:   all warps integer | all warps FP64 | warps alternate (warp-specialised: the two pipes in parallel)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mulmix mulmix.cu ; run on the GPU box.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../sandstorm_b200/csrc/fp252.cuh"
using namespace ss;

__device__ __forceinline__ void to_limbs52(const Fp &a, double (&d)[5]) {
    const unsigned long long w0 = a.l[0] | ((unsigned long long)a.l[1] << 32), w1 = a.l[2] | ((unsigned long long)a.l[3] << 32);
    const unsigned long long w2 = a.l[4] | ((unsigned long long)a.l[5] << 32), w3 = a.l[6] | ((unsigned long long)a.l[7] << 32);
    const unsigned long long M = (1ull << 52) - 1, E = 0x4330000000000000ull;
    const unsigned long long L[5] = {w0 & M, ((w0 >> 52) | (w1 << 12)) & M, ((w1 >> 40) | (w2 << 24)) & M, ((w2 >> 28) | (w3 << 36)) & M, w3 >> 16};
#pragma unroll
    for (int i = 0; i < 5; ++i) d[i] = __longlong_as_double((long long)(L[i] | E)) - 4503599627370496.0;
}

__device__ __forceinline__ Fp mul_f64(const Fp &a, const Fp &b) {
    double x[5], y[5];
    to_limbs52(a, x);
    to_limbs52(b, y);
    const double C1 = 20282409603651670423947251286016.0;            // 2^104
    const double C2 = 20282409603651674927546878656512.0;            // 2^104 + 2^52
    unsigned long long col[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        // bit patterns carry the exponent fields: 0x433 (lo, value 2^52 + lo) and 0x467 (hi, value 2^104 + hi 2^52)
        const int n_lo = k < 5 ? k + 1 : 9 - k, n_hi = k == 0 ? 0 : (k - 1 < 5 ? k : 10 - k);
        col[k] = 0ull - (unsigned long long)n_lo * 0x4330000000000000ull - (unsigned long long)n_hi * 0x4670000000000000ull;
    }
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            const double hi = __fma_rz(x[i], y[j], C1);
            const double lo = __fma_rz(x[i], y[j], C2 - hi);
            col[i + j] += (unsigned long long)__double_as_longlong(lo);
            col[i + j + 1] += (unsigned long long)__double_as_longlong(hi);
        }
    // normalise to 52-bit limbs, repack into 64-bit words, add p 2^256 (keeps the reduction's result positive, as fp::mul does)
    const unsigned long long M = (1ull << 52) - 1;
    unsigned long long n[10], carry = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const unsigned long long s = col[k] + carry;
        n[k] = k < 9 ? (s & M) : s;
        carry = s >> 52;
    }
    unsigned long long W[8];
    W[0] = n[0] | (n[1] << 52);
    W[1] = (n[1] >> 12) | (n[2] << 40);
    W[2] = (n[2] >> 24) | (n[3] << 28);
    W[3] = (n[3] >> 36) | (n[4] << 16);
    W[4] = (n[4] >> 48) | (n[5] << 4) | (n[6] << 56);
    W[5] = (n[6] >> 8) | (n[7] << 44);
    W[6] = (n[7] >> 20) | (n[8] << 32);
    W[7] = (n[8] >> 32) | (n[9] << 20);
    W[4] += 1ull;
    unsigned long long c = W[4] == 0ull;
    W[5] += c; c = c & (W[5] == 0ull);
    W[6] += c; c = c & (W[6] == 0ull);
    W[7] += c + (((unsigned long long)SS_P7 << 32) | SS_P6);
    uint32_t T16[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { T16[2 * i] = (uint32_t)W[i]; T16[2 * i + 1] = (uint32_t)(W[i] >> 32); }
    return fp::mont_reduce(T16);
}

#define ITERS 256
template <int MODE>      // 0: integer, 1: fp64, 2: alternate warps
__global__ void __launch_bounds__(512) bench(Fp *out, long long *cycles) {
    Fp b, w;
    for (int i = 0; i < 8; ++i) { b.l[i] = blockIdx.x * 13 + threadIdx.x * 7 + i * 3; w.l[i] = 0x1234567u * (i + 1); }
    b.l[7] &= 0x07ffffff; w.l[7] &= 0x07ffffff;
    const bool use_f = MODE == 1 || (MODE == 2 && ((threadIdx.x >> 5) & 1));
    __syncthreads();
    const long long t0 = clock64();
    if (use_f) { for (int it = 0; it < ITERS; ++it) b = mul_f64(b, w); }
    else       { for (int it = 0; it < ITERS; ++it) b = fp::mul(b, w); }
    __syncthreads();
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = b;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void check(const Fp *a, const Fp *b, int n, int *bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fp r0 = fp::canon(fp::mul(a[i], b[i])), r1 = fp::canon(mul_f64(a[i], b[i]));
    for (int k = 0; k < 8; ++k) if (r0.l[k] != r1.l[k]) { atomicAdd(bad, 1); break; }
}

template <int MODE>
void run(const char *name, Fp *out, long long *cyc) {
    for (int threads : {128, 256, 384, 512}) {
        bench<MODE><<<148, threads>>>(out, cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s threads=%d failed: %s\n", name, threads, cudaGetErrorString(cudaGetLastError())); continue; }
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        const double warp_muls = (double)ITERS * (threads / 32);
        printf("%-22s warps/SM=%2d  cycles per warp-mul per SMSP = %7.1f   (muls/clk/SM = %.3f)\n", name, threads / 32, c / (warp_muls / 4), warp_muls * 32 / c);
    }
}

int main() {
    const int n = 1 << 16;
    Fp *a, *b; int *bad;
    cudaMallocManaged(&a, n * sizeof(Fp)); cudaMallocManaged(&b, n * sizeof(Fp)); cudaMallocManaged(&bad, 4);
    srand(1);
    for (int i = 0; i < n; ++i) for (int k = 0; k < 8; ++k) { a[i].l[k] = (uint32_t)rand() * 2654435761u + rand(); b[i].l[k] = (uint32_t)rand() * 40503u + rand(); }
    for (int i = 0; i < n; ++i) { a[i].l[7] &= 0x7fffffff; b[i].l[7] &= 0x7fffffff; }      // < 2^255: the multiplier's input bound
    for (int k = 0; k < 8; ++k) { a[0].l[k] = 0xffffffffu; b[0].l[k] = 0xffffffffu; a[1].l[k] = 0; b[2].l[k] = k == 0; }
    a[0].l[7] = b[0].l[7] = 0x7fffffffu;
    *bad = 0;
    check<<<n / 256, 256>>>(a, b, n, bad);
    cudaDeviceSynchronize();
    printf("mul_f64 vs fp::mul on %d operand pairs: %d mismatches\n", n, *bad);
    Fp *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * sizeof(Fp)); cudaMalloc(&cyc, 148 * 8);
    run<0>("integer (fp::mul)", out, cyc);
    run<1>("fp64 (mul_f64)", out, cyc);
    run<2>("alternating warps", out, cyc);
    return *bad != 0;
}
