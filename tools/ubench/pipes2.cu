// Throughput of carry-chained wide multiply-adds (the form fp::mul uses) vs carry-free IMAD.WIDE.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a0 = seed + threadIdx.x, a1 = seed * 3 + 1, a2 = seed ^ 0x5555, a3 = seed + 77, b = seed * 5 + threadIdx.x;
    uint32_t r[4][9];
    uint64_t w[16];
    for (int c = 0; c < 4; ++c) for (int i = 0; i < 9; ++i) r[c][i] = a0 + c * 9 + i;
    for (int i = 0; i < 16; ++i) w[i] = a1 + i;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
        if (OP == 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                    "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                    "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                    "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                    : "+r"(r[c][0]), "+r"(r[c][1]), "+r"(r[c][2]), "+r"(r[c][3]), "+r"(r[c][4]), "+r"(r[c][5]), "+r"(r[c][6]), "+r"(r[c][7]), "+r"(r[c][8])
                    : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
        } else if (OP == 1) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a0), "r"(b));
        } else if (OP == 2) {   // carry-out only (first of a chain) x4 + plain
#pragma unroll
            for (int c = 0; c < 4; ++c)
                asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\taddc.cc.u32 %2, %2, 0;\n\taddc.u32 %3, %3, 0;"
                    : "+r"(r[c][0]), "+r"(r[c][1]), "+r"(r[c][2]), "+r"(r[c][3]) : "r"(a0), "r"(b));
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int c = 0; c < 4; ++c) for (int i = 0; i < 9; ++i) acc += r[c][i];
    for (int i = 0; i < 16; ++i) acc += (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char *name, int wide_per_iter) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {256, 512, 1024}) {
        k<OP><<<148, threads>>>(out, 12345, cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        double n = (double)ITERS * wide_per_iter * (threads / 32);
        printf("%-40s threads=%4d cycles=%9.0f  wide-mads/cycle/SM=%6.3f\n", name, threads, c, n / c);
    }
}
int main() {
    run<0>("4 chains x (4 wide w/ carry + addc)", 16);
    run<1>("16 independent IMAD.WIDE (64-bit acc)", 16);
    run<2>("4 x (1 wide carry-out + 2 addc)", 4);
    return 0;
}
