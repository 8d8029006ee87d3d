// Do carry-chained wide MADs and IADD3.X chains overlap (different pipes) or serialise?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int NW, int NA>   // NW chains of 4 carry-wide mads, NA chains of 8 add.cc/addc per iteration
__global__ void k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a0 = seed + threadIdx.x, a1 = seed * 3 + 1, a2 = seed ^ 0x5555, a3 = seed + 77, b = seed * 5 + threadIdx.x;
    uint32_t r[4][9], s[4][8];
    for (int c = 0; c < 4; ++c) { for (int i = 0; i < 9; ++i) r[c][i] = a0 + c * 9 + i; for (int i = 0; i < 8; ++i) s[c][i] = a1 * (c + 2) + i; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < NW; ++c)
            asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\tmadc.hi.cc.u32 %1, %9, %13, %1;\n\t"
                "madc.lo.cc.u32 %2, %10, %13, %2;\n\tmadc.hi.cc.u32 %3, %10, %13, %3;\n\t"
                "madc.lo.cc.u32 %4, %11, %13, %4;\n\tmadc.hi.cc.u32 %5, %11, %13, %5;\n\t"
                "madc.lo.cc.u32 %6, %12, %13, %6;\n\tmadc.hi.cc.u32 %7, %12, %13, %7;\n\taddc.u32 %8, %8, 0;"
                : "+r"(r[c][0]), "+r"(r[c][1]), "+r"(r[c][2]), "+r"(r[c][3]), "+r"(r[c][4]), "+r"(r[c][5]), "+r"(r[c][6]), "+r"(r[c][7]), "+r"(r[c][8])
                : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#pragma unroll
        for (int c = 0; c < NA; ++c)
            asm volatile("add.cc.u32 %0, %0, %8;\n\taddc.cc.u32 %1, %1, %9;\n\taddc.cc.u32 %2, %2, %8;\n\taddc.cc.u32 %3, %3, %9;\n\t"
                "addc.cc.u32 %4, %4, %8;\n\taddc.cc.u32 %5, %5, %9;\n\taddc.cc.u32 %6, %6, %8;\n\taddc.u32 %7, %7, %9;"
                : "+r"(s[c][0]), "+r"(s[c][1]), "+r"(s[c][2]), "+r"(s[c][3]), "+r"(s[c][4]), "+r"(s[c][5]), "+r"(s[c][6]), "+r"(s[c][7]) : "r"(a2), "r"(a3));
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int c = 0; c < 4; ++c) { for (int i = 0; i < 9; ++i) acc += r[c][i]; for (int i = 0; i < 8; ++i) acc += s[c][i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int NW, int NA>
void run() {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<NW, NA><<<148, 512>>>(out, 12345, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    printf("per iteration per warp per SMSP: %d wide-chains(4) + %d add-chains(8): %7.1f cycles\n", NW, NA, c / ITERS / 4.0);
}
int main() { run<4, 0>(); run<0, 4>(); run<4, 4>(); run<4, 2>(); run<2, 4>(); return 0; }
