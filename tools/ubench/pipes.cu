// Measures per-SM instruction throughput of the integer/FP64 instructions the Fp252 multiplier can be
// built from (IMAD, IMAD.WIDE.U32, carry variants, IADD3, DFMA) on sm_100a.  One block per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define CHAINS 8
template <int OP>
__global__ void k(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
    uint32_t r[CHAINS * 2];
    uint64_t w[CHAINS];
    double d[CHAINS];
    for (int i = 0; i < CHAINS; ++i) { r[2 * i] = a + i; r[2 * i + 1] = b + i; w[i] = a * 7 + i; d[i] = 1.0 + i; }
    double da = 1.0000001, db = 0.5;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[2 * i]) : "r"(a), "r"(b));
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a), "r"(b));
            if (OP == 2) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(r[2 * i]), "+r"(r[2 * i + 1]) : "r"(a), "r"(b));
            if (OP == 3) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(r[2 * i]) : "r"(a), "r"(b));
            if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[2 * i]) : "r"(a));
            if (OP == 5) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(r[2 * i]), "+r"(r[2 * i + 1]) : "r"(a));
            if (OP == 6) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
            if (OP == 7) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(r[2 * i]) : "r"(a), "r"(b));
            if (OP == 8) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(r[2 * i]) : "r"(a));
            if (OP == 9) asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((uint64_t)a << 32 | b));
            if (OP == 10) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(r[2*i]), "r"(b));
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < CHAINS; ++i) acc += r[2 * i] + r[2 * i + 1] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32) + (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char *name, int per_iter) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {128, 512, 1024}) {
        k<OP><<<148, threads>>>(out, 12345, cyc);
        cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
        double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
        double warp_instr = (double)ITERS * CHAINS * per_iter * (threads / 32);
        printf("%-34s threads=%4d  cycles=%9.0f  warp-instr/cycle/SM=%6.3f  (lane-ops/clk/SM=%6.1f)\n", name, threads, c, warp_instr / c, 32 * warp_instr / c);
    }
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("IMAD (mad.lo.u32)", 1);
    run<1>("IMAD.WIDE.U32 (64-bit acc)", 1);
    run<2>("IMAD.WIDE pair via mad.lo.cc/madc.hi", 1);
    run<3>("IMAD.HI.U32", 1);
    run<4>("IADD3 (add.u32)", 1);
    run<5>("IADD3 + IADD3.X (add.cc/addc)", 2);
    run<6>("DFMA", 1);
    run<7>("FFMA", 1);
    run<8>("SHF", 1);
    run<9>("add.u64", 1);
    run<10>("mul.wide.u32", 1);
    return 0;
}
